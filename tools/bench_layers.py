"""Config 5 of BASELINE.json: isolated conv fprop/dgrad/wgrad microbench over train.py's layer shapes.

Prints one JSON line per (layer, pass): CUDA-event time, TFLOP/s, fraction of the measured bf16 peak
(MEASURED_PEAKS.json, burst figure: kernels are timed alone) and the HBM-side fraction, so each layer can be read
against the bound that applies to it.  An L2 flush (256 MB write) runs between timed
launches unless --no-flush.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gan_class_transfer2_b200 import ops  # noqa: E402


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p["bf16_tflops"], p["hbm_gbs"], "measured"
    except Exception:  # noqa: BLE001
        return 1590.0, 6650.0, "fallback"


def layer_table(size, pixel_size, max_size, octaves):
    rows = []
    cin = 3
    for i in range(octaves):
        co = min(pixel_size * 2 ** i, max_size)
        rows.append(("down%d" % i, "down", size >> i, cin, co))
        cin = co
    for i in reversed(range(octaves)):
        ci = min(pixel_size * 2 ** i, max_size) if i == octaves - 1 else \
            min(pixel_size * 2 ** (i + 1) // 2, max_size) + min(pixel_size * 2 ** i, max_size)
        co = min(pixel_size * 2 ** i // 2, max_size)
        rows.append(("up%d" % i, "up", size >> (i + 1), ci, co))
    return rows


def time_fn(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        if flush is not None:
            flush.fill_(1.0)
        torch.cuda._sleep(200000)  # keep the GPU busy while the host prepares the launch (tensor maps, ctypes)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        tot += s.elapsed_time(e)
    return tot / iters * 1e3  # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--pixel-size", type=int, default=128)
    ap.add_argument("--max-size", type=int, default=512)
    ap.add_argument("--octaves", type=int, default=6)
    ap.add_argument("--only", default="")
    ap.add_argument("--debug", action="append", default=[], help="key=value for gct2_debug_set (2=verbose plans, "
                    "3=BN, 4=splits, 5=cluster M, 6=cluster N)")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    from gan_class_transfer2_b200 import _lib
    lib = _lib.init(0)
    for d in a.debug:
        k, v = (int(t) for t in d.split("="))
        lib.gct2_debug_set(k, v)
    tf_peak, hbm_peak, src = peaks()
    flush = None if a.no_flush else torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    B = a.batch
    g = torch.Generator(device=dev).manual_seed(0)
    for name, kind, Hin, Cin, Cout in layer_table(a.size, a.pixel_size, a.max_size, a.octaves):
        if Cin == 3 or (a.only and a.only not in name):
            continue
        Hout = Hin // 2 if kind == "down" else Hin * 2
        x = torch.randn(B, Hin, Hin, Cin, device=dev, generator=g).to(torch.bfloat16)
        y = torch.empty(B, Hout, Hout, Cout, device=dev, dtype=torch.bfloat16)
        dy = torch.randn(B, Hout, Hout, Cout, device=dev, generator=g).to(torch.bfloat16)
        dx = torch.empty_like(x)
        wshape = (4, 4, Cin, Cout) if kind == "down" else (4, 4, Cout, Cin)
        w = (torch.randn(wshape, device=dev, generator=g) * 0.02).to(torch.bfloat16)
        dw = torch.empty(wshape, device=dev, dtype=torch.float32)
        bias = torch.zeros(Cout, device=dev)
        ws = ops.Workspace(256 << 20, dev)
        lo = min(Hin, Hout)
        flops = 2.0 * B * lo * lo * 16 * Cin * Cout
        act_bytes = 2.0 * (x.numel() + y.numel())
        if kind == "down":
            passes = {"fprop": lambda: ops.conv4s2_fprop(x, w, bias, y, ws),
                      "dgrad": lambda: ops.conv4s2_dgrad(dy, w, dx, x, False, ws),
                      "wgrad": lambda: ops.conv4s2_wgrad(x, dy, dw, ws)}
        else:
            passes = {"fprop": lambda: ops.convT4s2_fprop(x, w, bias, y, ws),
                      "dgrad": lambda: ops.convT4s2_dgrad(dy, w, dx, x, Cin, ws),
                      "wgrad": lambda: ops.convT4s2_wgrad(x, dy, dw, ws)}
        for pname, fn in passes.items():
            us = time_fn(fn, a.iters, flush)
            algo_bytes = act_bytes + (4.0 if pname == "wgrad" else 2.0) * w.numel()
            t_tc = flops / (tf_peak * 1e12) * 1e6
            t_hbm = algo_bytes / (hbm_peak * 1e9) * 1e6
            print(json.dumps({"layer": name, "pass": pname, "B": B, "us": round(us, 2),
                              "tflops": round(flops / us * 1e-6, 1), "tc_frac": round(t_tc / us, 3),
                              "hbm_frac": round(t_hbm / us, 3), "bound": "tensor" if t_tc > t_hbm else "hbm",
                              "roofline_us": round(max(t_tc, t_hbm), 2), "peaks": src}), flush=True)


if __name__ == "__main__":
    main()
