"""Per-CTA phase timeline of the tensor-core conv launches (test hook gct2_debug_set(7,1) / gct2_debug_timeline).

For every (layer, pass) of the default model at the given batch: runs the op a few times, then reads the
%globaltimer stamps of the last launch and prints, in microseconds relative to the first CTA's entry:
  span      last CTA done - first CTA entry (the kernel's device time)
  entry     spread of CTA entry times (launch ramp)
  prologue  barrier init + TMEM alloc + (cluster) sync
  fill      prologue done -> first operands landed (TMA latency)
  mma       first operands -> first accumulator complete (main loop of the first tile)
  epi       first accumulator complete -> first epilogue done
  rest      first epilogue done -> CTA done (further tiles of a persistent CTA + teardown)
"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
# the production library carries no stamps in its main loops: use the -DGCT2_TIMELINE build (make -C csrc timeline)
_TL = os.path.join(ROOT, "gan_class_transfer2_b200", "libgct2_b200_timeline.so")
if "GCT2_LIB" not in os.environ:
    import subprocess
    subprocess.run(["make", "-C", os.path.join(ROOT, "gan_class_transfer2_b200", "csrc"), "timeline"], check=True,
                   capture_output=True)  # no-op when the library is newer than the sources
    os.environ["GCT2_LIB"] = _TL

from gan_class_transfer2_b200 import _lib, ops  # noqa: E402
from tools.bench_layers import layer_table  # noqa: E402


def read_timeline(lib):
    buf = (ctypes.c_ulonglong * (512 * 8))()
    n = lib.gct2_debug_timeline(buf, 512)
    return np.frombuffer(buf, dtype=np.uint64)[: n * 8].reshape(n, 8).astype(np.int64)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--only", default="")
    ap.add_argument("--debug", action="append", default=[])
    ap.add_argument("--cfg", action="append", default=[],
                    help='several plans in one process: "3=64 4=8 25=2" (space-separated gct2_debug_set key=value)')
    ap.add_argument("--passes", default="fprop,dgrad,wgrad")
    ap.add_argument("--weights-stable", action="store_true", help="pass GCT2_WEIGHTS_STABLE (early weight fetch)")
    ap.add_argument("--cold-weights", action="store_true",
                    help="flush L2 before the measured launch, then re-touch the activations only: the state a layer "
                         "finds inside a training step (the optimiser has streamed 1.3 GB through L2 since)")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    lib = _lib.init(0)
    for d in a.debug:
        k, v = (int(t) for t in d.split("="))
        lib.gct2_debug_set(k, v)
    lib.gct2_debug_set(7, 1)
    lib.gct2_debug_set(2, 1)
    B = a.batch
    for cfg in (a.cfg or [""]):
        keys = [tuple(int(t) for t in kv.split("=")) for kv in cfg.split()]
        for k, v in keys:
            lib.gct2_debug_set(k, v)
        if a.cfg:
            print(json.dumps({"cfg": cfg}), flush=True)
        run(a, lib, dev, B)
        for k, v in keys:
            lib.gct2_debug_set(k, 0)


def run(a, lib, dev, B):
    g = torch.Generator(device=dev).manual_seed(0)
    for name, kind, Hin, Cin, Cout in layer_table(256, 128, 512, 6):
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
        if kind == "down":
            passes = {"fprop": lambda: ops.conv4s2_fprop(x, w, bias, y, ws, a.weights_stable),
                      "dgrad": lambda: ops.conv4s2_dgrad(dy, w, dx, x, False, ws, a.weights_stable),
                      "wgrad": lambda: ops.conv4s2_wgrad(x, dy, dw, ws)}
        else:
            passes = {"fprop": lambda: ops.convT4s2_fprop(x, w, bias, y, ws, a.weights_stable),
                      "dgrad": lambda: ops.convT4s2_dgrad(dy, w, dx, x, Cin, ws, a.weights_stable),
                      "wgrad": lambda: ops.convT4s2_wgrad(x, dy, dw, ws)}
        for pname, fn in passes.items():
            if pname not in a.passes.split(","):
                continue
            lib.gct2_debug_set(2, 0)
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            lib.gct2_debug_set(2, 1)
            sys.stderr.flush()
            if a.cold_weights:
                flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
                flush.fill_(1)
                for t_ in (x, dy, y, dx):
                    t_.float().sum()
                torch.cuda.synchronize()
            fn()
            t = read_timeline(lib)
            t0 = t[:, 0].min()
            rel = (t - t0) / 1e3
            ok = t[:, 6] > 0
            # BAR.SYNC defers its blocking to the next consumer, so thread 0's "CTA done" stamp [6] can be taken before
            # the epilogue warps arrive: a CTA is done no earlier than its first epilogue [5]
            rel[:, 6] = np.maximum(rel[:, 6], rel[:, 5])

            def st(col_a, col_b):
                d = rel[ok, col_b] - rel[ok, col_a]
                return round(float(d.mean()), 2), round(float(d.max()), 2)

            print(json.dumps({"layer": name, "pass": pname, "ctas": int(t.shape[0]),
                              "span": round(float(rel[ok, 6].max()), 2),
                              "entry_spread": round(float(rel[ok, 0].max()), 2),
                              "prologue": st(0, 1), "fill": st(1, 2), "issue": st(1, 7), "mma_issue": st(2, 3), "mma": st(2, 4), "epi": st(4, 5), "rest": st(5, 6),
                              "cta_life": st(0, 6)}), flush=True)


if __name__ == "__main__":
    main()
