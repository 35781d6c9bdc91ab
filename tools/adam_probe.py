"""How much HBM bandwidth Keras-Adam draws from K SMs (gct2_debug_set key 23: one SM-exclusive 1024-thread CTA per SM).

The optimiser moves 30 bytes per parameter (w, m, v, g read; w, m, v and the bf16 shadow written) and needs no tensor
cores; the convs of the same step need no HBM.  This probe answers how many SMs the optimiser must be given to hide
behind them.  Prints one JSON line per K (K = 0: the default full-chip launch).
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gan_class_transfer2_b200 import _lib, ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--params", type=int, default=41_691_648)
    ap.add_argument("--sms", default="0,8,16,24,32,48,64,96,148")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    lib = _lib.init(0)
    n = a.params // 4 * 4
    w = torch.randn(n, device=dev)
    m = torch.zeros(n, device=dev)
    v = torch.zeros(n, device=dev)
    g = torch.randn(n, device=dev) * 1e-3
    wb = torch.zeros(n, dtype=torch.bfloat16, device=dev)
    hyper = torch.tensor([1e-5, 1e-5], device=dev)
    for k in (int(t) for t in a.sms.split(",")):
        lib.gct2_debug_set(23, k)
        for _ in range(3):
            ops.adam_apply(w, m, v, g, wb, hyper)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        s.record()
        reps = 10
        for _ in range(reps):
            ops.adam_apply(w, m, v, g, wb, hyper)
        e.record()
        torch.cuda.synchronize()
        us = s.elapsed_time(e) * 1e3 / reps
        gbs = n * 30 / (us * 1e-6) / 1e9
        print(json.dumps({"adam_sms": k, "us": round(us, 1), "GBps": round(gbs, 1),
                          "GBps_per_sm": round(gbs / k, 1) if k else None}), flush=True)
    lib.gct2_debug_set(23, 0)


if __name__ == "__main__":
    main()
