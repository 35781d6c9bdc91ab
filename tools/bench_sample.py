"""log_sample's diffusion loops (train.py:365-398, batch 1, t = 1..200; train.py:441-468, batch 6, t = 200..1) on one
B200: denoiser calls per second through Denoiser.sample (one CUDA graph per schedule), with the oracle loop timed on
the host cores beside it for a few steps.

    python tools/bench_sample.py [--cpu-steps 3]
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    from gan_class_transfer2_b200 import train as T
    from oracle import oracle as O
    T.use_cuda_graph = True
    den = T.Denoiser()
    steps = T.steps
    for name, batch, sched in (("forward diffusion (train.py:365-398)", 1, list(range(1, steps + 1))),
                               ("backward diffusion (train.py:441-468)", 6, list(range(steps, 0, -1)))):
        x0, _, e0 = O.synthetic_batch(O.DEFAULT, batch, 3)
        xd, ed = x0.cuda(), e0.cuda()
        den.sample(xd, ed, sched)  # captures the graph
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(a.reps):
            xt, et = den.sample(xd, ed, sched)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / a.reps
        line = {"loop": name, "batch": batch, "denoiser_calls": len(sched), "ms_per_loop": round(ms, 2),
                "calls_per_s": round(len(sched) / ms * 1e3, 1), "images_per_s": round(batch * len(sched) / ms * 1e3, 1),
                "finite": bool(torch.isfinite(xt).all().item())}
        if a.cpu_steps > 0:
            w = {k: v.cpu() for k, v in den.engine(batch, T.size).weights().items()}
            torch.set_num_threads(os.cpu_count() or 1)
            t0 = time.time()
            O.sample_loop(w, x0, e0, sched[:a.cpu_steps], O.DEFAULT)
            cpu_ms = (time.time() - t0) * 1e3 / a.cpu_steps
            line["cpu_oracle_ms_per_call"] = round(cpu_ms, 1)
            line["cpu_cores"] = os.cpu_count()
            line["speedup_vs_cpu_oracle"] = round(cpu_ms / (ms / len(sched)), 1)
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
