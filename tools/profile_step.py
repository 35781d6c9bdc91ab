"""One training step of the default model inside cudaProfilerStart/Stop, for `ncu --profile-from-start off`.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv \
        python tools/profile_step.py [--batch 1]
    ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_umma -o prof \
        python tools/profile_step.py

Eager launches (no CUDA graph, streams as in the real step); three warm-up steps run before the profiled one.
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    a = ap.parse_args()
    from gan_class_transfer2_b200.engine import NetConfig, UNetEngine
    eng = UNetEngine(NetConfig(), a.batch, use_graph=False)
    eng.init_glorot(0)
    x = torch.rand(a.batch, 256, 256, 3, device="cuda") * 2 - 1
    eng.set_batch(x)
    for _ in range(3):
        eng.run_step(draw=True)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    eng.run_step(draw=True)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("loss", float(eng.loss))


if __name__ == "__main__":
    main()
