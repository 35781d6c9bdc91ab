"""Whole-step launch timeline (test hook gct2_debug_set(11,1) / gct2_debug_trace): which launches overlap in time.

Runs the default model's training step (CUDA graph) a few times, traces ONE step and prints every launch of this
library with start / end (us, relative to the first launch), duration and how many other launches were running at
its start.  --csv writes the records for profiles/.
"""
import argparse
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

NAMES = {1: "noise", 2: "step_begin", 3: "down0_fprop", 4: "down0_wgrad", 5: "dense_mse", 6: "bias_grads", 7: "adam_prepare",
         8: "adam", 9: "cast_bf16", 20: "splitk_finish", 21: "wgrad_reduce"}
for mode, mname in ((0, "convS"), (1, "convP"), (2, "convW")):
    for bn in (64, 128, 256):
        NAMES[100 + mode * 10 + bn // 64] = f"{mname}<{bn}>"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--csv", default="")
    ap.add_argument("--no-graph", action="store_true")
    a = ap.parse_args()
    from gan_class_transfer2_b200 import _lib
    from gan_class_transfer2_b200.engine import NetConfig, UNetEngine
    lib = _lib.init(0)
    eng = UNetEngine(NetConfig(), a.batch, use_graph=not a.no_graph)
    eng.init_glorot(0)
    x = torch.rand(a.batch, 256, 256, 3, device="cuda") * 2 - 1
    eng.set_batch(x)
    for _ in range(5):
        eng.run_step(draw=True)
    torch.cuda.synchronize()
    lib.gct2_debug_set(11, 1)
    eng.run_step(draw=True)
    buf = (ctypes.c_ulonglong * (8192 * 4))()
    n = lib.gct2_debug_trace(buf, 8192)
    lib.gct2_debug_set(11, 0)
    rec = np.frombuffer(buf, dtype=np.uint64)[: n * 4].reshape(n, 4).astype(np.int64)
    ids, blk, grid = rec[:, 0] & 0xFFFF, rec[:, 1] & 0xFFFFFFFF, rec[:, 1] >> 32
    waited = rec[:, 0] >> 16  # ns between block entry and the return of griddepcontrol.wait (conv kernels only)
    t0 = rec[:, 2].min()
    # one launch = its first-block record (+ last-block record when the grid has more than one block)
    order = np.argsort(rec[:, 2])
    used = np.zeros(n, bool)
    launches = []
    for i in order:
        if used[i] or blk[i] != 0:
            continue
        used[i] = True
        start, end = rec[i, 2], rec[i, 3]
        ready = rec[i, 2] + waited[i] if waited[i] else 0
        if grid[i] > 1:
            cand = [j for j in order if not used[j] and ids[j] == ids[i] and grid[j] == grid[i] and blk[j] == grid[i] - 1]
            if cand:
                j = min(cand, key=lambda j: abs(rec[j, 2] - rec[i, 2]))
                used[j] = True
                start, end = min(start, rec[j, 2]), max(end, rec[j, 3])
                if waited[j]:
                    ready = max(ready, rec[j, 2] + waited[j])
        launches.append((int(ids[i]), int(grid[i]), (start - t0) / 1e3, (end - t0) / 1e3,
                         (ready - t0) / 1e3 if ready else -1.0))
    launches.sort(key=lambda r: r[2])
    lines = ["kernel,grid,start_us,end_us,dur_us,running_at_start,ready_us,work_us"]
    busy = 0.0
    last_end = 0.0
    for k, (kid, g, s, e, r) in enumerate(launches):
        running = sum(1 for (_, _, s2, e2, _) in launches if s2 < s < e2)
        tail = f",{r:.1f},{e - r:.1f}" if r >= 0 else ",,"
        lines.append(f"{NAMES.get(kid, kid)},{g},{s:.1f},{e:.1f},{e - s:.1f},{running}{tail}")
        if e > last_end:
            busy += e - max(s, last_end)
            last_end = e
    total = max(e for (_, _, _, e, _) in launches)
    print("\n".join(lines))
    print(f"# {len(launches)} launches, step span {total:.1f} us, union of kernel intervals {busy:.1f} us "
          f"(idle {total - busy:.1f} us), sum of durations {sum(e - s for (_, _, s, e, _) in launches):.1f} us")
    if a.csv:
        with open(a.csv, "w") as f:
            f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
