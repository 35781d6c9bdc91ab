"""Step-time sweep over engine knobs in ONE process (a fresh `import torch` costs more than the measurements).

    python tools/sweep_step.py --batch 1 --set GCT2_ADAM_SMS=0 --set GCT2_ADAM_SMS=48,GCT2_ADAM_WIDE=5 ...

Every --set is a comma-separated list of environment assignments applied before the engine is built (the engine
reads its knobs in __init__).  Prints one JSON line per configuration: ms per step (CUDA events over `steps` graph
replays after `warmup`), launches per step and the final loss (same inputs everywhere -> comparable)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--set", action="append", default=[])
    a = ap.parse_args()
    from gan_class_transfer2_b200 import _lib
    from gan_class_transfer2_b200.engine import NetConfig, UNetEngine
    lib = _lib.init(0)
    g = torch.Generator().manual_seed(1)
    x = (torch.randint(0, 256, (a.batch, 256, 256, 3), generator=g).float() / 128 - 1).cuda()
    for cfg in a.set or [""]:
        saved = dict(os.environ)
        dbg = []
        for item in filter(None, cfg.split(",")):
            k, v = item.split("=", 1)
            if k.startswith("key"):
                dbg.append((int(k[3:]), int(v)))
            else:
                os.environ[k] = v
        for k, v in dbg:
            lib.gct2_debug_set(k, v)
        eng = UNetEngine(NetConfig(), a.batch, use_graph=True)
        eng.init_glorot(0)
        eng.set_batch(x)
        for _ in range(a.warmup):
            eng.run_step(draw=True)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        s.record()
        for _ in range(a.steps):
            eng.run_step(draw=True)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / a.steps
        print(json.dumps({"config": cfg, "batch": a.batch, "ms_per_step": round(ms, 4), "img_per_s": round(a.batch / ms * 1e3, 1),
                          "launches": eng.launches_per_step(), "loss": float(eng.loss)}), flush=True)
        for k, v in dbg:
            lib.gct2_debug_set(k, {25: 1}.get(k, 0))  # back to the library's defaults
        del eng
        torch.cuda.empty_cache()
        os.environ.clear()
        os.environ.update(saved)


if __name__ == "__main__":
    main()
