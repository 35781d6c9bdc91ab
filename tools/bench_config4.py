"""BASELINE.json configs[3]: the doubled-resolution / widened-channel variant (size=512, pixel_size=256, max_size=1024,
octaves=7: 217,078,796 parameters, 2062.2 GFLOP per image per training step) -- training images/s on
1 GPU, or data parallel under torchrun (one rank per GPU, weak scaling).

    python tools/bench_config4.py [--batch-per-gpu 1] [--steps 30]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_config4.py
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
GFLOP_PER_IMAGE = 2062.2


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch-per-gpu", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    from gan_class_transfer2_b200.engine import DataParallel, NetConfig, UNetEngine, param_offsets
    dp = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dp = DataParallel()
    cfg = NetConfig(size=512, pixel_size=256, max_size=1024, octaves=7)
    eng = UNetEngine(cfg, a.batch_per_gpu, dp=dp, use_graph=True)
    eng.init_glorot(0)
    x = torch.rand(a.batch_per_gpu, 512, 512, 3, device="cuda") * 2 - 1
    eng.set_batch(x)
    for _ in range(a.warmup):
        eng.run_step(draw=True)
    torch.cuda.synchronize()
    if dp:
        dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(a.steps):
        eng.run_step(draw=True)
    e.record()
    torch.cuda.synchronize()
    ms = torch.tensor([s.elapsed_time(e)], device="cuda")
    if dp:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    if rank == 0:
        ips = a.batch_per_gpu * world * a.steps / (ms * 1e-3)
        print(json.dumps({"config": "size=512 pixel_size=256 max_size=1024 octaves=7", "params": param_offsets(cfg)[1],
                          "n_gpus": world, "batch_per_gpu": a.batch_per_gpu, "images_per_s": round(ips, 2),
                          "ms_per_step": round(ms / a.steps, 3),
                          "tflops_per_gpu": round(ips / world * GFLOP_PER_IMAGE / 1e3, 1),
                          "loss": float(eng.loss), "finite": bool(torch.isfinite(eng.loss).all().item())}), flush=True)
    if dp:
        eng.release_graphs()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
