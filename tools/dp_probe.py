"""Two-rank probe of the captured data-parallel step (debugging aid; needs 2 GPUs).

Runs a few CUDA-graph training steps of the tiny configuration on 2 ranks and prints how far each rank got.  A Python
stack dump fires after --deadline seconds so that a hang shows where the host is waiting.  Toggles come from the
environment like everywhere else (GCT2_DEBUG, GCT2_OVERLAP).

    python tools/dp_probe.py [--deadline 40] [--eager]
"""
import argparse
import faulthandler
import os
import socket
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, use_graph, deadline, steps):
    faulthandler.dump_traceback_later(deadline, exit=True)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from gan_class_transfer2_b200.engine import DataParallel, NetConfig, UNetEngine
    cfg = NetConfig(size=64, pixel_size=128, max_size=256, octaves=4)
    eng = UNetEngine(cfg, 2, dp=DataParallel(bucket_bytes=1 << 20), use_graph=use_graph)
    eng.init_glorot(0)
    x = torch.rand(2, cfg.size, cfg.size, 3, device="cuda") * 2 - 1
    for s in range(steps):
        t0 = time.time()
        loss = eng.train_step(x)
        torch.cuda.synchronize()
        print(f"rank {rank} step {s} loss {float(loss):.6f} t {time.time() - t0:.2f}s", flush=True)
    dist.barrier()
    eng.release_graphs()
    dist.destroy_process_group()
    faulthandler.cancel_dump_traceback_later()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--deadline", type=int, default=40)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--eager", action="store_true")
    a = ap.parse_args()
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), not a.eager, a.deadline, a.steps), nprocs=2, join=True)
    print("probe ok", flush=True)
