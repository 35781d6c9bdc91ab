"""CPU, world_size 2 over gloo: the host-side logic of the data-parallel path -- batch sharding, 1/N_global loss
scaling, flat-buffer packing and bucket-by-bucket all-reduce in backward order -- reproduces the single-process
full-batch gradient.  The compute inside each rank is the oracle (the CUDA engine needs a GPU); the sharding, layout and
bucket code under test is the product's (gan_class_transfer2_b200.engine)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gan_class_transfer2_b200 import engine as E
from oracle import oracle as O

OCFG = O.Config(size=16, pixel_size=4, max_size=16, octaves=2)
NCFG = E.NetConfig(size=16, pixel_size=4, max_size=16, octaves=2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, global_batch, bucket_bytes, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        weights = O.glorot_init(OCFG, 0)
        x, t, e = O.synthetic_batch(OCFG, global_batch, 1)
        lo, hi = E.shard_batch(global_batch, world, rank)
        loss, grads, _ = O.loss_and_grads(weights, x[lo:hi], t[lo:hi], e[lo:hi], OCFG, global_elems=x.numel())
        offsets, total = E.param_offsets(NCFG)
        flat = torch.zeros(total)
        for name, (off, cnt) in offsets.items():
            flat[off:off + cnt] = grads[name].reshape(-1)
        buckets = E.grad_buckets(NCFG, bucket_bytes)
        works = [dist.all_reduce(flat[start:end], async_op=True) for start, end, _ in buckets]  # tail-to-head order
        for w in works:
            w.wait()
        loss_t = loss.reshape(1).clone()
        dist.all_reduce(loss_t)
        if rank == 0:
            torch.save({"flat": flat, "loss": loss_t, "buckets": buckets}, out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("bucket_bytes", [1 << 8, 1 << 12, 1 << 30])
def test_two_rank_bucketed_allreduce_matches_full_batch(tmp_path, bucket_bytes):
    out = str(tmp_path / "dp.pt")
    mp.spawn(_worker, args=(2, _free_port(), 4, bucket_bytes, out), nprocs=2, join=True)
    got = torch.load(out)
    weights = O.glorot_init(OCFG, 0)
    x, t, e = O.synthetic_batch(OCFG, 4, 1)
    loss, grads, _ = O.loss_and_grads(weights, x, t, e, OCFG)
    offsets, total = E.param_offsets(NCFG)
    assert abs(float(got["loss"]) - float(loss)) <= 1e-5 * abs(float(loss))
    for name, (off, cnt) in offsets.items():
        assert torch.allclose(got["flat"][off:off + cnt], grads[name].reshape(-1), rtol=1e-4, atol=1e-8), name
    b = got["buckets"]
    assert b[0][1] == total and b[-1][0] == 0 and all(b[i][0] == b[i + 1][1] for i in range(len(b) - 1))
