"""-m gpu, needs >= 2 GPUs (skipped otherwise): two ranks over NCCL, batch sharded, bucketed gradient all-reduce
overlapped with backward, Adam chasing the buckets; checked against the oracle on the full batch."""
import os
import socket

import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out, use_graph, grad_dtype, transport):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    eng = None
    try:
        from gan_class_transfer2_b200.engine import DataParallel, NetConfig, UNetEngine, shard_batch
        cfg = O.TINY
        ncfg = NetConfig(size=cfg.size, pixel_size=cfg.pixel_size, max_size=cfg.max_size, octaves=cfg.octaves)
        per = 2
        eng = UNetEngine(ncfg, per, dp=DataParallel(bucket_bytes=1 << 20, grad_dtype=grad_dtype, nccl_ctas=16, transport=transport),
                         use_graph=use_graph)
        eng.load_weights(O.glorot_init(cfg, 0))
        x, t, e = O.synthetic_batch(cfg, per * world, 1)
        lo, hi = shard_batch(per * world, world, rank)
        loss = eng.loss_and_grads(x[lo:hi].cuda(), t[lo:hi].cuda(), e[lo:hi].cuda()).clone()
        grads = {k: v.cpu() for k, v in eng.grads().items()}
        losses = []
        for s in range(4):
            xs, ts, es = O.synthetic_batch(cfg, per * world, 100 + s)
            losses.append(float(eng.train_step(xs[lo:hi].cuda(), ts[lo:hi].cuda(), es[lo:hi].cuda())))
        torch.cuda.synchronize()
        stale_raises = False
        try:
            eng.weights()
        except RuntimeError:
            stale_raises = True  # sharded optimiser: reading stale masters must fail loudly, not return old values
        eng.gather_master_weights()  # ... every rank updates the fp32 masters of its own slice only (collective)
        w = eng.w.clone()
        gathered = [torch.empty_like(w) for _ in range(world)]
        dist.all_gather(gathered, w)
        if rank == 0:
            torch.save({"loss": float(loss), "grads": grads, "losses": losses, "stale_raises": stale_raises,
                        "transport": "p2p" if eng._p2p is not None else "nccl",
                        "multicast": bool(eng._p2p and eng._p2p["g_mc"]),
                        "replicas_equal": all(torch.equal(gathered[0], g) for g in gathered),
                        "weights": {k: v.cpu() for k, v in eng.weights().items()}}, out)
    finally:
        if eng is not None:
            eng.release_graphs()  # a live graph holding captured NCCL collectives blocks the communicator teardown
        dist.destroy_process_group()


@pytest.mark.parametrize("grad_dtype,transport", [("bf16", "p2p"), ("bf16", "nccl"), ("fp32", "nccl")])
@pytest.mark.parametrize("use_graph", [False, True], ids=["eager", "cuda_graph"])
def test_two_gpu_data_parallel_step_matches_oracle(tmp_path, use_graph, grad_dtype, transport):
    """transport "p2p": gradient sum + Keras-Adam + weight broadcast fused in one kernel over NVLink peer memory
    (gct2_adam_apply_p2p, multimem when the fabric offers it); "nccl": reduce-scatter / all-gather collectives."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    import torch.multiprocessing as mp
    from tests import engine_checks as E
    out = str(tmp_path / "dp.pt")
    mp.spawn(_worker, args=(2, _free_port(), out, use_graph, grad_dtype, transport), nprocs=2, join=True)
    got = torch.load(out)
    print("transport used:", got["transport"], "multicast:", got["multicast"])
    cfg = O.TINY
    weights = O.glorot_init(cfg, 0)
    x, t, e = O.synthetic_batch(cfg, 4, 1)
    loss, grads, _ = O.loss_and_grads(weights, x, t, e, cfg)
    assert abs(got["loss"] - float(loss)) <= 1e-3 * float(loss)
    for k, g in grads.items():
        assert E.rel(got["grads"][k], g) <= E.tol_f32_grad(cfg, k), k
    assert got["stale_raises"], "weights() returned stale fp32 masters under the sharded optimiser"
    assert got["replicas_equal"], "ranks diverged: the summed gradients (and so the weights) must be bit-identical"
    tr = O.OracleTrainer(cfg, weights=weights)
    ref = [tr.train_step(*O.synthetic_batch(cfg, 4, 100 + s)) for s in range(4)]
    assert max(abs(a - b) / b for a, b in zip(got["losses"], ref)) <= 1e-3, (got["losses"], ref)
