#!/usr/bin/env python
"""Headline benchmark: training images/sec of train.py's step (noising -> U-Net fwd -> MSE -> bwd -> Keras-Adam).

    python bench.py --gpus 1 --steps K --warmup W            # this implementation, one B200
    torchrun ... bench.py --gpus N --steps K --warmup W      # data parallel, one rank per GPU (weak scaling)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port) on host cores

Prints ONE JSON line (rank 0).  Workload: BASELINE.json configs[1] -- train.py's default model (256x256x3, 6 octaves,
41.69 M parameters, 128.525 GFLOP per image per step), batch 1 per GPU, synthetic images, glorot-init weights.
  value     images/s, inputs already resident in HBM (x staged in the engine; t_int/eps drawn on the device per step)
  e2e       images/s through the public API (train.Trainer.train_step) with a pinned HOST batch copied in every step
            and the scalar loss read back every step
  roofline  the dominant kernel family (tcgen05 implicit-GEMM convs): algorithmic FLOPs / CUDA-event time of its 33
            launches of one step replayed back to back from a CUDA graph, against the measured sustained bf16 peak
            (MEASURED_PEAKS.json); traffic = DRAM bytes per launch from the committed ncu capture (profiles/)
  cpu_baseline  the oracle (PyTorch-CPU fp32 restatement of train.py) timed on this box's host cores
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GFLOP_PER_IMAGE = 128.525  # SURVEY.md 8(d), default config
METRIC = "train_images_per_sec"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"bf16_burst": p["bf16_tflops"], "bf16_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "hbm": p["hbm_gbs"], "source": "measured"}
    except Exception:  # noqa: BLE001
        return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "source": "fallback"}


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed regions run."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.02)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference(batch: int, steps: int, warmup: int, budget_s: float):
    """Times the oracle's training step (the reference's CPU path as restated in oracle/oracle.py) on host cores."""
    import torch
    from oracle import oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = O.DEFAULT
    tr = O.OracleTrainer(cfg, seed=0)
    batches = [O.synthetic_batch(cfg, batch, 1 + i) for i in range(2)]
    t0 = time.perf_counter()
    tr.train_step(*batches[0])
    first = time.perf_counter() - t0
    done_w = 1
    # bound the run: the oracle needs ~0.5-1 s per image
    total = max(1, min(steps + warmup, int(budget_s / max(first, 1e-3))))
    eff_warm = min(warmup, max(1, total // 4))
    eff_steps = max(1, min(steps, total - eff_warm))
    while done_w < eff_warm:
        tr.train_step(*batches[done_w % 2])
        done_w += 1
    t0 = time.perf_counter()
    for i in range(eff_steps):
        tr.train_step(*batches[i % 2])
    dt = time.perf_counter() - t0
    return {"value": batch * eff_steps / dt, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{eff_steps} full training steps (fwd+loss+bwd+Keras-Adam) of the default model at batch {batch} "
                      f"after {eff_warm} warm-up, PyTorch-CPU fp32 (oneDNN) restatement of train.py -- not TensorFlow",
            "ms_per_step": dt / eff_steps * 1e3, "steps": eff_steps, "warmup": eff_warm}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = args.batch_per_gpu
    r = cpu_reference(batch, args.steps, args.warmup, budget_s=150.0)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "images/s", "n_gpus": args.gpus,
            "steps": r["steps"], "warmup": r["warmup"], "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(batch, 1),
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(batch_per_gpu: int, world: int):
    return {"workload": "train.py default Denoiser U-Net step (size=256, pixel_size=128, max_size=512, octaves=6; "
                        "41,691,660 params; 128.525 GFLOP/image/step), synthetic images, glorot-uniform init",
            "batch_per_gpu": batch_per_gpu, "global_batch": batch_per_gpu * world,
            "parallelism": f"dp{world}" if world > 1 else "single",
            "l2": "per-step working set ~1.4 GB (fp32 w/m/v/g + bf16 shadow + activations) exceeds the 126 MB L2; "
                  "no explicit flush"}


# ------------------------------------------------------------------------------------------------ GPU arm
def synthetic_images(batch: int, size: int, seed: int):
    """SURVEY.md 8(d) synthetic input: x = k/128 - 1, k ~ U{0..255} i.i.d. -- the value grid of decode_file
    (train.py:292).  Same generator call as oracle.synthetic_batch, restated here so that the GPU arm does not import
    the oracle."""
    import torch
    g = torch.Generator().manual_seed(seed)
    k = torch.randint(0, 256, (batch, size, size, 3), generator=g)
    return k.to(torch.float32) / 128 - 1


def kernel_sources_sha() -> str:
    from tools.sass_summary import kernel_sources_sha as f
    return f()


def instrumented_step(eng, torch, ops):
    """One eager step with every op bracketed by CUDA events (the GPU is kept busy first so that host launch latency
    is not inside the brackets).  Returns {op name: [durations in us]}."""
    saved = eng._save_state()
    overlap = (eng.overlap_wgrad, eng.overlap_adam)
    eng.overlap_wgrad = eng.overlap_adam = False  # serialise: per-op durations must be additive
    torch.cuda.synchronize()
    torch.cuda._sleep(int(3e7))
    ops.profile_ops(True)
    eng._step_body(False)
    rec = ops.profile_ops(False)
    torch.cuda.synchronize()
    eng.overlap_wgrad, eng.overlap_adam = overlap
    eng._restore_state(saved)
    out = {}
    for name, s, e in rec:
        out.setdefault(name, []).append(s.elapsed_time(e) * 1e3)
    out["__sequence__"] = [(name, round(s.elapsed_time(e) * 1e3, 1)) for name, s, e in rec]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch-per-gpu", type=int, default=1)
    ap.add_argument("--global-batch", type=int, default=0,
                    help="strong scaling (BASELINE config 3): a fixed global batch split over the ranks")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dump-ops", default="", help="write the instrumented step's per-launch (op, us) list here")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    from gan_class_transfer2_b200 import _lib, ops
    from gan_class_transfer2_b200 import train as T
    from gan_class_transfer2_b200.engine import DataParallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback for the training step "
                         "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        # the collectives run beside backward: cap their CTAs and keep the conv launches off those SMs (engine.DataParallel)
        os.environ.setdefault("NCCL_MAX_CTAS", "16")
        import torch.distributed as dist
        opts = None
        if os.environ.get("GCT2_NCCL_PRIORITY", "1") != "0":
            # the collectives' CTAs go first whenever an SM frees up: they are few, and everything waits for them
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), pg_options=opts)
        T.data_parallel = DataParallel(shard_optimizer=os.environ.get("GCT2_DP_SHARD", "1") != "0")
    _lib.init(local)

    B = args.batch_per_gpu
    scaling = "weak"
    if args.global_batch:
        if args.global_batch % world:
            raise SystemExit(f"--global-batch {args.global_batch} is not divisible by {world} ranks")
        B = args.global_batch // world
        scaling = "strong"
    T.use_cuda_graph = not args.no_graph
    denoiser = T.Denoiser()
    trainer = T.Trainer(denoiser)
    trainer.compile(T.optimizer, T.identity)
    eng = denoiser.engine(B, T.size)

    x_cpu = synthetic_images(B, T.size, 1 + rank)  # the same images the CPU arm sees; nothing of oracle/ is imported here
    x_host = x_cpu.pin_memory()
    loss_host = torch.zeros(1).pin_memory()
    eng.set_batch(x_host)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if dist is None:
            return ms
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- warm-up (also captures the CUDA graph)
    for _ in range(args.warmup):
        eng.run_step(draw=True)
    barrier()
    launches_per_step = eng.launches_per_step()

    sampler = ClockSampler(local)
    sampler.start()

    # ---- value: K steps, inputs resident in HBM
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    s.record()
    for _ in range(args.steps):
        eng.run_step(draw=True)
    e.record()
    barrier()
    ms_total = max_over_ranks(s.elapsed_time(e))
    ms_per_step = ms_total / args.steps
    value = B * world * args.steps / (ms_total * 1e-3)

    # ---- data parallel: correctness bit and communication breakdown
    comm = None
    if dist is not None:
        # every rank must hold bit-identical bf16 weights after the timed loop (summed gradients are identical everywhere)
        mx, mn = eng.w16.clone(), eng.w16.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        replicas_equal = bool(torch.equal(mx, mn))
        del mx, mn

        def timed(fn, n):
            for _ in range(3):
                fn()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n):
                fn()
            b.record()
            barrier()
            return max_over_ranks(a.elapsed_time(b)) / n

        # (a) the same step without its collectives (wrong numbers, right compute): what the GPU alone needs
        saved = eng._save_state()
        graphs = eng._graph
        eng._graph, T.data_parallel.dry_run = None, True
        try:
            compute_ms = timed(lambda: eng.run_step(draw=True), max(10, args.steps // 4))
        finally:
            eng.release_graphs()
            eng._graph, T.data_parallel.dry_run = graphs, False
            eng._restore_state(saved)
        # (b) the step's collectives alone, back to back: per bucket reduce-scatter (+ all-gather of the bf16 weights)
        from gan_class_transfer2_b200.engine import optimizer_shard
        dp = T.data_parallel

        head_p2p = [False]

        def comm_only():
            used_p2p = False
            for start, end, _ in eng._buckets:
                cut = optimizer_shard(start, end, eng.small, dp.world, dp.rank) if dp.shard_optimizer else None
                if cut is not None and eng._p2p is not None and (cut[3] - cut[2]) % 8 == 0 and cut[2] % 8 == 0:
                    # the fused kernel IS the exchange: cross-rank barrier + gradient sum / Adam / weight broadcast
                    lo, _, own, own_hi = cut
                    pp = eng._p2p
                    pp["hg"].barrier(channel=0)
                    ops.adam_apply_p2p(eng.w[own:own_hi], eng.m[own:own_hi], eng.v[own:own_hi], pp["g_ptrs"], pp["w_ptrs"],
                                       dp.world, own, eng.hyper, eng.cfg.beta1, eng.cfg.beta2, eng.cfg.epsilon, 1.0, True,
                                       pp["g_mc"], pp["w_mc"])
                    used_p2p = True
                    if start >= eng.small:
                        continue
                    if pp.get("head_p2p"):
                        # the head region and the loss: staging copies + one peer-load sum, no NCCL call
                        pp["head"][:eng.small].copy_(eng.g[:eng.small], non_blocking=True)
                        pp["head"][eng.small:eng.small + 1].copy_(eng.loss, non_blocking=True)
                        ops.sum_peers_f32(pp["head_ptrs"], dp.world, eng.g[:eng.small], eng.loss)
                        head_p2p[0] = True
                        continue
                    end = eng.small
                    cut = None
                if cut is not None:
                    lo, _, own, own_hi = cut
                    gsrc = eng.g16 if eng.g16 is not None else eng.g
                    dist.reduce_scatter_tensor(gsrc[own:own_hi], gsrc[lo:end])
                    dist.all_gather_into_tensor(eng.w16[lo:end], eng.w16[own:own_hi])
                    if start >= eng.small:
                        continue
                    end = eng.small
                dist.all_reduce(eng.g[start:end])
            if used_p2p:
                eng._p2p["hw"].barrier(channel=1)
            if not head_p2p[0]:
                dist.all_reduce(eng.loss)  # the reported loss (global mean)

        saved = eng._save_state()
        comm_ms = timed(comm_only, max(10, args.steps // 4))
        eng._restore_state(saved)
        exposed = max(0.0, ms_per_step - compute_ms)
        grad_bytes = (2 if eng.g16 is not None else 4) * eng.P
        comm = {"comm_exposed_ms": exposed, "comm_overlapped_ms": max(0.0, comm_ms - exposed), "comm_alone_ms": comm_ms,
                "compute_only_ms": compute_ms, "replicas_bit_equal": replicas_equal,
                "bytes_per_step": {"reduce_scatter": grad_bytes, "all_gather_bf16": 2 * eng.P},
                "grad_dtype": dp.grad_dtype, "shard_optimizer": dp.shard_optimizer, "nccl_max_ctas": dp.nccl_ctas,
                "transport": ("p2p+multimem" if eng._p2p["g_mc"] else "p2p") if eng._p2p is not None else "nccl",
                "head_and_loss": "peer loads" if (eng._p2p is not None and eng._p2p.get("head_p2p")) else "nccl all-reduce",
                "how": "compute_only = the captured step with its collectives left out; comm_alone = the step's "
                       "collectives back to back; exposed = step - compute_only; overlapped = comm_alone - exposed"}
        if not replicas_equal:
            raise SystemExit("data-parallel ranks hold different bf16 weights after the timed loop")

    # ---- e2e: public API, pinned host batch in, loss out, every step
    for _ in range(3):
        trainer.train_step((x_host, x_host))
    barrier()
    s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s2.record()
    for _ in range(args.steps):
        loss = trainer.train_step((x_host, x_host))["loss"]
        loss_host.copy_(loss.reshape(1), non_blocking=True)
    e2.record()
    barrier()
    ms_e2e = max_over_ranks(s2.elapsed_time(e2))
    # the same with the batch as decode_file's uint8 bytes (a quarter of the H2D traffic; decode fused into the prologue)
    u8_host = ((x_cpu + 1) * 128).round().clamp(0, 255).to(torch.uint8).pin_memory()
    for _ in range(3):
        trainer.train_step((u8_host, u8_host))
    barrier()
    s4, e4 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s4.record()
    for _ in range(args.steps):
        loss = trainer.train_step((u8_host, u8_host))["loss"]
        loss_host.copy_(loss.reshape(1), non_blocking=True)
    e4.record()
    barrier()
    ms_e2e_u8 = max_over_ranks(s4.elapsed_time(e4))
    clocks = sampler.stop()
    final_loss = float(loss_host.item())

    # ---- roofline of the dominant kernel family, measured live
    # (a) the family alone: its 33 launches of one step captured into a CUDA graph (same plans and launch chaining as
    #     in the step, nothing else running), replayed and timed with CUDA events -> average launch duration;
    # (b) one eager step with every op bracketed by events -> each op's share of the step (serialised).
    pk = peaks()
    fam_graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        eng.conv_family_pass()
    torch.cuda.current_stream().wait_stream(side)
    before = ops.launch_count()
    with torch.cuda.graph(fam_graph):
        eng.conv_family_pass()
    fam_launches = ops.launch_count() - before
    for _ in range(3):
        fam_graph.replay()
    reps = 50
    s3, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s3.record()
    for _ in range(reps):
        fam_graph.replay()
    e3.record()
    torch.cuda.synchronize()
    conv_us = s3.elapsed_time(e3) * 1e3 / reps
    del fam_graph
    prof = instrumented_step(eng, torch, ops)
    sequence = prof.pop("__sequence__")
    if args.dump_ops and rank == 0:
        with open(args.dump_ops, "w") as f:
            json.dump(sequence, f)
    conv_ops = [k for k in prof if k.startswith("conv") and "c3" not in k]
    conv_us_eager = sum(sum(prof[k]) for k in conv_ops)
    conv_launches = sum(len(prof[k]) for k in conv_ops)
    total_us = sum(sum(v) for v in prof.values())
    # FLOPs of the tensor-core family = step total minus down0 (fprop+wgrad, CUDA cores) and dense (fwd+2 bwd)
    flops_conv = (GFLOP_PER_IMAGE - 2 * 0.2013 - 3 * 0.0263) * 1e9 * B
    achieved = flops_conv / (conv_us * 1e-6) / 1e12 if conv_us else 0.0
    # DRAM bytes per launch come from an ncu capture (never taken inside a timed run); the file names the kernel
    # sources it was captured from, and a capture of other sources is not reported
    traffic, traffic_note = None, "no capture"
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("kernel_sha") == kernel_sources_sha():
            traffic, traffic_note = tj.get(f"traffic_bytes_per_launch_b{B}"), tj.get("source")
        else:
            traffic_note = f"stale capture (kernels {tj.get('kernel_sha')}, now {kernel_sources_sha()})"
    except Exception as exc:  # noqa: BLE001
        traffic_note = f"unreadable: {exc}"
    # the optimiser: the step's largest HBM consumer (30 bytes per parameter: w, m, v, g read; w, m, v, bf16 shadow written)
    adam_us = sum(prof.get("adam_apply", [])) + sum(prof.get("adam_keras", []))
    adam_bytes = 30.0 * eng.P / world if (world > 1 and T.data_parallel.shard_optimizer) else 30.0 * eng.P
    adam_gbs = adam_bytes / (adam_us * 1e-6) / 1e9 if adam_us else 0.0
    roofline = {"bound": "tensor", "kernel": "conv_umma_kernel<MODE,BN> (tcgen05 implicit-GEMM conv family, "
                                             f"{conv_launches} launches/step)",
                "achieved": achieved, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                "frac": achieved / pk["bf16_sustained"], "frac_burst": achieved / pk["bf16_burst"],
                "peak_burst": pk["bf16_burst"], "traffic": traffic, "traffic_source": traffic_note,
                "peak_source": pk["source"] + " (sustained; frac_burst is against the burst figure)",
                "how": f"{flops_conv / 1e9:.1f} GFLOP of the family per step / CUDA-event time of its launches replayed "
                       f"back to back from a graph ({fam_launches} launches incl. split-K passes, {reps} replays)",
                "avg_launch_us": conv_us / max(conv_launches, 1), "family_us_per_step": conv_us,
                "family_share_of_step": conv_us_eager / max(total_us, 1e-9),
                "step_achieved": value / world * GFLOP_PER_IMAGE * 1e9 / 1e12,
                "step_frac": value / world * GFLOP_PER_IMAGE * 1e9 / 1e12 / pk["bf16_sustained"],
                "per_op_us": {k: round(sum(v), 1) for k, v in prof.items()},
                "kernel_sha": kernel_sources_sha()}
    roofline_adam = {"bound": "hbm", "kernel": "adam_kernel (Keras-Adam, fp32 state + bf16 shadow)", "achieved": adam_gbs,
                     "peak": pk["hbm"], "unit": "GB/s", "frac": adam_gbs / pk["hbm"], "traffic": None,
                     "how": f"{adam_bytes / 1e6:.0f} MB algorithmic (30 B/param) / CUDA-event time of the optimiser launches "
                            "of one serialised eager step", "us_per_step": adam_us}

    def teardown():
        """Captured NCCL collectives keep the communicator busy: drop the graphs first, and never let a stuck
        communicator teardown keep the ranks (and the driver's clock) waiting."""
        if dist is None:
            return
        sys.stdout.flush()
        eng.release_graphs()
        barrier()
        timer = threading.Timer(20.0, lambda: os._exit(0))
        timer.daemon = True
        timer.start()
        dist.destroy_process_group()
        timer.cancel()

    if rank != 0:
        teardown()
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference(B, 4, 1, budget_s=25.0)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
    line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(B, world),
            "e2e": {"value": B * world * args.steps / (ms_e2e * 1e-3), "unit": "images/s",
                    "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps, "api": "train.Trainer.train_step((x_host, x_host))"},
            "e2e_uint8": {"value": B * world * args.steps / (ms_e2e_u8 * 1e-3), "unit": "images/s",
                          "h2d_bytes_per_step": u8_host.numel(), "d2h_bytes_per_step": 4,
                          "api": "train.Trainer.train_step((img_u8_host, img_u8_host)) -- decode_file's bytes, "
                                 "/128-1 on the device (SURVEY 8 f2)"},
            "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step,
            "cuda_graph": not args.no_graph, "clocks": clocks, "roofline": roofline, "roofline_adam": roofline_adam,
            "cpu_baseline": cpu, "comm": comm,
            "final_loss": final_loss}
    print(json.dumps(line), flush=True)
    teardown()


if __name__ == "__main__":
    main()
