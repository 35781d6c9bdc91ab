"""Step engine: owns the HBM-resident state of the reference's training step and sequences the kernels.

What it replaces: everything Keras does for ``trainer.fit`` per step in the reference (train.py:516-523) --
``Trainer.call`` (:223-272), ``Denoiser.call`` (:206-215), the GradientTape backward and ``Adam.apply_gradients``
(:75) -- as one fixed launch sequence over pre-allocated buffers, optionally captured into a CUDA graph.

Data layout in HBM (B = per-GPU batch, S = size, n = octaves; all activations NHWC bf16):
  noised            fp32 [B,S,S,3]                        train.py:231-234
  cat[j], j=1..n-1  bf16 [B,S/2^j,S/2^j, up_c[j]+down_c[j-1]]   the tf.concat of Residual (train.py:113-119) is never
                    materialised: up_j writes channels [0,up_c[j]), down_{j-1} writes the rest
  bot               bf16 [B,S/2^n,S/2^n,down_c[n-1]]      bottleneck (down_{n-1} output)
  u0                bf16 [B,S,S,up_c[0]]                  up_0 output; Dense(3) reads u0 and noised separately
  gcat[j], gbot, gu0   same shapes: gradients w.r.t. the *pre-activation* of the producing layer (ReLU mask applied)
  w, m, v, g        fp32 flat [P]: a small head region (down0's kernel, all biases, Dense) followed by the
                    tensor-core kernels down1..down{n-1}, up{n-1}..up0 (see param_offsets)
  w16               bf16 flat [P] shadow copy read by the tensor-core kernels (rewritten by the Adam kernel)
Backward completes the flat gradient buffer from its end to its start (up0..up{n-1}, down{n-1}..down1, small region),
so the data-parallel buckets are contiguous tail-to-head ranges that become ready in order.
"""
from __future__ import annotations

import dataclasses
import math
import os
from typing import Dict, List, Optional, Tuple

import torch

from . import ops


@dataclasses.dataclass(frozen=True)
class NetConfig:
    """Hyper-parameters of train.py:17-36 that shape the network and the optimiser."""
    size: int = 256
    pixel_size: int = 128
    max_size: int = 512
    octaves: int = 6
    steps: int = 200
    warm_up: int = 2000
    base_lr: float = 2e-5
    beta1: float = 0.9
    beta2: float = 0.999
    epsilon: float = 1e-7
    #: explicit per-octave filter counts (what a layer tree built by hand carries); default = train.py's formulas
    down_filters: Optional[Tuple[int, ...]] = None
    up_filters: Optional[Tuple[int, ...]] = None
    #: train.py:34,43-45,82-83: Keras 'mixed_float16' policy + dynamic LossScaleOptimizer.  True: activations, their
    #: gradients and the weights' shadow copy are fp16 (tensor cores on fp16 operands, fp32 accumulation), the loss is
    #: multiplied by a dynamic scale before backward, steps with non-finite gradients are skipped.  False (the
    #: reference's default): this implementation's bf16 storage, which needs no loss scaling.
    mixed_precision: bool = False
    loss_scale_init: float = 2.0 ** 15      # Keras LossScaleOptimizer defaults
    loss_scale_growth: int = 2000
    #: train.py:29-32 (predict_x / predict_scaled_epsilon / prediction_weighting / ordinary_differential_equation) as
    #: ops.target_mode bits: what the loss compares (train.py:238-252) and what log_sample derives from a prediction
    target_mode: int = 0
    #: train.py:20,131-139: Block = block_depth x Conv2D(filters, 3, 1, 'same', relu); 0 (the reference's default) makes
    #: every Block an identity.  > 0 is run by block_engine.BlockUNetEngine
    block_depth: int = 0
    #: train.py:27,113-119: Residual concatenates [module(x), x]; False = module(x) alone (no skip connections)
    concat: bool = True
    #: train.py:26,106-112: Residual returns input + Dense(input_channels, use_bias=False)(module(input)) (and ignores
    #: `concat`)
    residual: bool = False
    #: filters of the innermost Block (train.py:179); None = min(pixel_size * 2**octaves, max_size)
    mid_filters: Optional[int] = None
    #: filters of the two outermost Blocks (train.py:192,194); None = pixel_size
    outer_filters: Optional[int] = None

    @property
    def fused_default(self) -> bool:
        """The wiring the tuned UNetEngine implements (the reference's defaults)."""
        return self.block_depth == 0 and self.concat and not self.residual

    def mid_c(self) -> int:  # train.py:179
        return self.mid_filters if self.mid_filters is not None else min(self.pixel_size * 2 ** self.octaves, self.max_size)

    def outer_c(self) -> int:  # train.py:192,194
        return self.outer_filters if self.outer_filters is not None else self.pixel_size

    def level_in(self, i: int) -> int:
        """Channels entering Residual level i (the tensor the skip connection carries)."""
        if i == 0:
            return self.outer_c() if self.block_depth else 3
        return self.down_c(i - 1)

    def res_out(self, i: int) -> int:
        """Channels leaving Residual level i (train.py:110-121)."""
        if self.residual:
            return self.level_in(i)
        return self.up_c(i) + (self.level_in(i) if self.concat else 0)

    def down_c(self, i: int) -> int:  # train.py:181
        if self.down_filters is not None:
            return self.down_filters[i]
        return min(self.pixel_size * 2 ** i, self.max_size)

    def up_c(self, i: int) -> int:  # train.py:188
        if self.up_filters is not None:
            return self.up_filters[i]
        return min(self.pixel_size * 2 ** i // 2, self.max_size)

    def up_in(self, i: int) -> int:
        """Input channels of UpShuffle i: behind a Block (block_depth > 0) the Block's filters, otherwise whatever the
        inner Residual (or, innermost, the DownShuffle) delivers."""
        if self.block_depth:
            return self.down_c(i)
        return self.down_c(i) if i == self.octaves - 1 else self.res_out(i + 1)

    def validate(self) -> None:
        n = self.octaves
        if n < 1 or self.size % (2 ** n) or (self.size >> n) < 4:
            raise ValueError(f"size {self.size} must be a multiple of 2^octaves with a bottleneck of at least 4x4")
        if self.size & (self.size - 1):
            raise ValueError("size must be a power of two")
        for i in range(n):
            if self.down_c(i) % 64 or self.up_c(i) % 64:
                raise ValueError("channel counts must be multiples of 64 (tensor-core tile granularity)")
        if self.block_depth < 0:
            raise ValueError("block_depth must be >= 0")
        if self.block_depth == 0 and self.down_c(0) % 128:
            raise ValueError("down0 filters must be a multiple of 128")
        dense_in = self.outer_c() if self.block_depth else self.up_c(0)
        if dense_in not in (64, 128):
            raise ValueError("the fused Dense(3)+MSE kernel supports 64 or 128 16-bit input channels")
        # weight gradients put one side's channels on the 128-row M axis of the tensor-core tile
        pairs = [(s[2], s[3]) for name, s in variable_specs(self) if name.endswith("kernel") and len(s) == 4 and s[2] != 3]
        pairs += [s for name, s in variable_specs(self) if name.startswith("res") and s[1] != 3]
        for a, b in pairs:
            if a % 64 or b % 64:
                raise ValueError("channel counts must be multiples of 64 (tensor-core tile granularity)")
            if a % 128 and b % 128:
                raise ValueError(f"a conv layer with {a} -> {b} channels has no side that is a multiple of 128 "
                                 "(needed by the weight-gradient kernel)")


def variable_specs(cfg: NetConfig) -> List[Tuple[str, Tuple[int, ...]]]:
    """Variable list of trainer.trainable_variables (train.py:175-204) in construction order, Keras
    layouts: Conv2D [k,k,Cin,Cout], Conv2DTranspose [4,4,Cout,Cin], Dense [Cin,3].  With block_depth > 0 every Block of
    the recursion contributes block_depth 3x3 convolutions: block_in (train.py:192), block_down{i} (:185), block_mid
    (:179), block_up{i} (:187), block_out (:194)."""
    specs: List[Tuple[str, Tuple[int, ...]]] = []
    n, d = cfg.octaves, cfg.block_depth

    def block(prefix: str, cin: int, filters: int) -> int:
        for k in range(d):
            specs.extend([(f"{prefix}/conv{k}/kernel", (3, 3, cin, filters)), (f"{prefix}/conv{k}/bias", (filters,))])
            cin = filters
        return cin

    cin = block("block_in", 3, cfg.outer_c())
    for i in range(n):
        specs += [(f"down{i}/kernel", (4, 4, cin, cfg.down_c(i))), (f"down{i}/bias", (cfg.down_c(i),))]
        cin = block(f"block_down{i}", cfg.down_c(i), cfg.down_c(i))
    cin = block("block_mid", cin, cfg.mid_c())
    for i in reversed(range(n)):
        c = block(f"block_up{i}", cin if i == n - 1 else cfg.res_out(i + 1), cfg.down_c(i))
        specs += [(f"up{i}/kernel", (4, 4, cfg.up_c(i), c)), (f"up{i}/bias", (cfg.up_c(i),))]
        if cfg.residual:  # train.py:107: Dense(input_shape[-1], use_bias=False)
            specs.append((f"res{i}/dense/kernel", (cfg.up_c(i), cfg.level_in(i))))
    c = block("block_out", cfg.res_out(0), cfg.outer_c())
    specs += [("dense/kernel", (c, 3)), ("dense/bias", (3,))]
    return specs


def glorot_uniform(shape, generator) -> torch.Tensor:
    """Keras glorot_uniform (train.py:134,149,162; Dense default): U(+-sqrt(6/(fan_in+fan_out)))."""
    rf = math.prod(shape[:-2]) if len(shape) > 2 else 1
    limit = math.sqrt(6.0 / (shape[-2] * rf + shape[-1] * rf))
    return (torch.rand(shape, generator=generator, dtype=torch.float32) * 2 - 1) * limit


def small_names(cfg: NetConfig) -> List[str]:
    """Variables whose gradients are accumulated with atomics by HBM-bound kernels (down0's kernel, every bias, the
    Dense layer): they live together at the head of the flat buffers so that one memset zeroes their gradients."""
    specs = variable_specs(cfg)
    image_kernels = [n for n, s in specs if n.endswith("kernel") and len(s) == 4 and s[2] == 3]  # CUDA-core convs on the image
    names = image_kernels + [n for n, _ in specs if n.endswith("bias") and not n.startswith("dense")]
    if cfg.residual and cfg.block_depth == 0:
        names.append("res0/dense/kernel")  # [U,3]: updated with the Dense(3) it is folded into (gct2_res0_compose)
    return names + ["dense/kernel", "dense/bias"]


def param_offsets(cfg: NetConfig) -> Tuple[Dict[str, Tuple[int, int]], int]:
    """name -> (offset, count) in the flat buffers, and the total element count.

    Internal layout (the Keras variable order of variable_specs is the *naming* convention, not the address order):
      [ small region: down0/kernel, biases in Keras order, dense/kernel, dense/bias, padded to 64 elements ]
      [ down1/kernel .. down{n-1}/kernel, up{n-1}/kernel .. up0/kernel ]
    Backward completes the tensor-core kernels' gradients from the tail (up0) to down1 and the small region last, so
    data-parallel buckets are contiguous tail-to-head ranges.  Every tensor-core kernel starts on a 128-byte boundary
    (TMA needs 16)."""
    shapes = dict(variable_specs(cfg))
    offsets: Dict[str, Tuple[int, int]] = {}
    off = 0
    for name in small_names(cfg):
        cnt = math.prod(shapes[name])
        offsets[name] = (off, cnt)
        off += cnt
    off = (off + 63) // 64 * 64
    for name, shape in variable_specs(cfg):
        if name in offsets:
            continue
        cnt = math.prod(shape)
        offsets[name] = (off, cnt)
        off += cnt
    return offsets, off


def small_region(cfg: NetConfig) -> int:
    """Element count of the head region described in param_offsets (including its padding)."""
    offsets, _ = param_offsets(cfg)
    return min(off for name, (off, _) in offsets.items() if name not in small_names(cfg))


def grad_buckets(cfg: NetConfig, bucket_bytes: int) -> List[Tuple[int, int, str]]:
    """Data-parallel gradient buckets: contiguous [start, end) ranges of the flat fp32 gradient buffer, listed tail
    to head (the order backward completes them: up0..up{n-1}, down{n-1}..down1, then the small region), each closed by
    the variable whose gradient completes it.  Together they tile [0, P) exactly once."""
    offsets, total = param_offsets(cfg)
    order = [f"up{i}/kernel" for i in range(cfg.octaves)] + [f"down{i}/kernel" for i in reversed(range(1, cfg.octaves))]
    buckets: List[Tuple[int, int, str]] = []
    cur_end = total
    for name in order:
        start = offsets[name][0]
        if (cur_end - start) * 4 >= bucket_bytes:
            buckets.append((start, cur_end, name))
            cur_end = start
    buckets.append((0, cur_end, "down0/kernel"))  # remaining kernels (if any) + the small region, ready last
    return buckets


def optimizer_shard(start: int, end: int, small: int, world: int, rank: int) -> Optional[Tuple[int, int, int, int]]:
    """Sharded optimiser: the part [lo, end) of a gradient bucket above the replicated head region [0, small) is cut
    into `world` equal slices of whole float4 vectors; returns (lo, end, own_lo, own_hi) of `rank`, or None when the
    bucket has nothing above the head region or cannot be cut evenly (then it stays replicated: all-reduce + Adam on
    every rank)."""
    lo = max(start, small)
    if end <= lo or (end - lo) % (4 * world):
        return None
    chunk = (end - lo) // world
    return lo, end, lo + rank * chunk, lo + (rank + 1) * chunk


def shard_batch(global_batch: int, world: int, rank: int) -> Tuple[int, int]:
    """[lo, hi) sample range of one rank: the batch shards evenly, samples are independent (no cross-sample op but
    the loss mean, train.py:272)."""
    if global_batch % world:
        raise ValueError(f"global batch {global_batch} is not divisible by {world} ranks")
    per = global_batch // world
    return rank * per, (rank + 1) * per


def plan_table_key(cfg: NetConfig, batch: int, sms: int) -> str:
    """Which tuned table applies: the network's shape, the per-GPU batch, the storage policy and the SM count."""
    ch = ",".join(f"{cfg.down_c(i)}:{cfg.up_c(i)}" for i in range(cfg.octaves))
    return f"size{cfg.size}_oct{cfg.octaves}_ch{ch}_b{batch}_{'f16' if cfg.mixed_precision else 'bf16'}_sm{sms}"


def load_tuned_plans(cfg: NetConfig, batch: int, sms: int) -> Dict[str, Tuple[int, int]]:
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tuned_plans.json")
    if not os.path.exists(path):
        return {}
    with open(path) as f:
        tables = json.load(f)
    table = tables.get(plan_table_key(cfg, batch, sms), {}).get("plans", {})
    return {k: (int(v[0]), int(v[1])) for k, v in table.items()}


class DataParallel:
    """Data-parallel context: the batch shards over ranks, gradient buckets are summed with NCCL."""

    def __init__(self, group=None, bucket_bytes: Optional[int] = None, shard_optimizer: bool = True,
                 grad_dtype: Optional[str] = None, nccl_ctas: Optional[int] = None, transport: Optional[str] = None):
        import os
        if bucket_bytes is None:
            bucket_bytes = int(float(os.environ.get("GCT2_DP_BUCKET_MB", "48")) * (1 << 20))
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.bucket_bytes = bucket_bytes
        #: "bf16": the gradient reduce-scatter of the sharded optimiser runs on a bf16 copy of each bucket (half the
        #: NVLink bytes; the fp32 gradients already carry bf16 operand noise) and Keras-Adam reads the summed bf16 values;
        #: "fp32": the fp32 gradients themselves are reduced.  The replicated paths always reduce fp32.
        self.grad_dtype = grad_dtype or os.environ.get("GCT2_DP_GRAD", "bf16")
        if self.grad_dtype not in ("bf16", "fp32"):
            raise ValueError("grad_dtype must be 'bf16' or 'fp32'")
        #: SMs left to the NCCL kernels that run beside backward (NCCL_MAX_CTAS caps them): the tensor-core launches of
        #: backward keep to the other SMs (gct2_set_sm_budget) instead of queueing a second wave behind a collective
        self.nccl_ctas = nccl_ctas if nccl_ctas is not None else int(os.environ.get("NCCL_MAX_CTAS", "0") or 0)
        #: how the remaining SMs are shared during backward: "split" = the dgrad chain (main stream) and the wgrad chain
        #: (side stream) get half each, so that together with the NCCL kernels nothing ever queues behind anything;
        #: "shared" = each chain may use all of them (the chains then take turns and a collective waits for a gap)
        self.sm_sharing = os.environ.get("GCT2_DP_SM", "shared")
        #: how gradients and weights cross the GPUs: "p2p" = ONE kernel per bucket does the gradient sum, Keras-Adam and the
        #: weight broadcast with its own loads and stores over NVLink peer memory (symmetric-memory buffers; with NVLS the
        #: sum is a multimem.ld_reduce inside the switch and the broadcast a multimem.st) -- no NCCL kernel, no SMs set
        #: aside for a collective; "nccl" = reduce-scatter / all-gather calls.  "p2p" falls back to "nccl" when the
        #: buffers cannot be mapped.
        self.transport = transport or os.environ.get("GCT2_DP_TRANSPORT", "p2p")
        #: NVLS multicast addresses (multimem.ld_reduce / multimem.st) for the fused kernel's sum and broadcast.  Measured on
        #: 2 / 4 / 8 B200s at 1 image per GPU (profiles/r2_scaling_transports.jsonl): 2964 / 6238 / 12604 images/s with
        #: them, 3090 / 6215 / 11521 with plain peer loads and stores -- the switch-side reduction pays from four GPUs on,
        #: between two GPUs a peer load is cheaper than a round trip through the switch's reduction unit
        mc = os.environ.get("GCT2_DP_MULTICAST", "")
        self.multicast = (self.world > 2) if mc == "" else (mc != "0")
        #: measurement aid (bench.py's communication breakdown): when True the step is enqueued WITHOUT its collectives
        #: (wrong numbers, right compute time); read when a step is enqueued / captured
        self.dry_run = False
        #: SURVEY.md 8(e) "optimisation": per bucket, reduce-scatter the fp32 gradients, run Keras-Adam on this rank's
        #: 1/N slice only, all-gather the bf16 weights the tensor-core kernels read.  0.75x the bytes of an fp32
        #: all-reduce on the wire and 1/N of the optimiser's HBM traffic per GPU.  The fp32 masters of the other ranks'
        #: slices go stale until UNetEngine.gather_master_weights() (called by weights()).
        self.shard_optimizer = shard_optimizer


class _NoWork:
    def wait(self):
        pass


class _Collectives:
    """The three collectives of the data-parallel step; DataParallel.dry_run turns them into no-ops."""

    def __init__(self, dp: DataParallel):
        self.dp = dp

    def reduce_scatter(self, out, inp):
        if self.dp.dry_run:
            return _NoWork()
        return self.dp.dist.reduce_scatter_tensor(out, inp, group=self.dp.group, async_op=True)

    def all_gather(self, out, inp):
        if self.dp.dry_run:
            return _NoWork()
        return self.dp.dist.all_gather_into_tensor(out, inp, group=self.dp.group, async_op=True)

    def all_reduce(self, t):
        if self.dp.dry_run:
            return _NoWork()
        return self.dp.dist.all_reduce(t, group=self.dp.group, async_op=True)


class UNetEngine:
    def __init__(self, cfg: NetConfig, batch: int, device=None, dp: Optional[DataParallel] = None,
                 use_graph: bool = False, share_params_with: Optional["UNetEngine"] = None):
        cfg.validate()
        if type(self) is UNetEngine and not cfg.fused_default:
            raise NotImplementedError("UNetEngine implements the reference's default wiring (block_depth = 0, concat = True); "
                                      "use engine.make_engine / block_engine.BlockUNetEngine for the dormant switches")
        self.cfg = cfg
        self.B = batch
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.dp = dp
        self.half = torch.float16 if cfg.mixed_precision else torch.bfloat16  # the 16-bit storage format (ops.HALF)
        if cfg.mixed_precision and dp is not None and dp.world > 1:
            raise NotImplementedError("mixed_precision (fp16 + dynamic loss scaling) is implemented for single-GPU steps")
        self.use_graph = use_graph
        self.rng_seed = 0x5DEECE66D + 7919 * (dp.rank if dp else 0)  # every rank draws its own noise
        import os
        mode = os.environ.get("GCT2_OVERLAP", "all")  # test hook: none | wgrad | adam | all
        self.overlap_wgrad = mode in ("all", "wgrad")
        self.overlap_adam = mode in ("all", "adam")
        #: GCT2_WEIGHTS_STABLE for the step's tensor-core launches: the bf16 kernels are only ever written on the side
        #: streams (Keras-Adam, the data-parallel all-gather), which rejoin the main stream through events -- a full
        #: dependency -- so no launch that writes them can still be running when a conv launch of the main stream starts
        #: its prologue.  With the optimiser on the main stream (test hook GCT2_OVERLAP=none|wgrad) an Adam launch could
        #: still be draining under programmatic dependent launch, and the flag stays off.
        #: Off under mixed precision as well: there the whole update is one launch on the main stream after backward.
        self.weights_stable = (self.overlap_adam and not cfg.mixed_precision
                               and os.environ.get("GCT2_WEIGHTS_EARLY", "1") != "0")
        #: Keras-Adam beside backward on disjoint SMs (single GPU): the optimiser is HBM-bound and draws ~98 GB/s per SM,
        #: the tensor-core launches of backward need no HBM bandwidth to speak of -- so the first `adam_wide_buckets`
        #: gradient buckets (backward order: up0 .. up{n-1}, down{n-1} ..) are updated by `adam_sms` SM-exclusive CTAs
        #: while the conv launches keep to the other SMs (gct2_set_sm_budget); the remaining buckets, complete only when
        #: backward is over, use the whole chip.  0 = the optimiser always uses the whole chip (takes turns with the convs).
        self.adam_sms = int(os.environ.get("GCT2_ADAM_SMS", "0"))
        self.adam_wide_buckets = int(os.environ.get("GCT2_ADAM_WIDE", "5"))
        self._side = torch.cuda.Stream(device=self.device)
        self._side_adam = torch.cuda.Stream(device=self.device)
        self._copy_stream = torch.cuda.Stream(device=self.device)   # host batches in flight while the previous step runs
        self._staging: Dict[tuple, dict] = {}
        self._graph = None
        self._graph_launches = 0
        self._masters_stale = False
        self._sharded_ranges = set()  # (lo, end) ranges whose fp32 masters are updated on their owner rank only
        dev, n, S, B = self.device, cfg.octaves, cfg.size, batch
        from . import _lib
        lib = _lib.init(self.device.index or 0)
        self._lib = lib
        #: per-launch plan overrides {"<layer>/<pass>": (BN, splits)}; tuned_plans.json holds tables measured in the step
        #: (tools/tune_plans.py) for single-GPU configurations, GCT2_TUNED=0 ignores them
        self.plans: Dict[str, Tuple[int, int]] = {}
        if os.environ.get("GCT2_TUNED", "1") != "0" and not (dp is not None and dp.world > 1):
            self.plans = load_tuned_plans(cfg, batch, _lib.load().gct2_num_sms())
        if dp is not None and dp.world > 1:
            # The L2 form of the in-launch split-K finish makes the CTAs of a conv launch wait for one another, which is
            # only safe while every other kernel on the GPU terminates on its own.  An NCCL kernel waits for its peer GPU
            # and must itself be fully resident: a half-placed conv launch and a half-placed NCCL kernel can then hold
            # all 148 SMs between them forever (seen as a hang of the captured 2-GPU step in round 1).  Data-parallel
            # steps therefore finish split-K inside a thread-block cluster (co-scheduled by the hardware, so waiting
            # inside it is always safe) or with the separate finishing kernel.
            lib.gct2_debug_set(26, 1)
            lib.gct2_debug_set(25, 0)

        # ---- parameters, flat in Keras order
        self.specs = variable_specs(cfg)
        self.offsets, self.P = param_offsets(cfg)
        self.small = small_region(cfg)
        if self.P % 4:
            raise ValueError("flat parameter buffer must be a multiple of 4 elements")
        f32 = dict(dtype=torch.float32, device=dev)
        if share_params_with is not None:
            # a second batch size over the same variables (e.g. the reference's batch-6 sampling, train.py:432-450)
            o = share_params_with
            if o.specs != self.specs:
                raise ValueError("engines sharing parameters must have the same variable list")
            if o.cfg.mixed_precision != cfg.mixed_precision:
                raise ValueError("engines sharing parameters must use the same precision policy")
            self.w, self.m, self.v, self.g, self.w16 = o.w, o.m, o.v, o.g, o.w16
            self.iterations, self.hyper, self.ls, self._it_scratch = o.iterations, o.hyper, o.ls, o._it_scratch
        else:
            self.w = torch.zeros(self.P, **f32)
            self.m = torch.zeros(self.P, **f32)
            self.v = torch.zeros(self.P, **f32)
            self.g = torch.zeros(self.P, **f32)
            self.w16 = torch.zeros(self.P, dtype=self.half, device=dev)
            self.iterations = torch.zeros(1, dtype=torch.int64, device=dev)
            self.hyper = torch.zeros(2, **f32)
            # dynamic loss scaling state {scale, good steps, finite flag, 1/scale} (mixed precision only)
            self.ls = torch.tensor([cfg.loss_scale_init, 0.0, 1.0, 1.0 / cfg.loss_scale_init], **f32)
            self._it_scratch = torch.zeros(1, dtype=torch.int64, device=dev)
        # bf16 copy of the gradient buckets for the data-parallel reduce-scatter (DataParallel.grad_dtype)
        self.g16 = (torch.zeros(self.P, dtype=torch.bfloat16, device=dev)
                    if (dp is not None and dp.world > 1 and dp.shard_optimizer and dp.grad_dtype == "bf16") else None)
        self._p2p = None
        if (dp is not None and dp.world > 1 and dp.shard_optimizer and dp.transport == "p2p" and share_params_with is None
                and dp.world <= 8):
            self._p2p = self._map_peer_buffers()
        # ---- the step's inputs and outputs (activations and their gradients: _alloc_activations)
        self.x = torch.zeros(B, S, S, 3, **f32)
        self.x_u8 = torch.zeros(B, S, S, 3, dtype=torch.uint8, device=dev)   # decode_file's bytes (train.py:285-293)
        self.flip = torch.zeros(B, dtype=torch.uint8, device=dev)           # per-image left-right flip flags
        self.eps = torch.zeros(B, S, S, 3, **f32)
        self.t_int = torch.ones(B, dtype=torch.int32, device=dev)
        self.noised = torch.zeros(B, S, S, 3, **f32)
        self.pred = torch.zeros(B, S, S, 3, **f32)
        self.loss = torch.zeros(1, **f32)
        self.global_batch = B * (dp.world if dp else 1)
        self._alloc_activations()
        #: With the peer-memory transport no NCCL kernel runs beside a conv launch (the head region's all-reduce and the
        #: loss all-reduce start only after the last dgrad; what does run beside backward -- the gradient cast, the
        #: signal-pad barrier's single small CTA, the fused exchange kernel -- either terminates on its own or leaves room
        #: for a conv CTA on its SM), so the L2 rendezvous form of the split-K finish and the single-GPU plan table are
        #: usable again.  GCT2_DP_L2_FINISH=1 turns that on (an A/B switch until it has been measured on 8 GPUs).
        all_fused = self._p2p is not None and all(
            (cut := optimizer_shard(st, en, self.small, dp.world, dp.rank)) is not None
            and (cut[3] - cut[2]) % 8 == 0 and cut[2] % 8 == 0 for st, en, _ in self._buckets if en > self.small)
        self.dp_l2_finish = all_fused and os.environ.get("GCT2_DP_L2_FINISH", "0") == "1"
        if self.dp_l2_finish:
            lib.gct2_debug_set(26, 0)
            lib.gct2_debug_set(25, 1)
            if os.environ.get("GCT2_TUNED", "1") != "0":
                self.plans = load_tuned_plans(cfg, batch, _lib.load().gct2_num_sms())


    def _alloc_activations(self) -> None:
        """Activation / gradient buffers of the default wiring (see the module docstring) and everything sized by them."""
        import os
        cfg, dev, dp = self.cfg, self.device, self.dp
        n, S, B = cfg.octaves, cfg.size, self.B
        bf = dict(dtype=self.half, device=dev)
        self.cat: Dict[int, torch.Tensor] = {}
        self.gcat: Dict[int, torch.Tensor] = {}
        for j in range(1, n):
            shape = (B, S >> j, S >> j, cfg.up_c(j) + cfg.down_c(j - 1))
            self.cat[j] = torch.zeros(shape, **bf)
            self.gcat[j] = torch.zeros(shape, **bf)
        self.bot = torch.zeros(B, S >> n, S >> n, cfg.down_c(n - 1), **bf)
        self.gbot = torch.zeros_like(self.bot)
        self.u0 = torch.zeros(B, S, S, cfg.up_c(0), **bf)
        self.gu0 = torch.zeros_like(self.u0)
        biggest = max([self.u0.numel(), self.bot.numel()] + [c.numel() for c in self.cat.values()])
        # split-K scratch: one for the main stream (fprop / dgrad partial outputs), one for the side stream (wgrad)
        self.ws = ops.Workspace(max(4 * 4 * biggest, 64 << 20), dev)
        self.ws_w = ops.Workspace(64 << 20, dev)
        # gradient buckets: all-reduce granularity (data parallel) and the granularity at which Adam chases backward
        local_bucket = int(float(os.environ.get("GCT2_BUCKET_MB", "8")) * (1 << 20))  # Adam granularity (8 MB: per layer)
        self._buckets = grad_buckets(cfg, dp.bucket_bytes if dp else local_bucket)
        layers = [f"down{i}" for i in range(n)] + [f"up{i}" for i in range(n)]
        self._bias_plan = ops.BiasGradPlan([self.gdown_out(i) for i in range(n)] + [self.gup_out(i) for i in range(n)],
                                           [self.view(self.g, f"{l}/bias") for l in layers])

    def _map_peer_buffers(self):
        """Re-homes the bf16 gradient copy and the 16-bit weight shadow in symmetric memory (every rank's buffer mapped
        into every process, plus NVLS multicast addresses where the fabric offers them).  Returns None -- and the step
        falls back to NCCL -- when that is not possible."""
        dp = self.dp
        try:
            import torch.distributed._symmetric_memory as symm_mem
            group = dp.group if dp.group is not None else dp.dist.group.WORLD
            g16 = symm_mem.empty(self.P, dtype=torch.bfloat16, device=self.device)
            w16 = symm_mem.empty(self.P, dtype=self.half, device=self.device)
            hg = symm_mem.rendezvous(g16, group)
            hw = symm_mem.rendezvous(w16, group)
            g16.zero_()
            w16.copy_(self.w16)
            # fp32 staging of the replicated head region's gradients + the loss (summed by peer loads, no NCCL call)
            head = symm_mem.empty(self.small + 4, dtype=torch.float32, device=self.device)
            hh = symm_mem.rendezvous(head, group)
            head.zero_()
            g_mc = int(getattr(hg, "multicast_ptr", 0) or 0) if dp.multicast else 0
            w_mc = int(getattr(hw, "multicast_ptr", 0) or 0) if dp.multicast else 0
            if not (g_mc and w_mc):
                g_mc = w_mc = 0
            self.g16, self.w16 = g16, w16
            torch.cuda.synchronize(self.device)
            hw.barrier(channel=0)
            return dict(hg=hg, hw=hw, g_ptrs=[int(p) for p in hg.buffer_ptrs], w_ptrs=[int(p) for p in hw.buffer_ptrs],
                        g_mc=g_mc, w_mc=w_mc, head=head, hh=hh, head_ptrs=[int(p) for p in hh.buffer_ptrs],
                        head_p2p=os.environ.get("GCT2_DP_HEAD", "p2p") == "p2p")
        except Exception as exc:  # noqa: BLE001 -- any failure here means "no peer mapping": use the collectives
            import warnings
            warnings.warn(f"gct2: peer-memory transport unavailable ({type(exc).__name__}: {exc}); using NCCL collectives")
            return None

    # ------------------------------------------------------------------------------------------ parameter access
    def view(self, buf: torch.Tensor, name: str) -> torch.Tensor:
        off, cnt = self.offsets[name]
        return buf[off:off + cnt].view(dict(self.specs)[name])

    def load_weights(self, weights: Dict[str, torch.Tensor], reset_optimizer: bool = True) -> None:
        for name, shape in self.specs:
            t = weights[name]
            if tuple(t.shape) != tuple(shape):
                raise ValueError(f"{name}: expected shape {shape}, got {tuple(t.shape)}")
            self.view(self.w, name).copy_(t.to(torch.float32))
        ops.cast_bf16(self.w, self.w16)
        if reset_optimizer:
            self.m.zero_()
            self.v.zero_()
            self.iterations.zero_()
            self.ls.copy_(torch.tensor([self.cfg.loss_scale_init, 0.0, 1.0, 1.0 / self.cfg.loss_scale_init]))
        self._graph = None
        # the bf16 kernels must be at rest before a step may fetch them ahead of its dependencies (weights_stable)
        torch.cuda.current_stream().synchronize()

    def init_glorot(self, seed: int = 0) -> None:
        gen = torch.Generator().manual_seed(seed)
        self.load_weights({name: torch.zeros(shape) if name.endswith("bias") else glorot_uniform(shape, gen)
                           for name, shape in self.specs})

    def gather_master_weights(self) -> None:
        """Sharded optimiser (DataParallel.shard_optimizer): brings the fp32 masters of the other ranks' slices up to
        date (all ranks must call it).  The training step itself never needs them -- it reads the bf16 copies."""
        dp = self.dp
        if not dp or dp.world == 1 or not self._sharded_ranges:
            self._masters_stale = False
            return
        torch.cuda.synchronize(self.device)
        for lo, end in sorted(self._sharded_ranges):
            chunk = (end - lo) // dp.world
            own = lo + dp.rank * chunk
            dp.dist.all_gather_into_tensor(self.w[lo:end], self.w[own:own + chunk], group=dp.group)
        torch.cuda.synchronize(self.device)
        self._masters_stale = False

    def weights(self) -> Dict[str, torch.Tensor]:
        if self._masters_stale:
            raise RuntimeError("sharded optimiser: call gather_master_weights() on every rank before reading the fp32 "
                               "weights (each rank has updated the masters of its own slice only)")
        return {name: self.view(self.w, name).detach().clone() for name, _ in self.specs}

    def grads(self) -> Dict[str, torch.Tensor]:
        """The gradients of the last backward pass (mixed precision: with the loss scale removed)."""
        unscale = float(self.ls[3]) if self.cfg.mixed_precision else 1.0
        return {name: self.view(self.g, name).detach().clone() * unscale for name, _ in self.specs}

    # ------------------------------------------------------------------------------------------ buffer wiring
    def down_in(self, i: int) -> torch.Tensor:
        """Input of DownShuffle i (i >= 1): the skip slice of cat[i] (= down_{i-1}'s output)."""
        return self.cat[i][..., self.cfg.up_c(i):]

    def down_out(self, i: int) -> torch.Tensor:
        return self.bot if i == self.cfg.octaves - 1 else self.cat[i + 1][..., self.cfg.up_c(i + 1):]

    def gdown_out(self, i: int) -> torch.Tensor:
        return self.gbot if i == self.cfg.octaves - 1 else self.gcat[i + 1][..., self.cfg.up_c(i + 1):]

    def up_in_buf(self, i: int) -> torch.Tensor:
        return self.bot if i == self.cfg.octaves - 1 else self.cat[i + 1]

    def gup_in_buf(self, i: int) -> torch.Tensor:
        return self.gbot if i == self.cfg.octaves - 1 else self.gcat[i + 1]

    def up_out(self, i: int) -> torch.Tensor:
        return self.u0 if i == 0 else self.cat[i][..., :self.cfg.up_c(i)]

    def gup_out(self, i: int) -> torch.Tensor:
        return self.gu0 if i == 0 else self.gcat[i][..., :self.cfg.up_c(i)]

    # ------------------------------------------------------------------------------------------ tensor-core passes
    def plan_keys(self) -> List[str]:
        """The tensor-core launches of one step in step order: "<layer>/<fprop|dgrad|wgrad>" (down0 runs on CUDA cores)."""
        n = self.cfg.octaves
        keys = [f"down{i}/fprop" for i in range(1, n)] + [f"up{i}/fprop" for i in reversed(range(n))]
        for i in range(n):
            keys += [f"up{i}/wgrad", f"up{i}/dgrad"]
        for i in reversed(range(1, n)):
            keys += [f"down{i}/wgrad", f"down{i}/dgrad"]
        return keys

    def _conv_pass(self, key: str) -> None:
        """One tensor-core launch of the step on the current stream.  `self.plans[key] = (BN, splits)` (tile width and
        split-K factor; 0 = the library's cost model decides) overrides the plan: the table is measured *in the step*
        by tools/tune_plans.py, where launches compete for SMs with the two other chains -- something a per-launch cost
        model cannot see."""
        cfg, n = self.cfg, self.cfg.octaves
        layer, kind = key.split("/")
        i = int(layer[4:] if layer.startswith("down") else layer[2:])
        bn, splits = self.plans.get(key, (0, 0))
        if bn or splits:
            self._lib.gct2_debug_set(3, int(bn))
            self._lib.gct2_debug_set(4, int(splits))
        try:
            if layer.startswith("down"):
                if kind == "fprop":
                    ops.conv4s2_fprop(self.down_in(i), self.view(self.w16, f"down{i}/kernel"),
                                      self.view(self.w, f"down{i}/bias"), self.down_out(i), self.ws, self.weights_stable)
                elif kind == "wgrad":
                    ops.conv4s2_wgrad(self.down_in(i), self.gdown_out(i), self.view(self.g, f"down{i}/kernel"), self.ws_w)
                else:
                    # total gradient of down_{i-1}'s output = skip-path part (stored raw by up_{i-1}'s dgrad) + this
                    ops.conv4s2_dgrad(self.gdown_out(i), self.view(self.w16, f"down{i}/kernel"),
                                      self.gcat[i][..., cfg.up_c(i):], self.down_in(i), True, self.ws, self.weights_stable)
            else:
                if kind == "fprop":
                    ops.convT4s2_fprop(self.up_in_buf(i), self.view(self.w16, f"up{i}/kernel"),
                                       self.view(self.w, f"up{i}/bias"), self.up_out(i), self.ws, self.weights_stable)
                elif kind == "wgrad":
                    ops.convT4s2_wgrad(self.up_in_buf(i), self.gup_out(i), self.view(self.g, f"up{i}/kernel"), self.ws_w)
                else:
                    mask = cfg.down_c(i) if i == n - 1 else cfg.up_c(i + 1)
                    ops.convT4s2_dgrad(self.gup_out(i), self.view(self.w16, f"up{i}/kernel"), self.gup_in_buf(i),
                                       self.up_in_buf(i), mask, self.ws, self.weights_stable)
        finally:
            if bn or splits:
                self._lib.gct2_debug_set(3, 0)
                self._lib.gct2_debug_set(4, 0)

    # ------------------------------------------------------------------------------------------ forward / backward
    def _forward(self, want_pred: bool, backward: bool, inv_n: float) -> None:
        cfg, n = self.cfg, self.cfg.octaves
        ops.conv4s2_c3_fprop(self.noised, self.view(self.w, "down0/kernel"), self.view(self.w, "down0/bias"),
                             self.down_out(0))
        for i in range(1, n):
            self._conv_pass(f"down{i}/fprop")
        for i in reversed(range(n)):
            self._conv_pass(f"up{i}/fprop")
        ops.dense_mse(self.u0, self.noised, self.x, self.view(self.w, "dense/kernel"), self.view(self.w, "dense/bias"),
                      self.loss, inv_n, pred=self.pred if want_pred else None,
                      du0=self.gu0 if backward else None, dwd=self.view(self.g, "dense/kernel") if backward else None,
                      dbd=self.view(self.g, "dense/bias") if backward else None, accumulate=True,
                      loss_scale=self.ls if (backward and self.cfg.mixed_precision) else None,
                      eps=self.eps, t_int=self.t_int, mode=cfg.target_mode, steps=cfg.steps)

    def _backward(self, apply_adam: bool, inc_iterations: bool = False) -> None:
        """Three chains that share the GPU (at batch 1 no layer fills 148 SMs on its own):
          main stream  the dgrad chain (the only true dependency chain of backward), then down0's wgrad + bias grads;
          side_w       the 12 tensor-core weight gradients, each released as soon as its dz exists, and -- data
                       parallel -- the NCCL all-reduce of every gradient bucket the moment its last wgrad is enqueued;
          side_a       Keras-Adam on each bucket as soon as the bucket is complete AND the dgrad that still reads the
                       bucket's bf16 weights has been enqueued (HBM-bound, overlaps the L2/tensor-bound dgrad chain).
        Everything rejoins the main stream before the step ends, so the whole step captures into one CUDA graph."""
        cfg, n = self.cfg, self.cfg.octaves
        main = torch.cuda.current_stream()
        sw = self._side if self.overlap_wgrad else main
        sa = self._side_adam if self.overlap_adam else main
        dp = self.dp if (self.dp and self.dp.world > 1) else None
        coll = _Collectives(dp) if dp else None
        pending = []
        from . import _lib
        num_sms = _lib.load().gct2_num_sms()
        # [wide buckets left, conv SM budget active]
        side = [self.adam_wide_buckets if (apply_adam and dp is None and sa is not main and 0 < self.adam_sms < num_sms)
                else 0, False]
        chain_caps = False
        p2p_used = [False]
        head_done = [False]  # head region + loss already summed over the ranks by peer loads
        if dp is not None and 0 < dp.nccl_ctas < num_sms and self._p2p is None:
            ops.set_sm_budget(num_sms - dp.nccl_ctas)  # the collectives of backward keep their SMs
            side[1] = True
            if dp.sm_sharing == "split" and sw is not main:
                half = (num_sms - dp.nccl_ctas) // 2
                lib_ = _lib.load()
                lib_.gct2_debug_set(9, half)   # wgrad launches (side stream)
                lib_.gct2_debug_set(10, half)  # dgrad launches (main stream)
                chain_caps = True

        def on_side(fn):
            if sw is main:
                fn()
                return
            sw.wait_stream(main)  # dz of this layer is complete on the main stream
            with torch.cuda.stream(sw):
                fn()

        def bucket_done(trigger: str):
            """Called once the dgrad reading `trigger`'s weights is on the main stream."""
            for start, end, name in self._buckets:
                if name != trigger:
                    continue
                work = None
                if dp and apply_adam and dp.shard_optimizer:
                    # sharded optimiser: everything above the small head region of this bucket
                    cut = optimizer_shard(start, end, self.small, dp.world, dp.rank)
                    if cut is not None:
                        lo, _, own, own_hi = cut
                        chunk = own_hi - own
                        if sw is not main and trigger == "down0/kernel":
                            sw.wait_stream(main)
                        if self._p2p is not None and chunk % 8 == 0 and own % 8 == 0:
                            # fused: cast -> barrier -> ONE kernel (gradient sum over NVLink, Keras-Adam on my slice, weight
                            # broadcast into every rank's shadow); see gct2_adam_apply_p2p
                            pp = self._p2p
                            # the last bucket carries the replicated head region below it: its gradients and the loss go
                            # through the same barrier and are summed by peer loads right after the fused kernel
                            head_now = start < self.small and pp["head_p2p"]
                            with torch.cuda.stream(sw):
                                ops.cast_bf16(self.g[lo:end], self.g16[lo:end])
                                if head_now and not dp.dry_run:
                                    pp["head"][:self.small].copy_(self.g[:self.small], non_blocking=True)
                                    pp["head"][self.small:self.small + 1].copy_(self.loss, non_blocking=True)
                            if sa is not main:
                                sa.wait_stream(main)   # ... including the dgrad that still reads this bucket's weights
                                if sw is not main:
                                    sa.wait_stream(sw)
                            elif sw is not main:
                                main.wait_stream(sw)
                            with torch.cuda.stream(sa):
                                if dp.dry_run:
                                    ops.adam_apply(self.w[own:own + chunk], self.m[own:own + chunk], self.v[own:own + chunk],
                                                   self.g16[own:own + chunk], self.w16[own:own + chunk], self.hyper,
                                                   cfg.beta1, cfg.beta2, cfg.epsilon, 1.0)
                                else:
                                    pp["hg"].barrier(channel=0)   # every rank's gradients of this bucket are in place
                                    ops.adam_apply_p2p(self.w[own:own + chunk], self.m[own:own + chunk],
                                                       self.v[own:own + chunk], pp["g_ptrs"], pp["w_ptrs"], dp.world, own,
                                                       self.hyper, cfg.beta1, cfg.beta2, cfg.epsilon, 1.0, True,
                                                       pp["g_mc"], pp["w_mc"])
                            self._sharded_ranges.add((lo, end))
                            p2p_used[0] = True
                            if start >= self.small:
                                continue
                            if head_now:
                                with torch.cuda.stream(sa):
                                    if not dp.dry_run:
                                        ops.sum_peers_f32(pp["head_ptrs"], dp.world, self.g[:self.small], self.loss)
                                    ops.adam_apply(self.w[:self.small], self.m[:self.small], self.v[:self.small],
                                                   self.g[:self.small], self.w16[:self.small], self.hyper, cfg.beta1,
                                                   cfg.beta2, cfg.epsilon, 1.0,
                                                   iterations_inc=self.iterations if inc_iterations else None)
                                head_done[0] = True
                                continue
                            end = self.small  # the head region below stays replicated
                            cut = None
                    if cut is not None:
                        gsrc = self.g
                        with torch.cuda.stream(sw):
                            if self.g16 is not None:
                                ops.cast_bf16(self.g[lo:end], self.g16[lo:end])
                                gsrc = self.g16
                            # in place: my slice of the bucket receives the sum of everybody's slice
                            rs = coll.reduce_scatter(gsrc[own:own + chunk], gsrc[lo:end])
                        if sa is not main:
                            sa.wait_stream(main)
                            if sw is not main:
                                sa.wait_stream(sw)
                        elif sw is not main:
                            main.wait_stream(sw)
                        with torch.cuda.stream(sa):
                            rs.wait()
                            ops.adam_apply(self.w[own:own + chunk], self.m[own:own + chunk], self.v[own:own + chunk],
                                           gsrc[own:own + chunk], self.w16[own:own + chunk], self.hyper, cfg.beta1,
                                           cfg.beta2, cfg.epsilon, 1.0)
                            # every rank's freshly written bf16 slice to everybody (in place)
                            pending.append(coll.all_gather(self.w16[lo:end], self.w16[own:own + chunk]))
                        self._sharded_ranges.add((lo, end))
                        if start >= self.small:
                            continue
                        end = self.small  # the head region below stays replicated
                if dp:
                    if sw is not main and trigger == "down0/kernel":
                        sw.wait_stream(main)  # the small region is produced on the main stream
                    with torch.cuda.stream(sw):
                        work = coll.all_reduce(self.g[start:end])
                if not apply_adam:
                    if work is not None:
                        pending.append(work)
                    continue
                if sa is not main:
                    sa.wait_stream(main)
                    if sw is not main:
                        sa.wait_stream(sw)
                elif sw is not main:
                    main.wait_stream(sw)
                wide = side[0] > 0 and start != 0
                with torch.cuda.stream(sa):
                    if work is not None:
                        work.wait()
                    ops.set_adam_sms(self.adam_sms if wide else 0)
                    ops.adam_apply(self.w[start:end], self.m[start:end], self.v[start:end], self.g[start:end],
                                   self.w16[start:end], self.hyper, cfg.beta1, cfg.beta2, cfg.epsilon, 1.0,
                                   iterations_inc=self.iterations if (inc_iterations and start == 0) else None)
                    ops.set_adam_sms(0)
                if wide:
                    side[0] -= 1
                    if not side[1]:
                        # from here to the end of backward the conv launches leave the optimiser's SMs alone
                        ops.set_sm_budget(num_sms - self.adam_sms)
                        side[1] = True

        for i in range(n):  # up0 .. up{n-1}
            on_side(lambda: self._conv_pass(f"up{i}/wgrad"))
            self._conv_pass(f"up{i}/dgrad")
            bucket_done(f"up{i}/kernel")
        for i in reversed(range(1, n)):  # down{n-1} .. down1
            on_side(lambda: self._conv_pass(f"down{i}/wgrad"))
            self._conv_pass(f"down{i}/dgrad")
            bucket_done(f"down{i}/kernel")
        ops.conv4s2_c3_wgrad(self.noised, self.gdown_out(0), self.view(self.g, "down0/kernel"), None, accumulate=True)
        # every conv layer's BiasAddGrad in one launch: the pre-activation gradients all still sit in their buffers
        ops.bias_grad_multi(self._bias_plan, accumulate=True)
        if side[1]:
            ops.set_sm_budget(0)
        bucket_done("down0/kernel")
        if sw is not main:
            main.wait_stream(sw)
        if sa is not main and apply_adam:
            # (only when this call forked work onto the optimiser stream: joining a stream that holds nothing of this
            # step would tie a captured graph to uncaptured work)
            main.wait_stream(sa)
        for work in pending:
            work.wait()
        if p2p_used[0] and not dp.dry_run:
            # every rank has finished its fused kernels: all weight writes into my shadow have landed, and nobody still
            # reads my gradient copy (the next step may overwrite it)
            self._p2p["hw"].barrier(channel=1)
        if side[1]:
            ops.set_sm_budget(0)
        if chain_caps:
            _lib.load().gct2_debug_set(9, 0)
            _lib.load().gct2_debug_set(10, 0)
        if dp and not head_done[0]:
            coll.all_reduce(self.loss).wait()

    def _zero_small_grads(self) -> None:
        """One memset for everything the HBM-bound kernels accumulate atomically (param_offsets' head region) + loss."""
        self.g[:self.small].zero_()
        self.loss.zero_()

    # ------------------------------------------------------------------------------------------ public steps
    def _step_body(self, draw: bool, u8: bool = False) -> None:
        cfg = self.cfg
        inv_n = 1.0 / (self.global_batch * cfg.size * cfg.size * 3)
        if u8:
            # the batch as decode_file's uint8 bytes: decode (+ flip) happens inside the prologue launch
            ops.step_begin_u8(self.x_u8, self.flip, self.x, self.noised, self.iterations, self.hyper,
                              self.g[:self.small], self.loss, self.rng_seed, cfg.steps, cfg.base_lr, cfg.warm_up,
                              cfg.beta1, cfg.beta2, t_out=self.t_int, eps_out=self.eps if cfg.target_mode else None)
        elif draw:
            # train.py:224-234 in one launch: t_int ~ U{1..steps} and epsilon ~ N(0,1) drawn on the device (Philox,
            # offset by the optimiser iteration), noising, zeroing of the atomically-accumulated gradients + loss, and
            # this step's Adam alpha; the iteration counter is advanced by the step's last Adam launch
            ops.step_begin(self.x, self.noised, self.iterations, self.hyper, self.g[:self.small], self.loss,
                           self.rng_seed, cfg.steps, cfg.base_lr, cfg.warm_up, cfg.beta1, cfg.beta2, t_out=self.t_int,
                           eps_out=self.eps if cfg.target_mode else None)  # targets other than x need the drawn noise
        else:
            self._zero_small_grads()
            if cfg.mixed_precision:
                # the iteration counter advances only when the update is applied (below): alpha from a scratch copy
                self._it_scratch.copy_(self.iterations)
                ops.adam_prepare(self._it_scratch, self.hyper, cfg.base_lr, cfg.warm_up, cfg.beta1, cfg.beta2)
            else:
                ops.adam_prepare(self.iterations, self.hyper, cfg.base_lr, cfg.warm_up, cfg.beta1, cfg.beta2)
            ops.noise_images(self.x, self.eps, self.t_int, self.noised, cfg.steps)
        self._forward(want_pred=False, backward=True, inv_n=inv_n)
        if cfg.mixed_precision:
            # train.py:82-83 LossScaleOptimizer: the update can only start once EVERY gradient is known to be finite
            self._backward(apply_adam=False)
            ops.loss_scale_check(self.g, self.ls)
            ops.adam_apply(self.w, self.m, self.v, self.g, self.w16, self.hyper, cfg.beta1, cfg.beta2, cfg.epsilon, 1.0,
                           iterations_inc=self.iterations, loss_scale_state=self.ls)
            ops.loss_scale_update(self.ls, cfg.loss_scale_growth)
        else:
            self._backward(apply_adam=True, inc_iterations=draw or u8)

    def train_step_u8(self, img: torch.Tensor, flip: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One training step on the batch as it leaves the reference's decode_file before the cast
        (train.py:285-293): img uint8 [B,S,S,3] (host or device), flip uint8 [B] or None (no mirroring).  A quarter of
        the host-to-device bytes of train_step; t_int / eps are drawn on the device."""
        self._stage_in(self.x_u8, img)
        if flip is None:
            self.flip.zero_()
        else:
            self.flip.copy_(flip, non_blocking=True)
        self.run_step(draw=True, u8=True)
        return self.loss

    def _stage_in(self, dst: torch.Tensor, src: torch.Tensor) -> None:
        """Brings a batch into the step's input buffer.  A pinned host batch crosses PCIe on a copy stream into one of
        two staging buffers -- i.e. while the PREVIOUS step is still computing: nothing on the host waits for a step, so
        the host is a step ahead -- and the step itself only starts with a device-to-device copy (a few microseconds
        instead of the 16-30 us of the transfer on the step's critical path).  Device batches and pageable host memory are
        copied in stream order as before."""
        if src.is_cuda or not src.is_pinned() or os.environ.get("GCT2_STAGE_COPY", "1") == "0":
            dst.copy_(src, non_blocking=True)
            return
        key = (dst.data_ptr(), tuple(src.shape), src.dtype)
        st = self._staging.get(key)
        if st is None:
            st = self._staging[key] = {"buf": [torch.empty_like(dst) for _ in range(2)], "k": 0,
                                       "ready": [torch.cuda.Event() for _ in range(2)],
                                       "free": [torch.cuda.Event() for _ in range(2)], "used": [False, False]}
        k = st["k"]
        st["k"] ^= 1
        main = torch.cuda.current_stream()
        if st["used"][k]:
            self._copy_stream.wait_event(st["free"][k])  # the step that read this staging buffer two calls ago is past it
        with torch.cuda.stream(self._copy_stream):
            st["buf"][k].copy_(src, non_blocking=True)
            st["ready"][k].record(self._copy_stream)
        main.wait_event(st["ready"][k])
        dst.copy_(st["buf"][k], non_blocking=True)
        st["free"][k].record(main)
        st["used"][k] = True

    def set_batch(self, x: torch.Tensor, t_int: Optional[torch.Tensor] = None,
                  eps: Optional[torch.Tensor] = None) -> None:
        """Stages one batch into the engine's input buffers (host tensors are copied asynchronously).  t_int / eps,
        when given, are the injected RNG draws of the parity tests."""
        self._stage_in(self.x, x)
        if t_int is not None:
            self.t_int.copy_(t_int, non_blocking=True)
        if eps is not None:
            self.eps.copy_(eps, non_blocking=True)

    def train_step(self, x: torch.Tensor, t_int: Optional[torch.Tensor] = None,
                   eps: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One training step on the batch x [B,S,S,3] fp32 (host or device).  Returns the device scalar loss
        (global mean over all ranks).  Everything is enqueued on the current stream; nothing synchronises.
        Without t_int/eps the step draws them on the device like the reference (train.py:224-227)."""
        if (t_int is None) != (eps is None):
            raise ValueError("inject both t_int and eps, or neither")
        self.set_batch(x, t_int, eps)
        self.run_step(draw=t_int is None)
        return self.loss

    def _save_state(self):
        return [t.clone() for t in (self.w, self.m, self.v, self.w16, self.iterations, self.ls)]

    def _restore_state(self, saved) -> None:
        for dst, src in zip((self.w, self.m, self.v, self.w16, self.iterations, self.ls), saved):
            dst.copy_(src)
        torch.cuda.current_stream().synchronize()  # see load_weights

    def run_step(self, draw: bool = True, u8: bool = False) -> None:
        """The step on whatever set_batch staged (the part bench.py times as `value`)."""
        if self.dp is not None and self.dp.world > 1 and self.dp.shard_optimizer:
            self._masters_stale = True  # from here on every rank holds current fp32 masters for its own slices only
        if not self.use_graph:
            self._step_body(draw, u8)
            return
        if self._graph is None:
            self._graph = {}
        key = (draw, u8)
        if key not in self._graph:
            # warm up once eagerly on a side stream (lazy inits must not happen under capture), then capture
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                saved = self._save_state()
                self._step_body(draw, u8)
                self._restore_state(saved)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            before = ops.launch_count()
            with torch.cuda.graph(graph):
                self._step_body(draw, u8)
            self._graph_launches = ops.launch_count() - before
            self._graph[key] = graph
        self._graph[key].replay()

    def conv_family_pass(self) -> None:
        """Measurement aid (bench.py's roofline): the 33 tensor-core launches of one step -- 11 fprops, 11 dgrads,
        11 wgrads, in step order -- back to back on the current stream and nothing else, over whatever the buffers
        hold.  Same plans, same programmatic dependent launch chaining as inside the step; the optimiser, the
        CUDA-core kernels and the side streams are left out so that the time is the family's alone."""
        for key in self.plan_keys():
            self._conv_pass(key)

    def release_graphs(self) -> None:
        """Drops the captured step graphs.  Data-parallel callers do this before ``destroy_process_group``: NCCL keeps a
        communicator alive (and its destruction waits) while a CUDA graph that captured one of its collectives
        exists -- seen as a hang at teardown of the captured 2-GPU step."""
        self._graph = None
        torch.cuda.synchronize(self.device)

    def launches_per_step(self) -> int:
        """Kernels of ours one step enqueues (counted inside the C library; a graph replay re-issues the same nodes)."""
        if self.use_graph and self._graph:
            return self._graph_launches
        before = ops.launch_count()
        saved = self._save_state()
        self._step_body(False)
        self._restore_state(saved)
        return ops.launch_count() - before

    def loss_and_grads(self, x, t_int, eps) -> torch.Tensor:
        """Forward + backward without the optimiser update (parity tests compare self.g with the oracle)."""
        self.set_batch(x, t_int, eps)
        inv_n = 1.0 / (self.global_batch * self.cfg.size * self.cfg.size * 3)
        self._zero_small_grads()
        ops.noise_images(self.x, self.eps, self.t_int, self.noised, self.cfg.steps)
        self._forward(want_pred=True, backward=True, inv_n=inv_n)
        self._backward(apply_adam=False)
        return self.loss

    def sample(self, x_theta: torch.Tensor, eps_theta: torch.Tensor, t_values) -> Tuple[torch.Tensor, torch.Tensor]:
        """log_sample's diffusion loops (train.py:365-398 ascending t, :441-468 descending t; predict_x branch): for t in
        t_values:  fake = sqrt(abar_t) x_theta + sqrt(1-abar_t) eps_theta;  x_theta = Denoiser(fake);
        eps_theta = (fake - sqrt(abar_t) x_theta) / sqrt(1-abar_t).  Returns the final (x_theta, eps_theta), fp32
        [B,S,S,3] device tensors owned by the engine.  The whole loop -- len(t_values) forward passes and the fused
        update between them -- is one CUDA graph per schedule (when the engine uses graphs)."""
        t_values = tuple(int(t) for t in t_values)
        if not t_values or min(t_values) < 1 or max(t_values) > self.cfg.steps:
            raise ValueError("t_values must be a non-empty sequence inside [1, steps]")
        if not hasattr(self, "xt"):
            self.xt = torch.zeros_like(self.x)
            self.et = torch.zeros_like(self.x)
        self.xt.copy_(x_theta, non_blocking=True)
        self.et.copy_(eps_theta, non_blocking=True)

        def body():
            steps = self.cfg.steps
            mode = self.cfg.target_mode
            ops.sample_update(None, self.noised, self.xt, self.et, t_values[0], t_values[0], steps, mode)  # first mix
            for k, t in enumerate(t_values):
                self._forward(want_pred=True, backward=False, inv_n=1.0)
                ops.sample_update(self.pred, self.noised, self.xt, self.et, t,
                                  t_values[k + 1] if k + 1 < len(t_values) else 0, steps, mode)

        if not self.use_graph:
            body()
            return self.xt, self.et
        if self._graph is None:
            self._graph = {}
        key = ("sample", t_values)
        if key not in self._graph:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                saved = (self.xt.clone(), self.et.clone())
                body()
                self.xt.copy_(saved[0])
                self.et.copy_(saved[1])
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                body()
            self._graph[key] = graph
        self._graph[key].replay()
        return self.xt, self.et

    def denoise(self, noised: torch.Tensor) -> torch.Tensor:
        """Denoiser.call (train.py:206-215): forward only on an already-noised image; returns the fp32 prediction."""
        self.noised.copy_(noised, non_blocking=True)
        self.x.copy_(self.noised)
        self.loss.zero_()
        self._forward(want_pred=True, backward=False, inv_n=1.0)
        return self.pred


def make_engine(cfg: NetConfig, batch: int, **kw) -> UNetEngine:
    """The engine for a configuration: the tuned UNetEngine for the reference's default wiring, BlockUNetEngine for the
    dormant switches (block_depth > 0, concat = False, residual = True)."""
    if cfg.fused_default:
        return UNetEngine(cfg, batch, **kw)
    from .block_engine import BlockUNetEngine
    return BlockUNetEngine(cfg, batch, **kw)
