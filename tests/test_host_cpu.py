"""CPU: host-side logic of the drop-in surface (no kernels are launched)."""
import math
import os

import pytest
import torch

from gan_class_transfer2_b200 import engine as E
from gan_class_transfer2_b200 import train as T
from oracle import oracle as O


def test_train_py_names_and_defaults():
    for name, val in dict(size=256, pixel_size=128, max_size=512, block_depth=0, octaves=6, batch_size=1, steps=200,
                          residual=False, concat=True, predict_x=True, mixed_precision=False, warm_up=2000,
                          test_step=25).items():
        assert getattr(T, name) == val, name
    for cls in ("WarmUp", "Residual", "Block", "UpShuffle", "DownShuffle", "Denoiser", "Trainer"):
        assert isinstance(getattr(T, cls), type)
    assert callable(T.identity) and callable(T.alpha_dash)
    assert isinstance(T.optimizer.learning_rate, T.WarmUp)
    assert (T.optimizer.beta_1, T.optimizer.beta_2, T.optimizer.epsilon) == (0.9, 0.999, 1e-7)


def test_construction_recursion_matches_reference_shapes():
    d = T.Denoiser()
    cfg = d.net_config(256)
    assert cfg.down_filters == (128, 256, 512, 512, 512, 512)
    assert cfg.up_filters == (64, 128, 256, 512, 512, 512)
    assert E.variable_specs(cfg) == O.variable_specs(O.DEFAULT)
    offsets, total = E.param_offsets(cfg)
    assert sum(cnt for _, cnt in offsets.values()) == 41_691_660
    assert 0 <= total - 41_691_660 < 64 and total % 4 == 0          # head-region padding only
    assert all(off % 64 == 0 for name, (off, _) in offsets.items() if name.endswith("kernel") and name[:5] not in
               ("dense", "down0"))
    # no overlaps, and the head region holds exactly the atomically-accumulated variables
    spans = sorted((off, off + cnt, name) for name, (off, cnt) in offsets.items())
    assert all(spans[i][1] <= spans[i + 1][0] for i in range(len(spans) - 1))
    small = E.small_region(cfg)
    assert {n for n, (off, _) in offsets.items() if off < small} == set(E.small_names(cfg))


def test_variable_specs_follow_formulas_for_the_widened_variant():
    cfg = E.NetConfig(size=512, pixel_size=256, max_size=1024, octaves=7)
    ocfg = O.Config(size=512, pixel_size=256, max_size=1024, octaves=7)
    assert E.variable_specs(cfg) == O.variable_specs(ocfg)
    assert sum(cnt for _, cnt in E.param_offsets(cfg)[0].values()) == 217_078_796


def test_warmup_and_alpha_dash_match_oracle():
    wu, ou = T.WarmUp(2e-5, 2000), O.WarmUp(2e-5, 2000)
    for s in (0, 1, 999, 1999, 2000, 5000):
        assert math.isclose(wu(s), ou(s), rel_tol=1e-6)
    for t in (1, 25, 200):
        assert math.isclose(T.alpha_dash(t), O.alpha_dash(t))
    tt = torch.tensor([1.0, 100.0])
    assert torch.allclose(T.alpha_dash(tt), O.alpha_dash(tt))


def test_unsupported_structures_are_rejected():
    d = T.Denoiser()
    d.middle.layers[3] = T.Dense(5)
    with pytest.raises(NotImplementedError):
        d.net_config(256)
    with pytest.raises(ValueError):
        E.NetConfig(size=48, octaves=2).validate()
    with pytest.raises(ValueError):
        E.NetConfig(size=64, octaves=5).validate()  # bottleneck below 4x4
    with pytest.raises(ValueError):
        E.NetConfig(size=64, octaves=2, pixel_size=96).validate()


def test_wiring_switches_reach_the_engine_config():
    """train.py:20,26-27: block_depth and concat change what Denoiser.__init__ builds; both are run by the layer-list
    engine (block_engine.BlockUNetEngine) and must arrive in the engine's configuration with the oracle's variable list,
    and so does residual=True (train.py:106-112: the bias-free Dense projection added to the Residual's input)."""
    import dataclasses
    saved = {k: getattr(T, k) for k in ("concat", "residual", "block_depth", "octaves", "max_size")}
    try:
        T.octaves, T.max_size = 4, 256
        for kw in (dict(block_depth=1), dict(block_depth=2, concat=False), dict(concat=False), dict(residual=True),
                   dict(residual=True, block_depth=1)):
            T.block_depth, T.concat, T.residual = kw.get("block_depth", 0), kw.get("concat", True), kw.get("residual", False)
            cfg = T.Denoiser().net_config(64)
            assert (cfg.block_depth, cfg.concat, cfg.residual) == (T.block_depth, T.concat, T.residual)
            assert not cfg.fused_default
            cfg.validate()
            ocfg = dataclasses.replace(O.TINY, **kw)
            assert E.variable_specs(cfg) == O.variable_specs(ocfg)
            assert sum(cnt for _, cnt in E.param_offsets(cfg)[0].values()) == O.param_count(ocfg)
            with pytest.raises(NotImplementedError):
                E.UNetEngine(cfg, 1)  # the tuned engine is for the default wiring only
    finally:
        for k, v in saved.items():
            setattr(T, k, v)
    assert T.Denoiser().net_config(256).fused_default


def test_small_region_holds_everything_accumulated_with_atomics():
    cfg = E.NetConfig(size=64, pixel_size=128, max_size=256, octaves=4, block_depth=2)
    names = E.small_names(cfg)
    assert names[0] == "block_in/conv0/kernel" and "down0/kernel" not in names  # down0 reads 128 channels here
    assert all(n.endswith("bias") or n.startswith("dense") or n == names[0] for n in names)
    offs, total = E.param_offsets(cfg)
    small = E.small_region(cfg)
    assert all((off < small) == (name in names) for name, (off, _) in offs.items())
    assert small % 64 == 0 and total % 4 == 0


def test_compile_hands_the_optimizer_to_the_denoiser():
    d = T.Denoiser()
    tr = T.Trainer(d)
    tr.compile(T.Adam(T.WarmUp(1e-3, 7), beta_1=0.8), T.identity)
    cfg = d.net_config(256)
    assert (cfg.base_lr, cfg.warm_up, cfg.beta1) == (1e-3, 7, 0.8)
    with pytest.raises(NotImplementedError):
        tr.compile(object(), T.identity)


def test_mixed_precision_switch_reaches_the_engine_config():
    """train.py:34,43-45,82-83: `mixed_precision = True` selects fp16 storage and the LossScaleOptimizer's parameters."""
    old = T.mixed_precision
    T.mixed_precision = True
    try:
        d = T.Denoiser()
        tr = T.Trainer(d)
        tr.compile(T.LossScaleOptimizer(T.Adam(T.WarmUp(2e-5, 2000)), initial_scale=2.0 ** 12, dynamic_growth_steps=7),
                   T.identity)
        cfg = d.net_config(256)
        assert cfg.mixed_precision and cfg.loss_scale_init == 2.0 ** 12 and cfg.loss_scale_growth == 7
        assert (cfg.base_lr, cfg.warm_up) == (2e-5, 2000) and T.preferred_type() == torch.float16
    finally:
        T.mixed_precision = old
    assert not T.Denoiser().net_config(256).mixed_precision and T.preferred_type() == torch.bfloat16


def test_layers_refuse_cpu_tensors():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        T.DownShuffle(64)(torch.zeros(1, 8, 8, 64))


def test_buckets_tile_the_flat_gradient_buffer_tail_to_head():
    for cfg in (E.NetConfig(), E.NetConfig(size=64, max_size=256, octaves=4)):
        total = E.param_offsets(cfg)[1]
        for bb in (1 << 20, 16 << 20, 48 << 20, 1 << 40):
            b = E.grad_buckets(cfg, bb)
            assert b[0][1] == total and b[-1][0] == 0
            assert all(b[i][0] == b[i + 1][1] for i in range(len(b) - 1))
            assert b[-1][2] == "down0/kernel"


def test_shard_batch():
    assert E.shard_batch(8, 4, 1) == (2, 4)
    with pytest.raises(ValueError):
        E.shard_batch(6, 4, 0)


def test_glorot_limits():
    g = torch.Generator().manual_seed(0)
    k = E.glorot_uniform((4, 4, 128, 256), g)
    lim = math.sqrt(6 / (16 * 128 + 16 * 256))
    assert k.abs().max() <= lim and k.abs().max() > 0.99 * lim


def test_optimizer_shards_tile_every_bucket():
    """Sharded optimiser: for every bucket of the default and the tiny model and 2/4/8 ranks, the slices of
    all ranks tile the part above the replicated head region exactly once, on float4 boundaries; the head region is
    never sharded."""
    from gan_class_transfer2_b200 import engine as E
    for cfg in (E.NetConfig(), E.NetConfig(size=64, pixel_size=128, max_size=256, octaves=4)):
        small = E.small_region(cfg)
        _, total = E.param_offsets(cfg)
        for world in (2, 4, 8):
            covered = 0
            for start, end, _ in E.grad_buckets(cfg, 48 << 20):
                cuts = [E.optimizer_shard(start, end, small, world, r) for r in range(world)]
                assert all(c is not None for c in cuts), (start, end, world)
                lo = max(start, small)
                assert [c[2] for c in cuts] == [lo + r * (end - lo) // world for r in range(world)]
                assert cuts[-1][3] == end and all(a[3] == b[2] for a, b in zip(cuts, cuts[1:]))
                assert all(c[2] % 4 == 0 and (c[3] - c[2]) % 4 == 0 for c in cuts)
                covered += end - lo
            assert covered == total - small
    assert E.optimizer_shard(0, 100, 128, 2, 0) is None      # nothing above the head region
    assert E.optimizer_shard(128, 128 + 12, 128, 8, 0) is None  # not divisible: stays replicated


def test_fast_division_magic_numbers():
    """The multiply-shift constants of csrc/conv_umma.cuh:make_fastdiv (index decoding without integer divisions),
    restated here: q = umulhi(n, mul) >> shr with p = 31 + ceil(log2 d), mul = ceil(2^p / d), shr = p - 32 must equal
    n // d for every 0 <= n < 2^31 -- checked on edge values and random samples for small, power-of-two-adjacent and
    random divisors (non-power-of-two divisors occur with channel counts like 192 = 3 x 64)."""
    import random
    rnd = random.Random(0)

    def make(d):
        if d <= 1:
            return d, 0, 0
        lg = (d - 1).bit_length()
        p = 31 + lg
        mul = ((1 << p) + d - 1) // d
        assert mul < (1 << 32)
        return d, mul, p - 32

    def div(f, n):
        d, mul, shr = f
        return n if d == 1 else ((n * mul) >> 32) >> shr

    divisors = list(range(1, 600)) + [2 ** k + s for k in range(1, 20) for s in (-1, 0, 1) if 2 ** k + s > 0]
    divisors += [rnd.randrange(1, 1 << 20) for _ in range(300)]
    for d in divisors:
        f = make(d)
        ns = [0, 1, d - 1, d, d + 1, 2 * d - 1, 2 * d, (1 << 31) - 1, (1 << 31) - d, 1 << 30]
        ns += [rnd.randrange(0, 1 << 31) for _ in range(50)]
        for n in ns:
            if 0 <= n < (1 << 31):
                assert div(f, n) == n // d, (d, n)


def test_plan_table_is_keyed_by_shape_batch_policy_and_sm_count(tmp_path, monkeypatch):
    """tuned_plans.json (tools/tune_plans.py): a table applies to exactly the configuration it was measured on."""
    import json
    cfg = E.NetConfig()
    key = E.plan_table_key(cfg, 1, 148)
    assert key != E.plan_table_key(cfg, 8, 148) and key != E.plan_table_key(cfg, 1, 132)
    assert key != E.plan_table_key(E.NetConfig(mixed_precision=True), 1, 148)
    assert key != E.plan_table_key(E.NetConfig(size=64, max_size=256, octaves=4), 1, 148)
    shipped = E.load_tuned_plans(cfg, 1, 148)
    assert shipped and all(len(v) == 2 and v[0] in (64, 128, 256) and v[1] >= 1 for v in shipped.values())
    layers = {f"down{i}" for i in range(1, 6)} | {f"up{i}" for i in range(6)}
    assert all(k.split("/")[0] in layers and k.split("/")[1] in ("fprop", "dgrad", "wgrad") for k in shipped)
    assert E.load_tuned_plans(cfg, 3, 148) == {}  # no table for this batch: the cost model decides


def test_ncu_summary_reads_both_csv_shapes(tmp_path):
    """tools/ncu_summary.py: the --metrics log (one row per launch and metric) and the raw page (one row per launch)."""
    import subprocess
    import sys
    long_csv = tmp_path / "long.csv"
    long_csv.write_text(
        '==PROF== Connected\n'
        '"ID","Process ID","Process Name","Host Name","Kernel Name","Context","Stream","Block Size","Grid Size","Device","CC",'
        '"Section Name","Metric Name","Metric Unit","Metric Value"\n'
        + "".join(f'"{i}","1","python","h","void gct2::conv_umma_kernel<0, 64, 0, 0>(CUtensorMap_st)","1","7","(384, 1, 1)",'
                  f'"(128, 1, 1)","0","10.0","s","{m}","{u}","{v}"\n'
                  for i in range(2) for m, u, v in (("gpu__time_duration.sum", "ns", "12000"), ("dram__bytes_read.sum", "Mbyte", "5"),
                                                    ("dram__bytes_write.sum", "Kbyte", "250"))))
    out = tmp_path / "o.csv"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run([sys.executable, os.path.join(root, "tools", "ncu_summary.py"), "--raw", str(long_csv), "--out", str(out)],
                   check=True, capture_output=True)
    rows = out.read_text().strip().split("\n")
    assert rows[0].startswith("kernel,time_us,grid,block") and len(rows) == 3
    cells = rows[1].split('",')[1].split(",")
    assert cells[0] == "12.000" and cells[1] == "128.000" and cells[2] == "384.000" and cells[3] == "5.000" and cells[4] == "0.250"
    wide = tmp_path / "wide.csv"
    wide.write_text('"ID","Kernel Name","Block Size","SM_A.Sec.gpu__time_duration.sum","FBSP.Sec.dram__bytes_read.sum"\n'
                    '"","","","us","byte"\n"0","k<1>(int)","(384, 1, 1)","7.5","1000000"\n')
    subprocess.run([sys.executable, os.path.join(root, "tools", "ncu_summary.py"), "--raw", str(wide), "--out", str(out)],
                   check=True, capture_output=True)
    assert out.read_text().strip().split("\n")[1] == '"k<1>",7.500,1.000'
