"""-m gpu: every CUDA kernel, called through the C ABI, against the CPU oracle on the same seeded inputs."""
import pytest
import torch

from tests import kernel_checks as K

pytestmark = pytest.mark.gpu


def _id(case):
    fn, kw = case
    return fn.__name__.replace("check_", "") + "-" + "-".join(f"{k}{v}" for k, v in kw.items())


@pytest.mark.parametrize("case", K.CONV_CASES, ids=_id)
def test_conv_family_matches_oracle(case):
    fn, kw = case
    m = fn(**kw)
    assert m["err"] <= m["tol"], m
    assert m.get("pad_intact", True), f"kernel wrote outside its channel slice: {m}"


@pytest.mark.parametrize("case", K.EW_CASES, ids=_id)
def test_hbm_bound_kernels_match_oracle(case):
    fn, kw = case
    m = fn(**kw)
    assert m["err"] <= m["tol"], m
    assert m.get("pad_intact", True), m


def _fid(c):
    f = c[2]
    return (c[0].__name__.replace("check_", "") + "-" + "-".join(f"{k}{v}" for k, v in c[1].items() if k != "seed")
            + "-" + "-".join(f"{k}{v}" for k, v in f.items()))


@pytest.mark.parametrize("case", K.FORCED_CASES, ids=_fid)
def test_every_tile_width_and_split_k_path(case):
    fn, kw, force = case
    m = K.forced(fn, **force, **kw)
    assert m["err"] <= m["tol"], m
    assert m.get("pad_intact", True), m


@pytest.mark.parametrize("case", K.PAIR_CASES, ids=_fid)
def test_cta_pair_kernels_match_oracle(case):
    """All six cta_group::2 instantiations against the oracle (these were selected automatically for
    every launch with more tiles than SMs but only ever reached by a property test)."""
    fn, kw, force = case
    m = K.forced(fn, **force, **kw)
    assert m["plan"]["pair"] == 1, m
    assert m["err"] <= m["tol"], m
    assert m.get("pad_intact", True), m


# ---- stride-1 convs of Block (train.py:131-139, block_depth > 0)
@pytest.mark.parametrize("case", K.S1_CASES, ids=_fid)
def test_stride1_conv_family_matches_oracle(case):
    fn, kw, force = case
    m = K.forced(fn, **force, **kw)
    assert m["err"] <= m["tol"], m
    assert m.get("pad_intact", True), m


@pytest.mark.parametrize("case", K.S1_EW_CASES, ids=_id)
def test_block_cuda_core_kernels_match_oracle(case):
    fn, kw = case
    m = fn(**kw)
    assert m["err"] <= m["tol"], m
    assert m.get("pad_intact", True), m


def test_stride1_conv_family_fp16_storage():
    with K.half_format(torch.float16):
        for fn, kw in ((K.check_conv3_fprop, dict(B=2, H=16, Cin=128, Cout=256)),
                       (K.check_conv3_dgrad, dict(B=2, H=16, Cin=256, Cout=128, add_old=True)),
                       (K.check_conv3_wgrad, dict(B=2, H=16, Cin=128, Cout=256)), (K.check_conv3_c3, {}),
                       (K.check_dense_mse_noimage, {})):
            m = fn(**kw)
            assert m["err"] <= m["tol"], m
            assert m.get("pad_intact", True), m


# ---- fp16 storage (train.py:34,43-45: the reference's own reduced-precision mode)
F16_CASES = [
    (K.check_conv_fprop, dict(B=2, H=16, Cin=128, Cout=256), dict(BN=64, splits=1)),
    (K.check_conv_fprop, dict(B=2, H=16, Cin=128, Cout=256), dict(BN=128, splits=4, finish="l2")),
    (K.check_conv_fprop, dict(B=2, H=16, Cin=128, Cout=256), dict(BN=256, splits=8, finish="cluster")),
    (K.check_conv_fprop, dict(B=2, H=16, Cin=128, Cout=256), dict(BN=128, splits=4, nofuse=1)),
    (K.check_convT_fprop, dict(B=2, H=8, Cin=128, Cout=256), dict(BN=128, splits=1)),
    (K.check_convT_fprop, dict(B=2, H=16, Cin=128, Cout=256), dict(BN=256, splits=1, pair=1)),
    (K.check_conv_dgrad, dict(B=2, H=16, Cin=256, Cout=128, add_old=True), dict(BN=128, splits=1)),
    (K.check_conv_dgrad, dict(B=2, H=16, Cin=256, Cout=128, add_old=True), dict(BN=256, splits=4, finish="l2")),
    (K.check_convT_dgrad, dict(B=2, H=8, Cin=256, Cout=128), dict(BN=64, splits=4, finish="cluster")),
    (K.check_convT_dgrad, dict(B=2, H=16, Cin=256, Cout=128), dict(BN=128, splits=1, pair=1)),
    (K.check_conv_wgrad, dict(B=4, H=32, Cin=128, Cout=256), dict(BN=128, splits=1)),
    (K.check_conv_wgrad, dict(B=2, H=32, Cin=256, Cout=256), dict(BN=256, splits=1, pair=1)),
    (K.check_convT_wgrad, dict(B=4, H=16, Cin=256, Cout=64), dict(BN=64, splits=4)),
]


@pytest.mark.parametrize("case", F16_CASES, ids=_fid)
def test_conv_family_fp16_storage(case):
    fn, kw, force = case
    with K.half_format(torch.float16):
        m = K.forced(fn, **force, **kw)
    assert m["err"] <= m["tol"], m
    assert m.get("pad_intact", True), m


@pytest.mark.parametrize("case", [(K.check_c3_fprop, {}), (K.check_c3_wgrad, {}), (K.check_bias_grad, {}),
                                  (K.check_bias_grad_multi, {}), (K.check_dense_mse, {}),
                                  (K.check_dense_mse, dict(B=1, H=16, Cu=128)), (K.check_adam, {})], ids=_id)
def test_hbm_bound_kernels_fp16_storage(case):
    fn, kw = case
    with K.half_format(torch.float16):
        m = fn(**kw)
    assert m["err"] <= m["tol"], m
    assert m.get("pad_intact", True), m


def test_loss_scale_kernels_follow_keras_dynamics():
    """gct2_loss_scale_check / gct2_adam_apply(loss_scale_state) / gct2_loss_scale_update against the oracle's
    DynamicLossScale (tf.keras.mixed_precision.LossScaleOptimizer, train.py:82-83): an inf or NaN anywhere skips the whole
    update and halves the scale; `growth` good steps in a row double it; gradients are unscaled before Adam."""
    from gan_class_transfer2_b200 import ops
    from oracle import oracle as O
    dev = torch.device("cuda", 0)
    n = 8192
    g = torch.Generator().manual_seed(3)
    w0 = torch.randn(n, generator=g)
    cfg = O.Config(warm_up=0)
    ref = O.DynamicLossScale(2.0 ** 10, growth_steps=2)
    ls = torch.tensor([2.0 ** 10, 0.0, 1.0, 2.0 ** -10], device=dev)
    w, m, v = w0.clone().to(dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    wo, mo, vo = w0.clone(), torch.zeros(n), torch.zeros(n)
    wb = torch.zeros(n, dtype=torch.float16, device=dev)
    it = torch.zeros(1, dtype=torch.int64, device=dev)
    scratch = torch.zeros(1, dtype=torch.int64, device=dev)
    hyper = torch.zeros(2, device=dev)
    it_ref = 0
    poison = {1: float("inf"), 4: float("nan"), 5: float("-inf")}
    for step in range(8):
        grad = torch.randn(n, generator=g) * 1e-3
        scaled = grad * ref.scale
        if step in poison:
            scaled[(step * 977) % n] = poison[step]
        scratch.copy_(it)
        ops.adam_prepare(scratch, hyper, cfg.base_lr, cfg.warm_up)
        gd = scaled.to(dev)
        ops.loss_scale_check(gd, ls)
        ops.adam_apply(w, m, v, gd, wb, hyper, iterations_inc=it, loss_scale_state=ls)
        ops.loss_scale_update(ls, 2)
        scale_used = ref.scale
        applied = ref.update(step not in poison)
        if applied:  # the scaled gradient is an fp16 tensor in the reference (cast to fp32, then unscaled)
            O.keras_adam_update(wo, mo, vo, scaled.to(torch.float16).float() / scale_used, it_ref, cfg)
            it_ref += 1
        torch.cuda.synchronize()
        assert int(it) == it_ref, (step, int(it), it_ref)
        assert float(ls[0]) == ref.scale and float(ls[1]) == ref.good_steps and float(ls[2]) == 1.0, (step, ls.tolist(), ref.scale)
        assert float((w.cpu() - wo).abs().max()) <= 2e-7, step
        assert torch.isfinite(w).all() and torch.isfinite(m).all() and torch.isfinite(v).all()
    assert float((wb.float().cpu() - wo).abs().max()) <= 2e-3   # the fp16 shadow follows the masters
