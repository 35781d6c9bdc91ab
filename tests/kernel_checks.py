"""Per-kernel parity checks: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Shared by tests/test_kernels_gpu.py (pytest, -m gpu) and tools/diag_kernels.py (prints every metric even when a
case fails).  Each check returns {"name", "err": rel-L2 error, "max": max abs error / max |ref|, "tol"}.
Tolerances: bf16 outputs carry one rounding (2^-9 relative) on top of fp32 accumulation-order noise -> rel-L2
<= 4e-3; fp32 outputs (weight/bias gradients, loss) only differ by accumulation order -> rel-L2 <= 2e-5 * sqrt(K).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from oracle import oracle as O

BF16_TOL = 4e-3
F32_TOL = 3e-4
#: the 16-bit storage format the checks run in: bf16 (default) or fp16 (the reference's mixed_float16 policy, train.py:34,
#: 43-45); see half_format()
HALF = torch.bfloat16


class half_format:
    """with half_format(torch.float16): ...  -- runs the checks with fp16 tensors (the library follows the tensors' dtype)."""

    def __init__(self, dtype):
        self.dtype = dtype

    def __enter__(self):
        global HALF
        self.old, HALF = HALF, self.dtype

    def __exit__(self, *exc):
        global HALF
        HALF = self.old


def _dev():
    return torch.device("cuda", 0)


def _bf(t):
    return t.to(HALF)


def _metrics(name, got, ref, tol):
    got = got.detach().float().cpu()
    ref = ref.detach().float().cpu()
    denom = ref.norm().item() or 1.0
    err = (got - ref).norm().item() / denom
    mx = (got - ref).abs().max().item() / (ref.abs().max().item() or 1.0)
    bad = not math.isfinite(err)
    return {"name": name, "err": float("inf") if bad else err, "max": mx, "tol": tol}


def _rand(shape, gen, scale=1.0):
    return (torch.randn(shape, generator=gen) * scale)


def _ops():
    from gan_class_transfer2_b200 import ops
    return ops


def _slice_buf(B, H, W, C, pad_front, pad_back, dev, fill=None):
    """A [B,H,W,C] bf16 view living inside a wider NHWC buffer (exercises pixel strides / concat slices)."""
    full = torch.full((B, H, W, pad_front + C + pad_back), 7.0 if fill is None else fill, dtype=HALF,
                      device=dev)
    return full, full[..., pad_front:pad_front + C]


# ---------------------------------------------------------------------------------------------- stride-1 conv (Block)
def _s1_ref(x, w, ks):
    """Keras Conv2D(filters, ks, 1, 'same') without bias / activation, NHWC in, NHWC out (train.py:131-139)."""
    return F.conv2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), None, stride=1, padding=ks // 2).permute(0, 2, 3, 1)


def check_conv3_fprop(B, H, Cin, Cout, ks=3, seed=20, pad=64, weights_stable=False):
    ops = _ops()
    g = torch.Generator().manual_seed(seed)
    x = _bf(_rand((B, H, H, Cin), g))
    w = _bf(_rand((ks, ks, Cin, Cout), g, 1.0 / math.sqrt(ks * ks * Cin)))
    b = _rand((Cout,), g, 0.1)
    ref = torch.relu(_s1_ref(x.float(), w.float(), ks) + b)
    dev = _dev()
    _, xv = _slice_buf(B, H, H, Cin, pad, 0, dev)
    xv.copy_(x)
    yfull, yv = _slice_buf(B, H, H, Cout, 0, pad, dev)
    ws = ops.Workspace(64 << 20, dev)
    ops.conv3s1_fprop(xv, w.to(dev), b.to(dev), yv, ws, weights_stable)
    torch.cuda.synchronize()
    m = _metrics(f"conv3s1_fprop ks{ks} B{B} H{H} {Cin}->{Cout}", yv, ref, BF16_TOL)
    m["pad_intact"] = bool((yfull[..., Cout:] == 7.0).all().item()) if pad else True
    return m


def check_conv3_dgrad(B, H, Cin, Cout, ks=3, mask_channels=None, add_old=False, seed=21, pad=64, weights_stable=False):
    ops = _ops()
    g = torch.Generator().manual_seed(seed)
    mask_channels = Cin if mask_channels is None else mask_channels
    dy = _bf(_rand((B, H, H, Cout), g))
    w = _bf(_rand((ks, ks, Cin, Cout), g, 1.0 / math.sqrt(ks * ks * Cin)))
    act = _bf(_rand((B, H, H, Cin), g).clamp_min(0))
    old = _bf(_rand((B, H, H, Cin), g))
    xr = torch.zeros(B, H, H, Cin, requires_grad=True)
    _s1_ref(xr, w.float(), ks).backward(dy.float())
    ref = xr.grad
    if add_old:
        ref = ref + old.float()
    keep = act.float() > 0
    keep[..., mask_channels:] = True
    ref = ref * keep
    dev = _dev()
    _, dyv = _slice_buf(B, H, H, Cout, 0, pad, dev)
    dyv.copy_(dy)
    dxfull, dxv = _slice_buf(B, H, H, Cin, pad, 0, dev)
    dxv.copy_(old)
    _, actv = _slice_buf(B, H, H, Cin, pad, 0, dev)
    actv.copy_(act)
    ws = ops.Workspace(64 << 20, dev)
    ops.conv3s1_dgrad(dyv, w.to(dev), dxv, actv, mask_channels, add_old, ws, weights_stable)
    torch.cuda.synchronize()
    m = _metrics(f"conv3s1_dgrad ks{ks} B{B} H{H} {Cin}<-{Cout} mask{mask_channels} add{int(add_old)}", dxv, ref, BF16_TOL)
    m["pad_intact"] = bool((dxfull[..., :pad] == 7.0).all().item()) if pad else True
    return m


def check_conv3_wgrad(B, H, Cin, Cout, ks=3, seed=22, pad=64):
    ops = _ops()
    g = torch.Generator().manual_seed(seed)
    x = _bf(_rand((B, H, H, Cin), g))
    dy = _bf(_rand((B, H, H, Cout), g))
    wr = torch.zeros(ks, ks, Cin, Cout, requires_grad=True)
    _s1_ref(x.float(), wr, ks).backward(dy.float())
    dev = _dev()
    _, xv = _slice_buf(B, H, H, Cin, pad, 0, dev)
    xv.copy_(x)
    _, dyv = _slice_buf(B, H, H, Cout, 0, pad, dev)
    dyv.copy_(dy)
    dw = torch.full((ks, ks, Cin, Cout), 3.0, device=dev)
    ws = ops.Workspace(64 << 20, dev)
    ops.conv3s1_wgrad(xv, dyv, dw, ws)
    torch.cuda.synchronize()
    return _metrics(f"conv3s1_wgrad ks{ks} B{B} H{H} {Cin}x{Cout}", dw, wr.grad, F32_TOL * 2)


def check_proj_add(B, H, Cin, Cout, seed=25, pad=64):
    """train.py:110-111: y = res + Dense(Cout, use_bias=False)(x), the 1x1 map with the add in the epilogue."""
    ops = _ops()
    g = torch.Generator().manual_seed(seed)
    x = _bf(_rand((B, H, H, Cin), g))
    res = _bf(_rand((B, H, H, Cout), g))
    w = _bf(_rand((Cin, Cout), g, 1.0 / math.sqrt(Cin)))
    ref = res.float() + x.float() @ w.float()
    dev = _dev()
    _, xv = _slice_buf(B, H, H, Cin, pad, 0, dev)
    xv.copy_(x)
    _, rv = _slice_buf(B, H, H, Cout, 0, pad, dev)
    rv.copy_(res)
    yfull, yv = _slice_buf(B, H, H, Cout, 0, pad, dev)
    ops.conv3s1_fprop_add(xv, w.to(dev).view(1, 1, Cin, Cout), rv, yv)
    torch.cuda.synchronize()
    m = _metrics(f"conv3s1_fprop_add B{B} H{H} {Cin}->{Cout}", yv, ref, BF16_TOL)
    m["pad_intact"] = bool((yfull[..., Cout:] == 7.0).all().item())
    return m


def check_res0_fold(U=64, B=2, H=16, seed=26):
    """The image-level residual folded into Dense(3): pred, loss and all four gradients (dWp, dWd, dbd, du0) of
    pred = (noised + u0 . Wp) . Wd + bd through gct2_res0_compose -> gct2_dense_mse -> gct2_res0_decompose."""
    ops = _ops()
    g = torch.Generator().manual_seed(seed)
    u0 = _bf(_rand((B, H, H, U), g).clamp_min(0))
    noised, x = _rand((B, H, H, 3), g), _rand((B, H, H, 3), g)
    wp = _rand((U, 3), g, 0.2).requires_grad_(True)
    wd = _rand((3, 3), g, 0.5).requires_grad_(True)
    bd = _rand((3,), g, 0.1).requires_grad_(True)
    u0r = u0.float().requires_grad_(True)
    pred = (noised + u0r @ wp) @ wd + bd
    loss = ((x - pred) ** 2).mean()
    loss.backward()
    dev = _dev()
    u0v = u0.to(dev)
    weff = torch.zeros(U + 3, 3, device=dev)
    dweff = torch.zeros(U + 3, 3, device=dev)
    du0 = torch.full((B, H, H, U), 7.0, dtype=HALF, device=dev)
    predg, lossg, dbd = torch.empty(B, H, H, 3, device=dev), torch.zeros(1, device=dev), torch.zeros(3, device=dev)
    dwp, dwd = torch.full((U, 3), 3.0, device=dev), torch.full((3, 3), 3.0, device=dev)
    wpd, wdd = wp.detach().to(dev), wd.detach().to(dev)
    ops.res0_compose(wpd, wdd, weff)
    ops.dense_mse(u0v, noised.to(dev), x.to(dev), weff, bd.detach().to(dev), lossg, 1.0 / (B * H * H * 3), pred=predg,
                  du0=du0, dwd=dweff, dbd=dbd, accumulate=True)
    ops.res0_decompose(dweff, wpd, wdd, dwp, dwd)
    torch.cuda.synchronize()
    ms = [_metrics("pred", predg, pred, 2e-5), _metrics("loss", lossg, loss.reshape(1), 2e-5),
          _metrics("du0", du0, u0r.grad * (u0.float() > 0), BF16_TOL), _metrics("dWp", dwp, wp.grad, F32_TOL),
          _metrics("dWd", dwd, wd.grad, F32_TOL), _metrics("dbd", dbd, bd.grad, F32_TOL)]
    worst = dict(max(ms, key=lambda m: m["err"] / m["tol"]))
    worst["name"] = f"image-level residual folded into Dense(3) U{U} (worst: {worst['name']})"
    return worst


def check_conv3_c3(B=2, H=32, Cout=128, seed=23):
    """The 3-channel 3x3 / stride-1 conv of the outermost Block: forward and weight gradient."""
    ops = _ops()
    g = torch.Generator().manual_seed(seed)
    x = _rand((B, H, H, 3), g)
    w = _rand((3, 3, 3, Cout), g, 0.2)
    b = _rand((Cout,), g, 0.1)
    dz = _bf(_rand((B, H, H, Cout), g))
    wr = w.clone().requires_grad_(True)
    pre = _s1_ref(x, wr, 3) + b
    pre.backward(dz.float())
    dev = _dev()
    yfull, yv = _slice_buf(B, H, H, Cout, 64, 0, dev)
    ops.conv3s1_c3_fprop(x.to(dev), w.to(dev), b.to(dev), yv)
    _, dzv = _slice_buf(B, H, H, Cout, 0, 64, dev)
    dzv.copy_(dz)
    dw = torch.full((3, 3, 3, Cout), 3.0, device=dev)
    ops.conv3s1_c3_wgrad(x.to(dev), dzv, dw)
    dw2 = torch.full((3, 3, 3, Cout), 0.5, device=dev)
    ops.conv3s1_c3_wgrad(x.to(dev), dzv, dw2, accumulate=True)
    torch.cuda.synchronize()
    ms = [_metrics(f"conv3s1_c3_fprop B{B} H{H} 3->{Cout}", yv, torch.relu(pre.detach()), BF16_TOL),
          _metrics("conv3s1_c3_wgrad", dw, wr.grad, F32_TOL),
          _metrics("conv3s1_c3_wgrad accumulate", dw2, wr.grad + 0.5, F32_TOL)]
    worst = dict(max(ms, key=lambda m: m["err"] / m["tol"]))
    worst["name"] = f"conv3s1_c3 B{B} H{H} Cout{Cout} (worst: {worst['name']})"
    worst["pad_intact"] = bool((yfull[..., :64] == 7.0).all().item())
    return worst


def check_dense_mse_noimage(B=2, H=32, Cu=128, seed=24):
    """Dense(3) + MSE on the 16-bit channels only (behind a Block / without the concat skip)."""
    ops = _ops()
    g = torch.Generator().manual_seed(seed)
    u0 = _bf(_rand((B, H, H, Cu), g).clamp_min(0))
    x = _rand((B, H, H, 3), g)
    wd = _rand((Cu, 3), g, 0.2).requires_grad_(True)
    bd = _rand((3,), g, 0.1).requires_grad_(True)
    u0r = u0.float().requires_grad_(True)
    pred = O.dense(u0r, wd, bd)
    loss = ((x - pred) ** 2).mean()
    loss.backward()
    du0_ref = u0r.grad * (u0.float() > 0)
    dev = _dev()
    _, u0v = _slice_buf(B, H, H, Cu, 0, 64, dev)
    u0v.copy_(u0)
    du0 = torch.full((B, H, H, Cu), 7.0, dtype=HALF, device=dev)
    predg = torch.empty(B, H, H, 3, device=dev)
    lossg = torch.full((1,), 5.0, device=dev)
    dwd = torch.full((Cu + 3, 3), 3.0, device=dev)   # three guard rows: the kernel must not touch them
    dbd = torch.full((3,), 3.0, device=dev)
    ops.dense_mse(u0v, None, x.to(dev), wd.detach().to(dev), bd.detach().to(dev), lossg, 1.0 / (B * H * H * 3),
                  pred=predg, du0=du0, dwd=dwd, dbd=dbd)
    torch.cuda.synchronize()
    ms = [_metrics("dense pred", predg, pred, 2e-5), _metrics("mse loss", lossg, loss.reshape(1), 2e-5),
          _metrics("dense du0", du0, du0_ref, BF16_TOL), _metrics("dense dW", dwd[:Cu], wd.grad, F32_TOL),
          _metrics("dense db", dbd, bd.grad, F32_TOL)]
    worst = dict(max(ms, key=lambda m: m["err"] / m["tol"]))
    worst["name"] = f"dense_mse without image channels B{B} H{H} Cu{Cu} (worst: {worst['name']})"
    worst["pad_intact"] = bool((dwd[Cu:] == 3.0).all().item())
    return worst


# ---------------------------------------------------------------------------------------------- conv (down)
def check_conv_fprop(B, H, Cin, Cout, seed=0, pad=64, weights_stable=False):
    ops = _ops()
    g = torch.Generator().manual_seed(seed)
    x = _bf(_rand((B, H, H, Cin), g))
    w = _bf(_rand((4, 4, Cin, Cout), g, 1.0 / math.sqrt(16 * Cin)))
    b = _rand((Cout,), g, 0.1)
    ref = O.down_shuffle(x.float(), w.float(), b)
    dev = _dev()
    xfull, xv = _slice_buf(B, H, H, Cin, pad, 0, dev)
    xv.copy_(x)
    yfull, yv = _slice_buf(B, H // 2, H // 2, Cout, 0, pad, dev)
    ws = ops.Workspace(64 << 20, dev)
    ops.conv4s2_fprop(xv, w.to(dev), b.to(dev), yv, ws, weights_stable)
    torch.cuda.synchronize()
    m = _metrics(f"conv4s2_fprop B{B} H{H} {Cin}->{Cout}", yv, ref, BF16_TOL)
    m["pad_intact"] = bool((yfull[..., Cout:] == 7.0).all().item()) if pad else True
    return m


def check_conv_dgrad(B, H, Cin, Cout, add_old=True, seed=1, pad=64, weights_stable=False):
    ops = _ops()
    g = torch.Generator().manual_seed(seed)
    dy = _bf(_rand((B, H // 2, H // 2, Cout), g))
    w = _bf(_rand((4, 4, Cin, Cout), g, 1.0 / math.sqrt(16 * Cin)))
    act = _bf(_rand((B, H, H, Cin), g).clamp_min(0))
    old = _bf(_rand((B, H, H, Cin), g))
    xr = torch.zeros(B, H, H, Cin, requires_grad=True)
    y = F.conv2d(xr.permute(0, 3, 1, 2), w.float().permute(3, 2, 0, 1), None, stride=2, padding=1)
    y.backward(dy.float().permute(0, 3, 1, 2))
    ref = xr.grad
    if add_old:
        ref = ref + old.float()
    ref = ref * (act.float() > 0)
    dev = _dev()
    _, dyv = _slice_buf(B, H // 2, H // 2, Cout, 0, pad, dev)
    dyv.copy_(dy)
    dxfull, dxv = _slice_buf(B, H, H, Cin, pad, 0, dev)
    dxv.copy_(old)
    _, actv = _slice_buf(B, H, H, Cin, pad, 0, dev)
    actv.copy_(act)
    ws = ops.Workspace(64 << 20, dev)
    ops.conv4s2_dgrad(dyv, w.to(dev), dxv, actv, add_old, ws, weights_stable)
    torch.cuda.synchronize()
    m = _metrics(f"conv4s2_dgrad B{B} H{H} {Cin}<-{Cout} add{int(add_old)}", dxv, ref, BF16_TOL)
    m["pad_intact"] = bool((dxfull[..., :pad] == 7.0).all().item()) if pad else True
    return m


def check_conv_wgrad(B, H, Cin, Cout, seed=2, pad=64):
    ops = _ops()
    g = torch.Generator().manual_seed(seed)
    x = _bf(_rand((B, H, H, Cin), g))
    dy = _bf(_rand((B, H // 2, H // 2, Cout), g))
    wr = torch.zeros(4, 4, Cin, Cout, requires_grad=True)
    y = F.conv2d(x.float().permute(0, 3, 1, 2), wr.permute(3, 2, 0, 1), None, stride=2, padding=1)
    y.backward(dy.float().permute(0, 3, 1, 2))
    ref = wr.grad
    dev = _dev()
    _, xv = _slice_buf(B, H, H, Cin, pad, 0, dev)
    xv.copy_(x)
    _, dyv = _slice_buf(B, H // 2, H // 2, Cout, 0, pad, dev)
    dyv.copy_(dy)
    dw = torch.full((4, 4, Cin, Cout), 3.0, dtype=torch.float32, device=dev)
    ops.conv4s2_wgrad(xv, dyv, dw, ops.Workspace(64 << 20, dev))
    torch.cuda.synchronize()
    return _metrics(f"conv4s2_wgrad B{B} H{H} {Cin}x{Cout}", dw, ref, F32_TOL)


# ---------------------------------------------------------------------------------------------- convT (up)
def check_convT_fprop(B, H, Cin, Cout, seed=3, pad=64, weights_stable=False):
    ops = _ops()
    g = torch.Generator().manual_seed(seed)
    x = _bf(_rand((B, H, H, Cin), g))
    w = _bf(_rand((4, 4, Cout, Cin), g, 1.0 / math.sqrt(4 * Cin)))
    b = _rand((Cout,), g, 0.1)
    ref = O.up_shuffle(x.float(), w.float(), b)
    dev = _dev()
    _, xv = _slice_buf(B, H, H, Cin, 0, pad, dev)
    xv.copy_(x)
    yfull, yv = _slice_buf(B, 2 * H, 2 * H, Cout, 0, pad, dev)
    ws = ops.Workspace(64 << 20, dev)
    ops.convT4s2_fprop(xv, w.to(dev), b.to(dev), yv, ws, weights_stable)
    torch.cuda.synchronize()
    m = _metrics(f"convT4s2_fprop B{B} H{H} {Cin}->{Cout}", yv, ref, BF16_TOL)
    m["pad_intact"] = bool((yfull[..., Cout:] == 7.0).all().item()) if pad else True
    return m


def check_convT_dgrad(B, H, Cin, Cout, mask_channels=None, seed=4, pad=64, weights_stable=False):
    ops = _ops()
    g = torch.Generator().manual_seed(seed)
    if mask_channels is None:
        mask_channels = Cin // 2
    dy = _bf(_rand((B, 2 * H, 2 * H, Cout), g))
    w = _bf(_rand((4, 4, Cout, Cin), g, 1.0 / math.sqrt(16 * Cout)))
    act = _bf(_rand((B, H, H, Cin), g).clamp_min(0))
    xr = torch.zeros(B, H, H, Cin, requires_grad=True)
    y = F.conv_transpose2d(xr.permute(0, 3, 1, 2), w.float().permute(3, 2, 0, 1), None, stride=2, padding=1)
    y.backward(dy.float().permute(0, 3, 1, 2))
    ref = xr.grad.clone()
    ref[..., :mask_channels] *= (act.float()[..., :mask_channels] > 0)
    dev = _dev()
    _, dyv = _slice_buf(B, 2 * H, 2 * H, Cout, 0, pad, dev)
    dyv.copy_(dy)
    dxfull, dxv = _slice_buf(B, H, H, Cin, 0, pad, dev)
    _, actv = _slice_buf(B, H, H, Cin, 0, pad, dev)
    actv.copy_(act)
    ws = ops.Workspace(64 << 20, dev)
    ops.convT4s2_dgrad(dyv, w.to(dev), dxv, actv, mask_channels, ws, weights_stable)
    torch.cuda.synchronize()
    m = _metrics(f"convT4s2_dgrad B{B} H{H} {Cin}<-{Cout} mask{mask_channels}", dxv, ref, BF16_TOL)
    m["pad_intact"] = bool((dxfull[..., Cin:] == 7.0).all().item()) if pad else True
    return m


def check_convT_wgrad(B, H, Cin, Cout, seed=5, pad=64):
    ops = _ops()
    g = torch.Generator().manual_seed(seed)
    x = _bf(_rand((B, H, H, Cin), g))
    dy = _bf(_rand((B, 2 * H, 2 * H, Cout), g))
    wr = torch.zeros(4, 4, Cout, Cin, requires_grad=True)
    y = F.conv_transpose2d(x.float().permute(0, 3, 1, 2), wr.permute(3, 2, 0, 1), None, stride=2, padding=1)
    y.backward(dy.float().permute(0, 3, 1, 2))
    ref = wr.grad
    dev = _dev()
    _, xv = _slice_buf(B, H, H, Cin, 0, pad, dev)
    xv.copy_(x)
    _, dyv = _slice_buf(B, 2 * H, 2 * H, Cout, 0, pad, dev)
    dyv.copy_(dy)
    dw = torch.full((4, 4, Cout, Cin), 3.0, dtype=torch.float32, device=dev)
    ops.convT4s2_wgrad(xv, dyv, dw, ops.Workspace(64 << 20, dev))
    torch.cuda.synchronize()
    return _metrics(f"convT4s2_wgrad B{B} H{H} {Cout}x{Cin}", dw, ref, F32_TOL)


# ---------------------------------------------------------------------------------------------- HBM-bound kernels
def check_noise(B=3, H=32, seed=6):
    ops = _ops()
    cfg = O.Config(size=H)
    x, t_int, eps = O.synthetic_batch(cfg, B, seed)
    ref = O.noise_images(x, t_int, eps, cfg)
    dev = _dev()
    out = torch.empty_like(x, device=dev)
    ops.noise_images(x.to(dev), eps.to(dev), t_int.to(dev), out, cfg.steps)
    torch.cuda.synchronize()
    return _metrics(f"noise_images B{B} H{H}", out, ref, 2e-6)


def check_step_begin(B=4, H=64, seed=13):
    """Fused step prologue: the device-side draws have the right distributions (t_int ~ U{1..steps}, eps ~ N(0,1)),
    noised follows train.py:231-234 for the draws the kernel reports, the small-gradient region and the loss are zeroed,
    alpha matches the oracle's WarmUp/Adam arithmetic, draws differ between iterations and repeat for equal (seed, iteration)."""
    ops = _ops()
    cfg = O.Config(size=H)
    x, _, _ = O.synthetic_batch(cfg, B, seed)
    dev = _dev()
    xd = x.to(dev)
    outs = []
    for it in (0, 0, 5):
        noised = torch.empty_like(xd)
        eps = torch.empty_like(xd)
        t = torch.zeros(B, dtype=torch.int32, device=dev)
        iters = torch.full((1,), it, dtype=torch.int64, device=dev)
        hyper = torch.zeros(2, device=dev)
        gsmall = torch.full((4096,), 3.0, device=dev)
        loss = torch.full((1,), 3.0, device=dev)
        ops.step_begin(xd, noised, iters, hyper, gsmall, loss, 1234, cfg.steps, cfg.base_lr, cfg.warm_up, cfg.beta1,
                       cfg.beta2, eps_out=eps, t_out=t)
        torch.cuda.synchronize()
        outs.append((noised.cpu(), eps.cpu(), t.cpu(), hyper.cpu(), gsmall.cpu(), loss.cpu(), int(iters.item())))
    noised, eps, t, hyper, gsmall, loss, it_after = outs[2]
    ref = O.noise_images(x, t, eps, cfg)
    m = _metrics(f"step_begin B{B} H{H} noising", noised, ref, 2e-6)
    n = eps.numel()
    w = torch.zeros(1)
    m_, v_ = torch.zeros(1), torch.zeros(1)
    alpha = O.keras_adam_update(w, m_, v_, torch.zeros(1), 5, cfg)
    # many-draw statistics on a second, larger call
    big_x = torch.zeros(64, 64, 64, 3, device=dev)
    big_eps = torch.empty_like(big_x)
    big_t = torch.zeros(64, dtype=torch.int32, device=dev)
    ops.step_begin(big_x, torch.empty_like(big_x), torch.zeros(1, dtype=torch.int64, device=dev), torch.zeros(2, device=dev),
                   torch.zeros(4, device=dev), torch.zeros(1, device=dev), 99, 200, 2e-5, 2000, eps_out=big_eps, t_out=big_t)
    e = big_eps.double().cpu().flatten()
    ok = (abs(float(e.mean())) < 5e-3 and abs(float(e.var()) - 1) < 1e-2 and abs(float((e ** 4).mean()) - 3) < 0.1
          and float(e.abs().max()) > 4.0 and int(big_t.min()) >= 1 and int(big_t.max()) <= 200
          and len(set(big_t.cpu().tolist())) > 30
          and int(t.min()) >= 1 and int(t.max()) <= cfg.steps
          and bool((gsmall == 0).all()) and float(loss) == 0.0 and it_after == 5
          and abs(float(hyper[0]) - alpha) <= 1e-6 * alpha
          and torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])   # reproducible
          and not torch.equal(outs[0][1], outs[2][1]))                                       # fresh per iteration
    if not ok:
        m["err"] = float("inf")
        m["detail"] = dict(mean=float(e.mean()), var=float(e.var()), kurt=float((e ** 4).mean()), tmin=int(big_t.min()),
                           tmax=int(big_t.max()), alpha=float(hyper[0]), alpha_ref=alpha, it_after=it_after)
    return m


def check_step_begin_u8(B=3, H=32, seed=15):
    """The step prologue fed with decode_file's bytes (train.py:285-293): the decoded image (with the per-image
    left-right flip) must equal the oracle's decode bit for bit (u8/128 - 1 is exact in fp32), noised must follow
    train.py:231-234 for the draws the kernel reports, and the draws must equal those of the fp32 entry point."""
    ops = _ops()
    cfg = O.Config(size=H)
    g = torch.Generator().manual_seed(seed)
    img = torch.randint(0, 256, (B, H, H, 3), generator=g, dtype=torch.uint8)
    flip = torch.tensor([1, 0, 1][:B] + [0] * max(0, B - 3), dtype=torch.uint8)
    dev = _dev()
    worst = None
    for fl in (flip, None):
        ref_x = O.decode_u8(img, fl)
        x_out = torch.full((B, H, H, 3), 9.0, device=dev)
        noised = torch.empty_like(x_out)
        eps = torch.empty_like(x_out)
        t = torch.zeros(B, dtype=torch.int32, device=dev)
        iters = torch.full((1,), 7, dtype=torch.int64, device=dev)
        ops.step_begin_u8(img.to(dev), None if fl is None else fl.to(dev), x_out, noised, iters, torch.zeros(2, device=dev),
                          torch.zeros(8, device=dev), torch.zeros(1, device=dev), 1234, cfg.steps, cfg.base_lr, cfg.warm_up,
                          cfg.beta1, cfg.beta2, eps_out=eps, t_out=t)
        # same seed and iteration through the fp32 entry point: identical draws, identical noised
        noised32 = torch.empty_like(x_out)
        eps32 = torch.empty_like(x_out)
        t32 = torch.zeros(B, dtype=torch.int32, device=dev)
        ops.step_begin(ref_x.to(dev), noised32, iters, torch.zeros(2, device=dev), torch.zeros(8, device=dev),
                       torch.zeros(1, device=dev), 1234, cfg.steps, cfg.base_lr, cfg.warm_up, cfg.beta1, cfg.beta2,
                       eps_out=eps32, t_out=t32)
        torch.cuda.synchronize()
        m = _metrics(f"step_begin_u8 B{B} H{H} flip={'yes' if fl is not None else 'no'} noising", noised,
                     O.noise_images(ref_x, t.cpu(), eps.cpu(), cfg), 2e-6)
        exact = (torch.equal(x_out.cpu(), ref_x) and torch.equal(eps.cpu(), eps32.cpu()) and torch.equal(t.cpu(), t32.cpu())
                 and torch.equal(noised.cpu(), noised32.cpu()))
        if not exact:
            m["err"] = float("inf")
            m["detail"] = "decoded image / draws differ from the oracle decode or from the fp32 entry point"
        if worst is None or m["err"] / m["tol"] > worst["err"] / worst["tol"]:
            worst = m
    return worst


def check_sample_update_modes(B=2, H=16, seed=17, target=None):
    """gct2_sample_update for the objective switches of train.py:29-32 against oracle.sample_update (train.py:382-413)."""
    import dataclasses
    ops = _ops()
    cfg = dataclasses.replace(O.Config(size=H), **(target or {}))
    mode = ops.target_mode(cfg.predict_x, cfg.predict_scaled_epsilon, cfg.prediction_weighting,
                           cfg.ordinary_differential_equation)
    g = torch.Generator().manual_seed(seed)
    dev = _dev()
    shape = (B, H, H, 3)
    x0, e0, pred = _rand(shape, g), _rand(shape, g), _rand(shape, g)
    t, tn = 37, 36
    a, an = O.alpha_dash(float(t), cfg.steps), O.alpha_dash(float(tn), cfg.steps)
    fake_ref = a ** 0.5 * x0 + (1 - a) ** 0.5 * e0
    x_ref, e_ref = O.sample_update(pred, fake_ref, x0, e0, t, cfg)
    fake = torch.zeros(shape, device=dev)
    xt, et = x0.to(dev), e0.to(dev)
    ops.sample_update(None, fake, xt, et, t, t, cfg.steps, mode)
    ops.sample_update(pred.to(dev), fake, xt, et, t, tn, cfg.steps, mode)
    torch.cuda.synchronize()
    tol = 3e-5 if cfg.ordinary_differential_equation else 5e-6   # the ODE form divides by a difference of close terms
    ms = [_metrics("x_theta", xt, x_ref, tol), _metrics("eps_theta", et, e_ref, tol),
          _metrics("next fake", fake, an ** 0.5 * x_ref + (1 - an) ** 0.5 * e_ref, tol)]
    worst = dict(max(ms, key=lambda q: q["err"] / q["tol"]))
    worst["name"] = f"sample_update mode {mode} (worst: {worst['name']})"
    return worst


def check_latent_edits(S=32, K=8, seed=18):
    """gct2_latent_edits against oracle.latent_edits (train.py:418-432): the copies, the roll and the dictionary gather
    must be exact, the 4x4 average within fp32 summation order."""
    ops = _ops()
    g = torch.Generator().manual_seed(seed)
    e = _rand((1, S, S, 3), g)
    d = _rand((S, S, K, 3), g)
    ref = O.latent_edits(e, d)
    dev = _dev()
    out = torch.full((4, S, S, 3), 9.0, device=dev)
    ops.latent_edits(e.to(dev), d.to(dev), out)
    torch.cuda.synchronize()
    got = out.cpu()
    m = _metrics(f"latent_edits S{S} K{K} pixelated", got[1], ref[1], 1e-6)
    if not (torch.equal(got[0], ref[0]) and torch.equal(got[2], ref[2]) and torch.equal(got[3], ref[3])):
        m["err"] = float("inf")
        m["detail"] = "copy / roll / quantised planes differ"
    return m


def check_rmse(n=3 * 64 * 64, seed=19):
    ops = _ops()
    g = torch.Generator().manual_seed(seed)
    a, b = _rand((n,), g), _rand((n,), g)
    dev = _dev()
    out = torch.zeros(1, device=dev)
    ops.rmse(a.to(dev), b.to(dev), out)
    torch.cuda.synchronize()
    return _metrics(f"rmse n{n}", out, (((a - b) ** 2).mean() ** 0.5).reshape(1), 2e-6)


def check_sample_update(B=2, H=16, seed=16):
    """gct2_sample_update against the loop arithmetic of train.py:369-372 / :394-397 (fp32 elementwise)."""
    ops = _ops()
    cfg = O.Config(size=H)
    g = torch.Generator().manual_seed(seed)
    dev = _dev()
    shape = (B, H, H, 3)
    x0, e0, pred = _rand(shape, g), _rand(shape, g), _rand(shape, g)
    t, tn = 7, 8
    a, an = O.alpha_dash(float(t), cfg.steps), O.alpha_dash(float(tn), cfg.steps)
    fake_ref = a ** 0.5 * x0 + (1 - a) ** 0.5 * e0
    fake = torch.zeros(shape, device=dev)
    xt, et = x0.to(dev), e0.to(dev)
    ops.sample_update(None, fake, xt, et, t, t, cfg.steps)            # first mix
    m1 = _metrics("sample_update mix", fake, fake_ref, 2e-6)
    ops.sample_update(pred.to(dev), fake, xt, et, t, tn, cfg.steps)   # update + next mix
    e_ref = (fake_ref - a ** 0.5 * pred) / (1 - a) ** 0.5
    m2 = _metrics("sample_update eps_theta", et, e_ref, 2e-6)
    m3 = _metrics("sample_update next fake", fake, an ** 0.5 * pred + (1 - an) ** 0.5 * e_ref, 2e-6)
    keep = fake.clone()
    ops.sample_update(pred.to(dev), fake, xt, et, tn, 0, cfg.steps)   # last step: fake untouched
    torch.cuda.synchronize()
    worst = dict(max([m1, m2, m3], key=lambda q: q["err"] / q["tol"]))
    if not (torch.equal(xt.cpu(), pred) and torch.equal(keep, fake)):
        worst["err"] = float("inf")
    worst["name"] = f"sample_update B{B} H{H} (worst: {worst['name']})"
    return worst


def check_c3_fprop(B=2, H=32, Cout=128, seed=7):
    ops = _ops()
    g = torch.Generator().manual_seed(seed)
    x = _rand((B, H, H, 3), g)
    w = _rand((4, 4, 3, Cout), g, 0.2)
    b = _rand((Cout,), g, 0.1)
    ref = O.down_shuffle(x, w, b)
    dev = _dev()
    yfull, yv = _slice_buf(B, H // 2, H // 2, Cout, 64, 0, dev)
    ops.conv4s2_c3_fprop(x.to(dev), w.to(dev), b.to(dev), yv)
    torch.cuda.synchronize()
    m = _metrics(f"conv4s2_c3_fprop B{B} H{H} 3->{Cout}", yv, ref, BF16_TOL)
    m["pad_intact"] = bool((yfull[..., :64] == 7.0).all().item())
    return m


def check_c3_wgrad(B=2, H=32, Cout=128, seed=8):
    ops = _ops()
    g = torch.Generator().manual_seed(seed)
    x = _rand((B, H, H, 3), g)
    dz = _bf(_rand((B, H // 2, H // 2, Cout), g))
    wr = torch.zeros(4, 4, 3, Cout, requires_grad=True)
    y = F.conv2d(x.permute(0, 3, 1, 2), wr.permute(3, 2, 0, 1), None, stride=2, padding=1)
    y.backward(dz.float().permute(0, 3, 1, 2))
    dev = _dev()
    _, dzv = _slice_buf(B, H // 2, H // 2, Cout, 64, 0, dev)
    dzv.copy_(dz)
    dw = torch.full((4, 4, 3, Cout), 3.0, device=dev)
    db = torch.full((Cout,), 3.0, device=dev)
    ops.conv4s2_c3_wgrad(x.to(dev), dzv, dw, db)
    torch.cuda.synchronize()
    m = _metrics(f"conv4s2_c3_wgrad B{B} H{H}", dw, wr.grad, F32_TOL)
    m2 = _metrics("c3 bias grad", db, dz.float().sum((0, 1, 2)), F32_TOL)
    m["err"] = max(m["err"], m2["err"])
    return m


def check_bias_grad(B=2, H=16, C=256, seed=9):
    ops = _ops()
    g = torch.Generator().manual_seed(seed)
    dz = _bf(_rand((B, H, H, C), g))
    dev = _dev()
    _, dzv = _slice_buf(B, H, H, C, 0, 64, dev)
    dzv.copy_(dz)
    db = torch.full((C,), 3.0, device=dev)
    ops.bias_grad(dzv, db)
    torch.cuda.synchronize()
    return _metrics(f"bias_grad B{B} H{H} C{C}", db, dz.float().sum((0, 1, 2)), F32_TOL)


def check_bias_grad_multi(seed=12):
    """All layers' BiasAddGrad in one launch, ragged shapes, accumulate mode on top of a pre-filled output."""
    ops = _ops()
    g = torch.Generator().manual_seed(seed)
    dev = _dev()
    shapes = [(2, 16, 16, 64), (1, 4, 4, 512), (3, 8, 8, 256), (1, 32, 32, 128), (2, 4, 4, 1024)]
    dzs, dbs, refs = [], [], []
    for i, sh in enumerate(shapes):
        dz = _bf(_rand(sh, g))
        _, v = _slice_buf(*sh, 64 * (i % 2), 64, dev)
        v.copy_(dz)
        dzs.append(v)
        dbs.append(torch.full((sh[3],), 0.5, device=dev))
        refs.append(dz.float().sum((0, 1, 2)) + 0.5)
    plan = ops.BiasGradPlan(dzs, dbs)
    ops.bias_grad_multi(plan, accumulate=True)
    torch.cuda.synchronize()
    ms = [_metrics(f"bias_grad_multi seg{i}", d, r, F32_TOL) for i, (d, r) in enumerate(zip(dbs, refs))]
    worst = dict(max(ms, key=lambda m: m["err"]))
    worst["name"] = f"bias_grad_multi 5 tensors (worst: {worst['name']})"
    return worst


def check_dense_mse(B=2, H=32, Cu=64, seed=10, target=None):
    """target: None (predict_x) or the objective switches of train.py:29-32 as a dict for oracle.Config."""
    import dataclasses
    ops = _ops()
    g = torch.Generator().manual_seed(seed)
    u0 = _bf(_rand((B, H, H, Cu), g).clamp_min(0))
    noised = _rand((B, H, H, 3), g)
    x = _rand((B, H, H, 3), g)
    eps = _rand((B, H, H, 3), g)
    t_int = torch.randint(1, 201, (B,), generator=g, dtype=torch.int32)
    ocfg = dataclasses.replace(O.Config(size=H), **(target or {}))
    mode = ops.target_mode(ocfg.predict_x, ocfg.predict_scaled_epsilon, ocfg.prediction_weighting,
                           ocfg.ordinary_differential_equation)
    wd = _rand((Cu + 3, 3), g, 0.2).requires_grad_(True)
    bd = _rand((3,), g, 0.1).requires_grad_(True)
    u0r = u0.float().requires_grad_(True)
    pred = O.dense(torch.cat([u0r, noised], -1), wd, bd)
    tgt, wpred = O.loss_target(x, t_int, eps, pred, ocfg)
    loss = ((tgt - wpred) ** 2).mean()
    loss.backward()
    du0_ref = u0r.grad * (u0.float() > 0)
    dev = _dev()
    _, u0v = _slice_buf(B, H, H, Cu, 0, 64, dev)
    u0v.copy_(u0)
    du0 = torch.full((B, H, H, Cu), 7.0, dtype=HALF, device=dev)
    predg = torch.empty(B, H, H, 3, device=dev)
    lossg = torch.full((1,), 5.0, device=dev)
    dwd = torch.full((Cu + 3, 3), 3.0, device=dev)
    dbd = torch.full((3,), 3.0, device=dev)
    ops.dense_mse(u0v, noised.to(dev), x.to(dev), wd.detach().to(dev), bd.detach().to(dev), lossg,
                  1.0 / (B * H * H * 3), pred=predg, du0=du0, dwd=dwd, dbd=dbd, eps=eps.to(dev), t_int=t_int.to(dev),
                  mode=mode, steps=ocfg.steps)
    torch.cuda.synchronize()
    ms = [_metrics("dense pred", predg, pred, 2e-5), _metrics("mse loss", lossg, loss.reshape(1), 2e-5),
          _metrics("dense du0", du0, du0_ref, BF16_TOL), _metrics("dense dW", dwd, wd.grad, F32_TOL),
          _metrics("dense db", dbd, bd.grad, F32_TOL)]
    worst = max(ms, key=lambda m: m["err"] / m["tol"])
    worst = dict(worst)
    worst["name"] = f"dense_mse B{B} H{H} mode {mode} (worst: {worst['name']})"
    return worst


def check_adam(n=4096 + 128, steps=3, seed=11):
    ops = _ops()
    g = torch.Generator().manual_seed(seed)
    cfg = O.Config(warm_up=2)
    w = _rand((n,), g)
    # Keras-epsilon-sensitive magnitudes: most gradients of this model are below 1e-6
    gr = [_rand((n,), g) * (10.0 ** torch.randint(-9, -2, (n,), generator=g).float()) for _ in range(steps)]
    wo, m, v = w.clone(), torch.zeros(n), torch.zeros(n)
    dev = _dev()
    wg, mg, vg = w.clone().to(dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    wb = torch.zeros(n, dtype=HALF, device=dev)
    it = torch.zeros(1, dtype=torch.int64, device=dev)
    hyper = torch.zeros(2, device=dev)
    for s in range(steps):
        O.keras_adam_update(wo, m, v, gr[s], s, cfg)
        ops.adam_keras(wg, mg, vg, gr[s].to(dev), wb, it, hyper, cfg.base_lr, cfg.warm_up, cfg.beta1, cfg.beta2,
                       cfg.epsilon, 1.0)
    torch.cuda.synchronize()
    # w ~ 1 and |delta| ~ 1e-5: the fp32 ulp of w (1.2e-7) caps how well the *delta* can agree (two fp32 pipelines
    # that differ by one ulp in sqrt/div round w differently) -> compare w tightly and the delta loosely.
    md = _metrics("adam delta-w", wg.cpu() - w, wo - w, 5e-3)
    mw = _metrics("adam w", wg, wo, 2e-7)
    mm = _metrics("adam m", mg, m, 1e-5)
    mv = _metrics("adam v", vg, v, 1e-5)
    mb = _metrics("adam bf16 shadow", wb, wo, BF16_TOL)
    worst = dict(max([md, mw, mm, mv, mb], key=lambda q: q["err"] / q["tol"]))
    worst["name"] = f"adam_keras n{n} steps{steps} (worst: {worst['name']})"
    worst["iterations"] = int(it.item())
    return worst


# (function, kwargs) -- sized so the CPU references finish in seconds.
CONV_CASES = [
    (check_conv_fprop, dict(B=2, H=16, Cin=64, Cout=64)),
    (check_conv_fprop, dict(B=1, H=32, Cin=128, Cout=256)),
    (check_conv_fprop, dict(B=3, H=8, Cin=256, Cout=128)),
    (check_convT_fprop, dict(B=2, H=8, Cin=64, Cout=64)),
    (check_convT_fprop, dict(B=1, H=16, Cin=256, Cout=128)),
    (check_convT_fprop, dict(B=3, H=4, Cin=128, Cout=256)),
    (check_conv_dgrad, dict(B=2, H=16, Cin=64, Cout=64, add_old=True)),
    (check_conv_dgrad, dict(B=1, H=32, Cin=128, Cout=256, add_old=False)),
    (check_conv_dgrad, dict(B=3, H=8, Cin=256, Cout=128, add_old=True)),
    (check_convT_dgrad, dict(B=2, H=8, Cin=128, Cout=64)),
    (check_convT_dgrad, dict(B=1, H=16, Cin=256, Cout=128, mask_channels=256)),
    (check_convT_dgrad, dict(B=3, H=4, Cin=128, Cout=256, mask_channels=0)),
    (check_conv_wgrad, dict(B=2, H=16, Cin=128, Cout=64)),
    (check_conv_wgrad, dict(B=1, H=32, Cin=64, Cout=128)),
    (check_conv_wgrad, dict(B=3, H=8, Cin=256, Cout=256)),
    (check_convT_wgrad, dict(B=2, H=8, Cin=128, Cout=64)),
    (check_convT_wgrad, dict(B=1, H=16, Cin=64, Cout=128)),
    (check_convT_wgrad, dict(B=3, H=4, Cin=256, Cout=256)),
]
EW_CASES = [
    (check_noise, {}),
    (check_step_begin, {}),
    (check_step_begin_u8, {}),
    (check_sample_update, {}),
    (check_c3_fprop, {}),
    (check_c3_wgrad, {}),
    (check_bias_grad, {}),
    (check_bias_grad, dict(B=1, H=8, C=1024)),
    (check_bias_grad_multi, {}),
    (check_dense_mse, {}),
    (check_dense_mse, dict(B=1, H=16, Cu=128)),
    (check_dense_mse, dict(B=3, H=5, Cu=64)),
    (check_adam, {}),
    # the objective switches of train.py:29-32 (loss targets, train.py:238-252)
    (check_dense_mse, dict(target=dict(predict_x=False))),
    (check_dense_mse, dict(target=dict(predict_x=False, predict_scaled_epsilon=True))),
    (check_dense_mse, dict(target=dict(predict_x=False, prediction_weighting=True))),
    (check_dense_mse, dict(target=dict(predict_x=False, predict_scaled_epsilon=True, prediction_weighting=True), B=3, H=8)),
    (check_dense_mse, dict(target=dict(ordinary_differential_equation=True))),
    (check_sample_update_modes, dict(target=dict(predict_x=False))),
    (check_sample_update_modes, dict(target=dict(predict_x=False, predict_scaled_epsilon=True))),
    (check_sample_update_modes, dict(target=dict(ordinary_differential_equation=True))),
    (check_sample_update_modes, {}),
    # the rest of log_sample (train.py:325-361, 418-432)
    (check_latent_edits, {}),
    (check_latent_edits, dict(S=64, K=3)),
    (check_rmse, {}),
]


def forced(fn, BN=0, splits=0, pair=0, nofuse=0, budget=0, finish=None, **kw):
    """Runs a conv check with the tile width / split-K factor / CTA-pair mode pinned (test hooks gct2_debug_set keys 3,
    4, 19, 12 and gct2_set_sm_budget) so that every template instantiation and every split-K finishing path is exercised
    regardless of the heuristics; then asks the library which plan the launch actually used (gct2_debug_last_plan) --
    a case that asked for CTA pairs / a fused finish and silently got something else fails."""
    from gan_class_transfer2_b200 import _lib, ops
    lib = _lib.init(0)
    # how split-K is finished: "cluster" = the splits are one thread-block cluster (partials through distributed shared
    # memory), "l2" = in-launch rendezvous over fp32 slabs in global memory, "kernel" = separate finishing kernel
    cs = {None: 1, "l2": 1, "cluster": 2, "kernel": 1}[finish]
    if finish == "kernel":
        nofuse = 1
    for key, val in ((3, BN), (4, splits), (19, pair), (12, nofuse), (22, budget), (25, cs)):
        lib.gct2_debug_set(key, val)
    try:
        m = fn(**kw)
        plan = ops.last_plan()
    finally:
        for key in (3, 4, 19, 12, 22):
            lib.gct2_debug_set(key, 0)
        lib.gct2_debug_set(25, 1)  # the library's default
    m["name"] += (f" [BN={BN or 'auto'} splits={splits or 'auto'} pair={pair} finish={finish or ('kernel' if nofuse else 'auto')}"
                  f"{f' budget={budget}' if budget else ''} -> {plan}]")
    want = {}
    if BN:
        want["BN"] = BN
    if splits:
        want["splits"] = splits
    if pair == 1:
        want["pair"] = 1
    if pair == 2:
        want["pair"] = 0
    if nofuse:
        want["fused"] = 0
    if finish == "l2":
        want["fused"] = 1
    if finish == "cluster":
        want["fused"] = 2
    if budget:
        assert plan["grid"] <= budget, (plan, budget)
    wrong = {k: (plan[k], v) for k, v in want.items() if plan[k] != v}
    if wrong:
        m["err"] = float("inf")
        m["detail"] = f"plan differs from the forced one: {wrong}"
    m["plan"] = plan
    return m


# (check, shape kwargs, forced plan)
FORCED_CASES = [
    (check_conv_fprop, dict(B=2, H=16, Cin=128, Cout=256), dict(BN=64, splits=1)),
    (check_conv_fprop, dict(B=2, H=16, Cin=128, Cout=256), dict(BN=128, splits=4, finish="l2")),
    (check_conv_fprop, dict(B=2, H=16, Cin=128, Cout=256), dict(BN=256, splits=16, finish="l2")),
    (check_convT_fprop, dict(B=2, H=8, Cin=128, Cout=256), dict(BN=64, splits=2, finish="l2")),
    (check_convT_fprop, dict(B=2, H=8, Cin=128, Cout=256), dict(BN=128, splits=1)),
    (check_convT_fprop, dict(B=2, H=8, Cin=128, Cout=256), dict(BN=256, splits=8, finish="l2")),
    (check_conv_dgrad, dict(B=2, H=16, Cin=256, Cout=128, add_old=True), dict(BN=256, splits=4, finish="l2")),
    (check_conv_dgrad, dict(B=2, H=16, Cin=256, Cout=128, add_old=True), dict(BN=128, splits=1)),
    (check_convT_dgrad, dict(B=2, H=8, Cin=256, Cout=128), dict(BN=256, splits=2, finish="l2")),
    (check_convT_dgrad, dict(B=2, H=8, Cin=256, Cout=128), dict(BN=64, splits=32, finish="l2")),
    # split-K inside a thread-block cluster (partials through distributed shared memory): every tile width x mode x
    # epilogue, cluster sizes 2 / 4 / 8, a ragged batch, more clusters than the chip holds at once (waves), early weights
    (check_conv_fprop, dict(B=2, H=16, Cin=128, Cout=256), dict(BN=64, splits=2, finish="cluster")),
    (check_conv_fprop, dict(B=2, H=16, Cin=128, Cout=256), dict(BN=128, splits=4, finish="cluster")),
    (check_conv_fprop, dict(B=2, H=16, Cin=128, Cout=256), dict(BN=256, splits=8, finish="cluster")),
    (check_convT_fprop, dict(B=2, H=8, Cin=128, Cout=256), dict(BN=64, splits=8, finish="cluster")),
    (check_convT_fprop, dict(B=2, H=8, Cin=128, Cout=256), dict(BN=128, splits=2, finish="cluster")),
    (check_convT_fprop, dict(B=2, H=8, Cin=128, Cout=256), dict(BN=256, splits=4, finish="cluster")),
    (check_conv_dgrad, dict(B=2, H=16, Cin=256, Cout=128, add_old=True), dict(BN=256, splits=4, finish="cluster")),
    (check_conv_dgrad, dict(B=2, H=16, Cin=256, Cout=128, add_old=False), dict(BN=64, splits=2, finish="cluster")),
    (check_convT_dgrad, dict(B=2, H=8, Cin=256, Cout=128), dict(BN=128, splits=8, finish="cluster")),
    (check_convT_dgrad, dict(B=2, H=8, Cin=256, Cout=128, mask_channels=64), dict(BN=64, splits=4, finish="cluster")),
    (check_conv_fprop, dict(B=3, H=8, Cin=256, Cout=128), dict(BN=64, splits=8, finish="cluster")),             # ragged
    (check_convT_dgrad, dict(B=3, H=4, Cin=128, Cout=256, mask_channels=64), dict(BN=128, splits=4, finish="cluster")),
    (check_conv_fprop, dict(B=16, H=32, Cin=128, Cout=256), dict(BN=64, splits=4, finish="cluster")),           # 512 CTAs: waves
    (check_convT_fprop, dict(B=8, H=16, Cin=128, Cout=128), dict(BN=64, splits=2, finish="cluster")),               # 256 CTAs
    (check_conv_fprop, dict(B=2, H=16, Cin=128, Cout=256, weights_stable=True), dict(BN=64, splits=4, finish="cluster")),
    (check_convT_fprop, dict(B=2, H=8, Cin=192, Cout=64), dict(BN=64, splits=4, finish="cluster")),             # 3 chunks per split
    (check_conv_wgrad, dict(B=4, H=32, Cin=128, Cout=256), dict(BN=64, splits=8)),
    (check_conv_wgrad, dict(B=4, H=32, Cin=128, Cout=256), dict(BN=128, splits=1)),
    (check_conv_wgrad, dict(B=4, H=32, Cin=128, Cout=256), dict(BN=256, splits=2)),
    (check_convT_wgrad, dict(B=4, H=16, Cin=256, Cout=64), dict(BN=64, splits=4)),
    (check_convT_wgrad, dict(B=4, H=16, Cin=256, Cout=128), dict(BN=256, splits=16)),
    # odd number of k-chunks per work item: the last ring round of a two-chunk slot is half full (BN <= 128)
    # four filter taps per work item when the gathered N side has 64 channels (up0's weight gradient): a 256-wide tile whose
    # 64-column blocks are four taps; forcing BN = 64 above keeps the one-tap path covered
    (check_convT_wgrad, dict(B=4, H=16, Cin=256, Cout=64), dict(BN=256, splits=4)),
    (check_convT_wgrad, dict(B=2, H=32, Cin=256, Cout=64), dict(BN=256, splits=1)),
    (check_convT_wgrad, dict(B=1, H=128, Cin=256, Cout=64), dict(BN=256, splits=16)),   # up0 itself at batch 1
    (check_conv_wgrad, dict(B=3, H=16, Cin=64, Cout=128), dict(BN=256, splits=1)),      # 3 pixel chunks
    (check_conv_wgrad, dict(B=3, H=16, Cin=128, Cout=64), dict(BN=64, splits=1)),       # 3 pixel chunks
    (check_convT_wgrad, dict(B=1, H=4, Cin=128, Cout=128), dict(BN=128, splits=1)),     # 1 chunk, a quarter full
    (check_conv_fprop, dict(B=1, H=16, Cin=64, Cout=128), dict(BN=128, splits=16, finish="l2")),     # 1 chunk per split
    (check_convT_fprop, dict(B=2, H=8, Cin=192, Cout=64), dict(BN=64, splits=4, finish="l2")),       # 3 chunks per split
    # split-K finished by the separate kernel (the fallback when a CTA owns more than one work item)
    (check_conv_fprop, dict(B=2, H=16, Cin=128, Cout=256), dict(BN=128, splits=4, nofuse=1)),
    (check_convT_fprop, dict(B=2, H=8, Cin=128, Cout=256), dict(BN=256, splits=8, nofuse=1)),
    (check_conv_dgrad, dict(B=2, H=16, Cin=256, Cout=128, add_old=True), dict(BN=256, splits=4, nofuse=1)),
    (check_convT_dgrad, dict(B=2, H=8, Cin=256, Cout=128), dict(BN=64, splits=32, nofuse=1)),
    # split-K with a ragged batch: rows of the tile beyond the batch are neither stored nor finished
    (check_conv_fprop, dict(B=3, H=8, Cin=256, Cout=128), dict(BN=64, splits=16, finish="l2")),
    (check_convT_dgrad, dict(B=3, H=4, Cin=128, Cout=256, mask_channels=64), dict(BN=128, splits=8, finish="l2")),
    # persistent CTAs: more work items than the SM budget (several tiles per CTA, TMEM double buffering, ring wrap)
    (check_conv_fprop, dict(B=16, H=64, Cin=64, Cout=128), dict(BN=64, splits=1, budget=24)),
    (check_convT_fprop, dict(B=16, H=16, Cin=64, Cout=64), dict(BN=64, splits=1, budget=20)),
    (check_conv_dgrad, dict(B=4, H=32, Cin=256, Cout=128, add_old=False), dict(BN=256, splits=1, budget=6)),
    (check_conv_wgrad, dict(B=4, H=32, Cin=128, Cout=256), dict(BN=128, splits=2, budget=10)),
    # weights fetched before the programmatic dependency resolves (GCT2_WEIGHTS_STABLE)
    (check_conv_fprop, dict(B=2, H=16, Cin=128, Cout=256, weights_stable=True), dict(BN=64, splits=1)),
    (check_conv_fprop, dict(B=2, H=16, Cin=128, Cout=256, weights_stable=True), dict(BN=128, splits=16, finish="l2")),
    (check_convT_fprop, dict(B=2, H=8, Cin=128, Cout=256, weights_stable=True), dict(BN=256, splits=1)),
    (check_convT_fprop, dict(B=16, H=16, Cin=64, Cout=64, weights_stable=True), dict(BN=64, splits=1, budget=20)),
    (check_conv_dgrad, dict(B=2, H=16, Cin=256, Cout=128, add_old=True, weights_stable=True), dict(BN=128, splits=2, finish="l2")),
    (check_convT_dgrad, dict(B=2, H=8, Cin=256, Cout=128, weights_stable=True), dict(BN=64, splits=4, finish="l2")),
]

# CTA pairs (tcgen05 cta_group::2): all six PAIR instantiations ({S, P, W} x BN {128, 256}), each with the direct
# epilogue, split-K finished inside the launch, split-K finished by the separate kernel, a ragged batch, persistent
# CTA pairs (more tile pairs than resident clusters) and the early weight fetch.  The heuristic picks pairs for every
# launch with more tiles than SMs (>= 8 images per GPU, BASELINE config 4), so these are the kernels behind those numbers.
PAIR_CASES = [
    # MODE_S (DownShuffle fprop / UpShuffle dgrad)
    (check_conv_fprop, dict(B=2, H=32, Cin=128, Cout=256), dict(BN=128, splits=1, pair=1)),
    (check_conv_fprop, dict(B=2, H=32, Cin=128, Cout=256), dict(BN=256, splits=1, pair=1)),
    (check_conv_fprop, dict(B=2, H=32, Cin=128, Cout=256), dict(BN=128, splits=4, pair=1)),
    (check_conv_fprop, dict(B=2, H=32, Cin=128, Cout=256), dict(BN=256, splits=8, pair=1)),
    (check_conv_fprop, dict(B=2, H=32, Cin=128, Cout=256), dict(BN=128, splits=4, pair=1, nofuse=1)),
    (check_conv_fprop, dict(B=2, H=32, Cin=128, Cout=256), dict(BN=256, splits=2, pair=1, nofuse=1)),
    (check_conv_fprop, dict(B=3, H=16, Cin=128, Cout=256), dict(BN=128, splits=1, pair=1)),            # ragged: 3 of 4 images
    (check_conv_fprop, dict(B=3, H=16, Cin=128, Cout=256), dict(BN=256, splits=4, pair=1)),
    (check_conv_fprop, dict(B=16, H=64, Cin=64, Cout=128), dict(BN=128, splits=1, pair=1, budget=12)),  # persistent pairs
    (check_conv_fprop, dict(B=2, H=32, Cin=128, Cout=256, weights_stable=True), dict(BN=128, splits=2, pair=1)),
    (check_convT_dgrad, dict(B=2, H=16, Cin=256, Cout=128), dict(BN=128, splits=1, pair=1)),
    (check_convT_dgrad, dict(B=2, H=16, Cin=256, Cout=128), dict(BN=256, splits=1, pair=1)),
    (check_convT_dgrad, dict(B=2, H=16, Cin=256, Cout=128, mask_channels=64), dict(BN=128, splits=4, pair=1)),
    (check_convT_dgrad, dict(B=2, H=16, Cin=256, Cout=128), dict(BN=256, splits=2, pair=1, nofuse=1)),
    (check_convT_dgrad, dict(B=3, H=8, Cin=256, Cout=128, weights_stable=True), dict(BN=256, splits=4, pair=1)),
    # MODE_P (UpShuffle fprop / DownShuffle dgrad)
    (check_convT_fprop, dict(B=2, H=16, Cin=128, Cout=256), dict(BN=128, splits=1, pair=1)),
    (check_convT_fprop, dict(B=2, H=16, Cin=128, Cout=256), dict(BN=256, splits=1, pair=1)),
    (check_convT_fprop, dict(B=2, H=16, Cin=128, Cout=256), dict(BN=128, splits=2, pair=1)),
    (check_convT_fprop, dict(B=2, H=16, Cin=128, Cout=256), dict(BN=256, splits=4, pair=1)),
    (check_convT_fprop, dict(B=2, H=16, Cin=128, Cout=256), dict(BN=256, splits=4, pair=1, nofuse=1)),
    (check_convT_fprop, dict(B=3, H=8, Cin=128, Cout=256), dict(BN=128, splits=1, pair=1)),
    (check_convT_fprop, dict(B=8, H=32, Cin=64, Cout=128), dict(BN=128, splits=1, pair=1, budget=16)),
    (check_convT_fprop, dict(B=2, H=16, Cin=128, Cout=256, weights_stable=True), dict(BN=256, splits=2, pair=1)),
    (check_conv_dgrad, dict(B=2, H=32, Cin=256, Cout=128, add_old=True), dict(BN=128, splits=1, pair=1)),
    (check_conv_dgrad, dict(B=2, H=32, Cin=256, Cout=128, add_old=True), dict(BN=256, splits=1, pair=1)),
    (check_conv_dgrad, dict(B=2, H=32, Cin=256, Cout=128, add_old=False), dict(BN=256, splits=2, pair=1)),
    (check_conv_dgrad, dict(B=2, H=32, Cin=256, Cout=128, add_old=True), dict(BN=128, splits=2, pair=1, nofuse=1)),
    (check_conv_dgrad, dict(B=3, H=16, Cin=256, Cout=128, add_old=True, weights_stable=True), dict(BN=128, splits=2, pair=1)),
    # MODE_W (both weight gradients; the gathered operand on the M side and on the N side)
    (check_conv_wgrad, dict(B=2, H=32, Cin=256, Cout=256), dict(BN=128, splits=1, pair=1)),
    (check_conv_wgrad, dict(B=2, H=32, Cin=256, Cout=256), dict(BN=256, splits=1, pair=1)),
    (check_conv_wgrad, dict(B=2, H=32, Cin=256, Cout=256), dict(BN=128, splits=4, pair=1)),
    (check_conv_wgrad, dict(B=3, H=16, Cin=256, Cout=256), dict(BN=256, splits=1, pair=1)),             # 3 chunks
    (check_conv_wgrad, dict(B=2, H=32, Cin=256, Cout=128), dict(BN=128, splits=2, pair=1, budget=10)),
    (check_convT_wgrad, dict(B=2, H=16, Cin=256, Cout=256), dict(BN=128, splits=1, pair=1)),
    (check_convT_wgrad, dict(B=2, H=16, Cin=256, Cout=256), dict(BN=256, splits=2, pair=1)),
    (check_convT_wgrad, dict(B=3, H=8, Cin=512, Cout=256), dict(BN=256, splits=1, pair=1)),             # 3 chunks, 4 M tiles
]


# stride-1 index maps (train.py:131-139): every tile width, split-K through L2 / the finishing kernel, 1x1, ragged batch,
# extents from 4x4 (tile spans several images) to 64x64 (several tiles per image), a concat-style partial mask
S1_CASES = [
    (check_conv3_fprop, dict(B=2, H=16, Cin=128, Cout=256), dict(BN=64, splits=1)),
    (check_conv3_fprop, dict(B=2, H=16, Cin=128, Cout=256), dict(BN=128, splits=2, finish="l2")),
    (check_conv3_fprop, dict(B=2, H=16, Cin=128, Cout=256), dict(BN=256, splits=1)),
    (check_conv3_fprop, dict(B=3, H=8, Cin=256, Cout=128), dict(BN=128, splits=4, finish="kernel")),
    (check_conv3_fprop, dict(B=3, H=4, Cin=512, Cout=512), dict()),
    (check_conv3_fprop, dict(B=1, H=64, Cin=192, Cout=128), dict()),
    (check_conv3_fprop, dict(B=2, H=32, Cin=64, Cout=64, weights_stable=True), dict()),
    (check_conv3_fprop, dict(B=2, H=16, Cin=128, Cout=64, ks=1), dict()),
    (check_conv3_dgrad, dict(B=2, H=16, Cin=256, Cout=128), dict(BN=64, splits=1)),
    (check_conv3_dgrad, dict(B=2, H=16, Cin=256, Cout=128, add_old=True), dict(BN=128, splits=2, finish="l2")),
    (check_conv3_dgrad, dict(B=2, H=16, Cin=256, Cout=128, mask_channels=128), dict(BN=256, splits=1)),
    (check_conv3_dgrad, dict(B=3, H=8, Cin=128, Cout=256, add_old=True), dict(BN=128, splits=4, finish="kernel")),
    (check_conv3_dgrad, dict(B=3, H=4, Cin=512, Cout=512), dict()),
    (check_conv3_dgrad, dict(B=1, H=64, Cin=192, Cout=128, mask_channels=64), dict()),
    (check_conv3_dgrad, dict(B=2, H=16, Cin=64, Cout=128, ks=1, weights_stable=True), dict()),
    (check_conv3_wgrad, dict(B=2, H=16, Cin=128, Cout=256), dict(BN=64, splits=1)),
    (check_conv3_wgrad, dict(B=2, H=16, Cin=128, Cout=256), dict(BN=128, splits=2)),
    (check_conv3_wgrad, dict(B=2, H=16, Cin=256, Cout=256), dict(BN=256, splits=1)),
    (check_conv3_wgrad, dict(B=3, H=8, Cin=64, Cout=128), dict()),
    (check_conv3_wgrad, dict(B=3, H=4, Cin=512, Cout=512), dict()),
    (check_conv3_wgrad, dict(B=1, H=64, Cin=192, Cout=128), dict()),
    (check_conv3_wgrad, dict(B=2, H=16, Cin=128, Cout=64, ks=1), dict()),
]

S1_EW_CASES = [
    (check_proj_add, dict(B=2, H=16, Cin=64, Cout=128)),
    (check_proj_add, dict(B=3, H=4, Cin=256, Cout=256)),
    (check_proj_add, dict(B=1, H=64, Cin=128, Cout=64)),
    (check_res0_fold, {}),
    (check_res0_fold, dict(U=128, B=1)),
    (check_conv3_c3, {}),
    (check_conv3_c3, dict(B=1, H=16, Cout=64)),
    (check_conv3_c3, dict(B=3, H=8, Cout=512)),
    (check_dense_mse_noimage, {}),
    (check_dense_mse_noimage, dict(B=1, H=16, Cu=64)),
]


def passed(m) -> bool:
    return m["err"] <= m["tol"] and m.get("pad_intact", True)
