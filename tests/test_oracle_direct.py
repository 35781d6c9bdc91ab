"""CPU: the PyTorch-CPU oracle (oracle/oracle.py) against an independent plain-C restatement written from the
TensorFlow definitions of the ops (oracle/direct.c: explicit loops, SAME padding computed from TF's rule,
Conv2DTranspose as the scatter form of Conv2D's input gradient).  Two restatements that share no code agreeing does
not pin either against TensorFlow (parity stays unpinned), but it does pin the layer semantics of SURVEY.md A.1-A.6
against a second, library-free derivation."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
ODIR = os.path.join(os.path.dirname(HERE), "oracle")
F, I, LL, D = ctypes.POINTER(ctypes.c_float), ctypes.c_int, ctypes.c_longlong, ctypes.c_double


@pytest.fixture(scope="module")
def lib():
    subprocess.run(["make", "-C", ODIR], check=True, capture_output=True)
    so = ctypes.CDLL(os.path.join(ODIR, "_direct.so"))
    so.gct2_direct_conv2d_same.argtypes = [F, F, F, F, I, I, I, I, I, I, I, I]
    so.gct2_direct_conv2d_transpose_same.argtypes = [F, F, F, F, I, I, I, I, I, I, I, I]
    so.gct2_direct_dense.argtypes = [F, F, F, F, LL, I, I]
    so.gct2_direct_mse.argtypes = [F, F, LL]
    so.gct2_direct_mse.restype = D
    so.gct2_direct_alpha_dash.argtypes = [D, I]
    so.gct2_direct_alpha_dash.restype = D
    so.gct2_direct_noise.argtypes = [F, F, ctypes.POINTER(ctypes.c_int), F, I, LL, I]
    so.gct2_direct_adam.argtypes = [F, F, F, F, LL, D, D, D, D, LL]
    return so


def _p(a):
    return a.ctypes.data_as(F)


def _np(t):
    return np.ascontiguousarray(t.detach().numpy().astype(np.float32))


def conv(lib, x, w, b, relu=True):
    B, H, W, Cin = x.shape
    Cout = w.shape[3]
    y = np.empty((B, (H + 1) // 2, (W + 1) // 2, Cout), np.float32)
    lib.gct2_direct_conv2d_same(_p(x), _p(w), _p(b), _p(y), B, H, W, Cin, Cout, 4, 2, int(relu))
    return y


def convT(lib, x, w, b, relu=True):
    B, H, W, Cin = x.shape
    Cout = w.shape[2]
    y = np.empty((B, 2 * H, 2 * W, Cout), np.float32)
    lib.gct2_direct_conv2d_transpose_same(_p(x), _p(w), _p(b), _p(y), B, H, W, Cin, Cout, 4, 2, int(relu))
    return y


@pytest.mark.parametrize("shape", [(2, 8, 8, 5, 7), (1, 4, 4, 3, 4), (1, 16, 8, 2, 3)])
def test_down_and_up_shuffle_match_the_direct_definition(lib, shape):
    B, H, W, Ci, Co = shape
    g = torch.Generator().manual_seed(H * 100 + Ci)
    x = torch.randn(B, H, W, Ci, generator=g)
    wd, bd = torch.randn(4, 4, Ci, Co, generator=g) * 0.3, torch.randn(Co, generator=g)
    wu, bu = torch.randn(4, 4, Co, Ci, generator=g) * 0.3, torch.randn(Co, generator=g)
    np.testing.assert_allclose(conv(lib, _np(x), _np(wd), _np(bd)), _np(O.down_shuffle(x, wd, bd)), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(convT(lib, _np(x), _np(wu), _np(bu)), _np(O.up_shuffle(x, wu, bu)), rtol=1e-5, atol=1e-5)


def test_dense_mse_noise_and_adam_match(lib):
    g = torch.Generator().manual_seed(4)
    cfg = O.Config(size=8)
    x = torch.randn(2, 8, 8, 67, generator=g)
    wk, bk = torch.randn(67, 3, generator=g), torch.randn(3, generator=g)
    y = np.empty((2, 8, 8, 3), np.float32)
    lib.gct2_direct_dense(_p(_np(x)), _p(_np(wk)), _p(_np(bk)), _p(y), 2 * 64, 67, 3)
    np.testing.assert_allclose(y, _np(O.dense(x, wk, bk)), rtol=1e-5, atol=1e-5)
    a, b = _np(torch.randn(1000, generator=g)), _np(torch.randn(1000, generator=g))
    assert abs(lib.gct2_direct_mse(_p(a), _p(b), 1000) - float(((torch.tensor(a) - torch.tensor(b)) ** 2).mean())) < 1e-6
    for t in (1, 57, 200):
        assert abs(lib.gct2_direct_alpha_dash(float(t), cfg.steps) - O.alpha_dash(float(t), cfg.steps)) < 1e-12
    img, t_int, eps = O.synthetic_batch(cfg, 3, 2)
    out = np.empty_like(_np(img))
    ti = np.ascontiguousarray(t_int.numpy().astype(np.int32))
    lib.gct2_direct_noise(_p(_np(img)), _p(_np(eps)), ti.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), _p(out), 3,
                          8 * 8 * 3, cfg.steps)
    np.testing.assert_allclose(out, _np(O.noise_images(img, t_int, eps, cfg)), rtol=1e-6, atol=1e-6)
    # three Keras-Adam steps with warm-up, gradients spanning the epsilon-sensitive magnitudes
    n = 512
    w0 = torch.randn(n, generator=g)
    wo, mo, vo = w0.clone(), torch.zeros(n), torch.zeros(n)
    wd_, md_, vd_ = _np(w0), np.zeros(n, np.float32), np.zeros(n, np.float32)
    warm = O.WarmUp(cfg.base_lr, cfg.warm_up)
    for step in range(3):
        gr = torch.randn(n, generator=g) * (10.0 ** torch.randint(-9, -2, (n,), generator=g).float())
        O.keras_adam_update(wo, mo, vo, gr, step, cfg)
        lib.gct2_direct_adam(_p(wd_), _p(md_), _p(vd_), _p(_np(gr)), n, float(warm(step)), cfg.beta1, cfg.beta2,
                             cfg.epsilon, step + 1)
    np.testing.assert_allclose(wd_ - _np(w0), _np(wo - w0), rtol=2e-3, atol=1e-10)
    np.testing.assert_allclose(md_, _np(mo), rtol=1e-5, atol=1e-12)
    np.testing.assert_allclose(vd_, _np(vo), rtol=1e-5, atol=1e-20)


def test_whole_forward_pass_matches_with_the_wiring_restated(lib):
    """Denoiser.call on a 2-octave 16x16 variant: the U-Net wiring of train.py:175-215 restated here with the direct C
    ops -- down0..down1, up1..up0, every Residual concatenating [module output, skip] on the last axis (train.py:113-119),
    Dense(3) on the full-resolution concat -- must reproduce oracle.denoiser_forward."""
    cfg = O.Config(size=16, pixel_size=8, max_size=16, octaves=2)
    w = O.glorot_init(cfg, 3)
    g = torch.Generator().manual_seed(8)
    for k in list(w):
        if k.endswith("bias"):
            w[k] = torch.randn(w[k].shape, generator=g) * 0.1  # Keras initialises biases to 0: make them matter
    x = torch.rand(2, 16, 16, 3, generator=g) * 2 - 1
    ref = _np(O.denoiser_forward(w, x, cfg))
    n = cfg.octaves
    acts = [_np(x)]
    for i in range(n):  # the down path; acts[i] is the input (and skip) of octave i
        acts.append(conv(lib, acts[-1], _np(w[f"down{i}/kernel"]), _np(w[f"down{i}/bias"])))
    h = acts[n]
    for i in reversed(range(n)):
        up = convT(lib, h, _np(w[f"up{i}/kernel"]), _np(w[f"up{i}/bias"]))
        h = np.ascontiguousarray(np.concatenate([up, acts[i]], axis=-1))  # module output first, skip second
    out = np.empty((2, 16, 16, 3), np.float32)
    lib.gct2_direct_dense(_p(h), _p(_np(w["dense/kernel"])), _p(_np(w["dense/bias"])), _p(out), 2 * 256, h.shape[-1], 3)
    np.testing.assert_allclose(out, ref, rtol=2e-4, atol=2e-5)


# ---- the dormant switches (train.py:20,26-27) against the direct C ops
def conv3(lib, x, w, b, relu=True):
    """Conv2D(filters, 3, 1, 'same') through the generic direct loop (SAME padding from TensorFlow's rule: 1 before)."""
    B, H, W, Cin = x.shape
    y = np.empty((B, H, W, w.shape[3]), np.float32)
    lib.gct2_direct_conv2d_same(_p(x), _p(w), _p(b), _p(y), B, H, W, Cin, w.shape[3], 3, 1, int(relu))
    return y


@pytest.mark.parametrize("shape", [(2, 8, 8, 5, 7), (1, 4, 4, 3, 4), (1, 5, 9, 2, 3)])
def test_block_convolution_matches_the_direct_definition(lib, shape):
    B, H, W, Ci, Co = shape
    g = torch.Generator().manual_seed(H * 10 + Ci)
    x = torch.randn(B, H, W, Ci, generator=g)
    w, b = torch.randn(3, 3, Ci, Co, generator=g) * 0.3, torch.randn(Co, generator=g)
    np.testing.assert_allclose(conv3(lib, _np(x), _np(w), _np(b)), _np(O.conv3x3(x, w, b)), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("kw", [dict(block_depth=1), dict(block_depth=2, concat=False), dict(residual=True),
                                dict(residual=True, block_depth=1)],
                         ids=["depth1", "depth2-no-concat", "residual", "residual-depth1"])
def test_whole_forward_pass_of_the_dormant_wirings(lib, kw):
    """train.py:175-204 with block_depth > 0 / residual = True / concat = False, the recursion restated here over the direct
    C ops: Sequential[Block, Residual(Sequential[DownShuffle, Block, inner, Block, UpShuffle]), Block, Dense(3)], a Block =
    block_depth 3x3 convolutions, a Residual = input + Dense_nobias(module(input)) | concat([module(input), input]) |
    module(input) -- must reproduce oracle.denoiser_forward."""
    import dataclasses
    cfg = dataclasses.replace(O.Config(size=16, pixel_size=8, max_size=16, octaves=2), **kw)
    w = O.glorot_init(cfg, 3)
    g = torch.Generator().manual_seed(8)
    for k in list(w):
        if k.endswith("bias"):
            w[k] = torch.randn(w[k].shape, generator=g) * 0.1
    x = torch.rand(2, 16, 16, 3, generator=g) * 2 - 1
    ref = _np(O.denoiser_forward(w, x, cfg))

    def block(prefix, h):
        for k in range(cfg.block_depth):
            h = conv3(lib, h, _np(w[f"{prefix}/conv{k}/kernel"]), _np(w[f"{prefix}/conv{k}/bias"]))
        return h

    def residual(i, h):
        d = block(f"block_down{i}", conv(lib, h, _np(w[f"down{i}/kernel"]), _np(w[f"down{i}/bias"])))
        inner = residual(i + 1, d) if i + 1 < cfg.octaves else block("block_mid", d)
        u = convT(lib, block(f"block_up{i}", inner), _np(w[f"up{i}/kernel"]), _np(w[f"up{i}/bias"]))
        if cfg.residual:
            kern = _np(w[f"res{i}/dense/kernel"])
            proj = np.empty(h.shape, np.float32)
            zero = np.zeros(kern.shape[1], np.float32)
            lib.gct2_direct_dense(_p(u), _p(kern), _p(zero), _p(proj), u.size // u.shape[-1], u.shape[-1], kern.shape[1])
            return np.ascontiguousarray(h + proj)
        if cfg.concat:
            return np.ascontiguousarray(np.concatenate([u, h], axis=-1))
        return u

    top = block("block_out", residual(0, block("block_in", _np(x))))
    out = np.empty((2, 16, 16, 3), np.float32)
    lib.gct2_direct_dense(_p(top), _p(_np(w["dense/kernel"])), _p(_np(w["dense/bias"])), _p(out), 2 * 256, top.shape[-1], 3)
    np.testing.assert_allclose(out, ref, rtol=3e-4, atol=3e-5)
