"""Tensor-level wrappers over the C ABI (include/gct2_b200.h).

Each function takes torch CUDA tensors (used only as device-memory handles), derives the pixel strides the
kernels need from the tensors' own strides -- so a channel slice ``cat[..., a:b]`` of a concat buffer is
passed as-is and the reference's ``tf.concat`` (train.py:113-119) never becomes a copy -- and enqueues the
kernel on the current CUDA stream.  There is no fallback: a missing library or a non-sm_100 device raises.

Activation tensors are NHWC bf16; kernels are in the Keras layouts (Conv2D [4,4,Cin,Cout], Conv2DTranspose
[4,4,Cout,Cin]) as bf16 shadow copies; gradients and optimizer state are fp32.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from ._lib import check, current_stream, ptr


def launch_count() -> int:
    """Kernels this library has enqueued so far in this process (counted inside the C library)."""
    return int(_lib.load().gct2_launch_count())


def last_plan() -> dict:
    """Test hook: the plan of the most recent tensor-core conv launch (gct2_debug_last_plan)."""
    import ctypes
    out = (ctypes.c_int * 8)()
    _lib.load().gct2_debug_last_plan(out)
    keys = ("BN", "splits", "pair", "fused", "grid", "stages", "rounds", "early")
    return dict(zip(keys, list(out)))


def set_sm_budget(sms: int) -> None:
    """CTAs a tensor-core conv launch may occupy (0 = all SMs); see gct2_set_sm_budget."""
    _lib.load().gct2_set_sm_budget(int(sms))


def set_adam_sms(sms: int) -> None:
    """SMs of the following adam_apply launches (0 = whole chip); see gct2_set_adam_sms."""
    _lib.load().gct2_set_adam_sms(int(sms))


#: when a list, every op appends (op name, start event, end event) -- bench.py's per-kernel timing pass
_profile = None


def profile_ops(enable: bool):
    """Starts (True) or stops (False) per-op CUDA-event bracketing; stopping returns the recorded list."""
    global _profile
    if enable:
        _profile = []
        return None
    rec, _profile = _profile, None
    return rec


def _timed(fn):
    import functools

    @functools.wraps(fn)
    def wrapper(*a, **kw):
        if _profile is None:
            return fn(*a, **kw)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        out = fn(*a, **kw)
        e.record()
        _profile.append((fn.__name__, s, e))
        return out

    return wrapper


def _lib_for(t: torch.Tensor):
    if not t.is_cuda:
        raise _lib.Gct2Error("gct2 ops need CUDA tensors: there is no CPU path")
    return _lib.init(t.device.index if t.device.index is not None else torch.cuda.current_device())


#: the two 16-bit storage formats: bf16 (default) and fp16 (the reference's mixed_float16 policy, train.py:34,43-45)
HALF = (torch.bfloat16, torch.float16)
_policy_now = [None]


def _policy(t: torch.Tensor) -> None:
    """gct2_set_policy from the dtype of a 16-bit tensor handed to an op (host-side state, read at launch)."""
    f16 = t.dtype == torch.float16
    if _policy_now[0] != f16:
        _lib.load().gct2_set_policy(int(f16))
        _policy_now[0] = f16


def _nhwc(t: torch.Tensor, dtype) -> int:
    """Checks a [B,H,W,C] (possibly channel-sliced) view and returns its pixel stride in elements.  dtype
    torch.bfloat16 stands for "the 16-bit storage format": fp16 tensors are accepted too and select the fp16 policy."""
    if dtype == torch.bfloat16 and t.dtype in HALF:
        _policy(t)
        dtype = t.dtype
    if t.dtype != dtype or t.dim() != 4 or t.stride(3) != 1:
        raise _lib.Gct2Error(f"expected an NHWC {dtype} tensor with unit channel stride, got {t.dtype} {tuple(t.shape)} "
                             f"strides {t.stride()}")
    ld = t.stride(2)
    if t.stride(1) != ld * t.shape[2] or (t.shape[0] > 1 and t.stride(0) != ld * t.shape[1] * t.shape[2]):
        raise _lib.Gct2Error(f"tensor is not a channel slice of a dense NHWC buffer: strides {t.stride()}")
    return ld


class Workspace:
    """fp32 split-K scratch shared by the conv calls of ONE stream (contents irrelevant between calls)."""

    def __init__(self, nbytes: int, device):
        self.buf = torch.empty(max(nbytes, 16) // 4, dtype=torch.float32, device=device)

    @property
    def nbytes(self) -> int:
        return self.buf.numel() * 4


@_timed
def noise_images(x, eps, t_int, out, steps: int = 200):
    """train.py:224-234: out = x*sqrt(abar(t)) + eps*sqrt(1-abar(t)); x, eps, out fp32 [B,H,W,3]; t_int int32 [B]."""
    lib = _lib_for(x)
    B = x.shape[0]
    check(lib.gct2_noise_images(ptr(x), ptr(eps), ptr(t_int), ptr(out), B, x.numel() // B, steps, current_stream()))
    return out


@_timed
def conv4s2_c3_fprop(x, w, bias, y):
    """DownShuffle on the fp32 3-channel image (train.py:158-169, down0). w fp32 [4,4,3,Cout]."""
    lib = _lib_for(x)
    B, H, W, _ = x.shape
    check(lib.gct2_conv4s2_c3_fprop(ptr(x), ptr(w), ptr(bias), ptr(y), _nhwc(y, torch.bfloat16), B, H, W, y.shape[3],
                                    current_stream()))
    return y


@_timed
def conv4s2_c3_wgrad(x, dz, dw, db=None, accumulate: bool = False):
    """Weight (and, when db is given, bias) gradient of down0.  accumulate=True adds into pre-zeroed outputs."""
    lib = _lib_for(x)
    B, H, W, _ = x.shape
    check(lib.gct2_conv4s2_c3_wgrad(ptr(x), ptr(dz), _nhwc(dz, torch.bfloat16), ptr(dw), ptr(db), B, H, W, dz.shape[3],
                                    int(accumulate), current_stream()))


@_timed
def conv4s2_fprop(x, w, bias, y, ws: Workspace, weights_stable: bool = False):
    """DownShuffle forward (train.py:158-169): y = relu(conv2d(x, w, s=2, SAME) + b).  w bf16 [4,4,Cin,Cout].
    weights_stable: GCT2_WEIGHTS_STABLE (w is not being written by anything that may still run when this launch starts)."""
    lib = _lib_for(x)
    B, H, W, Cin = x.shape
    check(lib.gct2_conv4s2_fprop(ptr(x), _nhwc(x, torch.bfloat16), ptr(w), ptr(bias), ptr(y), _nhwc(y, torch.bfloat16),
                                 B, H, W, Cin, y.shape[3], ptr(ws.buf), ws.nbytes, int(weights_stable), current_stream()))
    return y


@_timed
def conv4s2_dgrad(dy, w, dx, act, add_old: bool, ws: Workspace, weights_stable: bool = False):
    """Backward-data of DownShuffle, fused with the ReLU mask of the layer that produced the input and the add of
    the skip-path gradient already sitting in dx."""
    lib = _lib_for(dy)
    B, H, W, Cin = dx.shape
    check(lib.gct2_conv4s2_dgrad(ptr(dy), _nhwc(dy, torch.bfloat16), ptr(w), ptr(dx), _nhwc(dx, torch.bfloat16),
                                 ptr(act), _nhwc(act, torch.bfloat16), int(add_old), B, H, W, Cin, dy.shape[3],
                                 ptr(ws.buf), ws.nbytes, int(weights_stable), current_stream()))
    return dx


@_timed
def conv4s2_wgrad(x, dy, dw, ws: Optional[Workspace] = None):
    lib = _lib_for(x)
    B, H, W, Cin = x.shape
    check(lib.gct2_conv4s2_wgrad(ptr(x), _nhwc(x, torch.bfloat16), ptr(dy), _nhwc(dy, torch.bfloat16), ptr(dw), B, H, W,
                                 Cin, dy.shape[3], ptr(ws.buf) if ws else 0, ws.nbytes if ws else 0, current_stream()))
    return dw


@_timed
def convT4s2_fprop(x, w, bias, y, ws: Workspace, weights_stable: bool = False):
    """UpShuffle forward (train.py:145-156): y = relu(conv2d_transpose(x, w, s=2, SAME) + b). w bf16 [4,4,Cout,Cin]."""
    lib = _lib_for(x)
    B, H, W, Cin = x.shape
    check(lib.gct2_convT4s2_fprop(ptr(x), _nhwc(x, torch.bfloat16), ptr(w), ptr(bias), ptr(y),
                                  _nhwc(y, torch.bfloat16), B, H, W, Cin, y.shape[3], ptr(ws.buf), ws.nbytes,
                                  int(weights_stable), current_stream()))
    return y


@_timed
def convT4s2_dgrad(dy, w, dx, act, mask_channels: int, ws: Workspace, weights_stable: bool = False):
    """Backward-data of UpShuffle; channels [0, mask_channels) of dx are ReLU-masked by act, the rest stored raw."""
    lib = _lib_for(dy)
    B, H, W, Cin = dx.shape
    check(lib.gct2_convT4s2_dgrad(ptr(dy), _nhwc(dy, torch.bfloat16), ptr(w), ptr(dx), _nhwc(dx, torch.bfloat16),
                                  ptr(act), _nhwc(act, torch.bfloat16), mask_channels, B, H, W, Cin, dy.shape[3],
                                  ptr(ws.buf), ws.nbytes, int(weights_stable), current_stream()))
    return dx


@_timed
def convT4s2_wgrad(x, dy, dw, ws: Optional[Workspace] = None):
    lib = _lib_for(x)
    B, H, W, Cin = x.shape
    check(lib.gct2_convT4s2_wgrad(ptr(x), _nhwc(x, torch.bfloat16), ptr(dy), _nhwc(dy, torch.bfloat16), ptr(dw), B, H, W,
                                  Cin, dy.shape[3], ptr(ws.buf) if ws else 0, ws.nbytes if ws else 0, current_stream()))
    return dw


@_timed
def conv3s1_fprop(x, w, bias, y, ws: Workspace, weights_stable: bool = False):
    """Block's Conv2D(filters, ks, 1, 'same', relu) (train.py:131-139; ks = 3): y = relu(conv2d(x, w, s=1, SAME) + b).
    w 16-bit [ks,ks,Cin,Cout] (ks = 3 or 1)."""
    lib = _lib_for(x)
    B, H, W, Cin = x.shape
    check(lib.gct2_conv3s1_fprop(ptr(x), _nhwc(x, torch.bfloat16), ptr(w), ptr(bias), ptr(y), _nhwc(y, torch.bfloat16),
                                 B, H, W, Cin, y.shape[3], w.shape[0], ptr(ws.buf), ws.nbytes, int(weights_stable),
                                 current_stream()))
    return y


@_timed
def conv3s1_fprop_add(x, w, res, y, weights_stable: bool = False):
    """train.py:110-111 (residual = True): y = res + Dense(C, use_bias=False)(x) -- the 1x1 case of the stride-1 map with
    the add fused into the epilogue.  w 16-bit [1,1,Cin,Cout] (a Dense kernel [Cin,Cout] viewed as a 1x1 conv)."""
    lib = _lib_for(x)
    B, H, W, Cin = x.shape
    check(lib.gct2_conv3s1_fprop_add(ptr(x), _nhwc(x, torch.bfloat16), ptr(w), ptr(res), _nhwc(res, torch.bfloat16), ptr(y),
                                     _nhwc(y, torch.bfloat16), B, H, W, Cin, y.shape[3], w.shape[0], int(weights_stable),
                                     current_stream()))
    return y


@_timed
def conv3s1_dgrad(dy, w, dx, act, mask_channels: int, add_old: bool, ws: Workspace, weights_stable: bool = False):
    """Backward-data of the stride-1 conv: dx = mask(conv2d_backprop_input(dy, w) (+ dx)); channels [0, mask_channels)
    of dx are ReLU-masked by act, the rest stored raw (a concat buffer's skip slice)."""
    lib = _lib_for(dy)
    B, H, W, Cin = dx.shape
    check(lib.gct2_conv3s1_dgrad(ptr(dy), _nhwc(dy, torch.bfloat16), ptr(w), ptr(dx), _nhwc(dx, torch.bfloat16),
                                 ptr(act), _nhwc(act, torch.bfloat16), int(mask_channels), int(add_old), B, H, W, Cin,
                                 dy.shape[3], w.shape[0], ptr(ws.buf), ws.nbytes, int(weights_stable), current_stream()))
    return dx


@_timed
def conv3s1_wgrad(x, dy, dw, ws: Optional[Workspace] = None):
    """dw fp32 [ks,ks,Cin,Cout] = conv2d_backprop_filter(x, dy) of the stride-1 conv."""
    lib = _lib_for(x)
    B, H, W, Cin = x.shape
    check(lib.gct2_conv3s1_wgrad(ptr(x), _nhwc(x, torch.bfloat16), ptr(dy), _nhwc(dy, torch.bfloat16), ptr(dw), B, H, W,
                                 Cin, dy.shape[3], dw.shape[0], ptr(ws.buf) if ws else 0, ws.nbytes if ws else 0,
                                 current_stream()))
    return dw


@_timed
def conv3s1_c3_fprop(x, w, bias, y):
    """The first conv of the outermost Block (block_depth > 0) on the fp32 3-channel image. w fp32 [3,3,3,Cout]."""
    lib = _lib_for(x)
    B, H, W, _ = x.shape
    check(lib.gct2_conv3s1_c3_fprop(ptr(x), ptr(w), ptr(bias), ptr(y), _nhwc(y, torch.bfloat16), B, H, W, y.shape[3],
                                    current_stream()))
    return y


@_timed
def conv3s1_c3_wgrad(x, dz, dw, accumulate: bool = False):
    lib = _lib_for(x)
    B, H, W, _ = x.shape
    check(lib.gct2_conv3s1_c3_wgrad(ptr(x), ptr(dz), _nhwc(dz, torch.bfloat16), ptr(dw), B, H, W, dz.shape[3],
                                    int(accumulate), current_stream()))
    return dw


@_timed
def bias_grad(dz, db):
    lib = _lib_for(dz)
    ld = _nhwc(dz, torch.bfloat16)
    rows = dz.shape[0] * dz.shape[1] * dz.shape[2]
    check(lib.gct2_bias_grad(ptr(dz), ld, rows, dz.shape[3], ptr(db), current_stream()))
    return db


class BiasGradPlan:
    """Host-side argument arrays of gct2_bias_grad_multi, built once for a fixed set of buffers."""

    def __init__(self, dzs, dbs):
        import ctypes
        n = len(dzs)
        self.n = n
        self.keep = (list(dzs), list(dbs))
        self.dz = (ctypes.c_void_p * n)(*[ptr(t) for t in dzs])
        self.db = (ctypes.c_void_p * n)(*[ptr(t) for t in dbs])
        self.ld = (ctypes.c_int * n)(*[_nhwc(t, torch.bfloat16) for t in dzs])
        self.rows = (ctypes.c_longlong * n)(*[t.shape[0] * t.shape[1] * t.shape[2] for t in dzs])
        self.C = (ctypes.c_int * n)(*[t.shape[3] for t in dzs])
        for t, d in zip(dzs, dbs):
            if d.numel() != t.shape[3] or d.dtype != torch.float32:
                raise _lib.Gct2Error("bias gradient outputs must be fp32 [C]")


@_timed
def bias_grad_multi(plan: BiasGradPlan, accumulate: bool = False):
    """BiasAddGrad of every conv layer in one launch: db_i[c] = sum over pixels of dz_i[..., c]."""
    lib = _lib_for(plan.keep[0][0])
    _policy(plan.keep[0][0])
    check(lib.gct2_bias_grad_multi(plan.n, plan.dz, plan.ld, plan.rows, plan.C, plan.db, int(accumulate),
                                   current_stream()))


TARGET_X, TARGET_EPSILON, TARGET_SCALED, TARGET_WEIGHTED, TARGET_ODE = 0, 1, 2, 4, 8


def target_mode(predict_x: bool = True, predict_scaled_epsilon: bool = False, prediction_weighting: bool = False,
                ordinary_differential_equation: bool = False) -> int:
    """train.py:29-32 -> the GCT2_TARGET_* bits (train.py:238-252: ODE takes precedence, then predict_x)."""
    if ordinary_differential_equation:
        return TARGET_ODE
    if predict_x:
        return TARGET_X
    return TARGET_EPSILON | (TARGET_SCALED if predict_scaled_epsilon else 0) | (TARGET_WEIGHTED if prediction_weighting else 0)


@_timed
def dense_mse(u0, noised, x, wd, bd, loss, inv_n: float, pred=None, du0=None, dwd=None, dbd=None,
              accumulate: bool = False, loss_scale=None, eps=None, t_int=None, mode: int = 0, steps: int = 200):
    """Dense(3) on concat([u0, noised]) (train.py:198-202) fused with the MSE (train.py:262-272) and, when du0 is
    given, their backward.  noised=None: the layer reads u0's channels only (behind a Block, or without the skip)."""
    lib = _lib_for(u0)
    backward = du0 is not None
    pixels = u0.shape[0] * u0.shape[1] * u0.shape[2]
    check(lib.gct2_dense_mse(ptr(u0), _nhwc(u0, torch.bfloat16), ptr(noised), ptr(x), ptr(wd), ptr(bd), ptr(pred),
                             ptr(loss), ptr(du0), _nhwc(du0, torch.bfloat16) if backward else 0, ptr(dwd), ptr(dbd),
                             pixels, u0.shape[3], inv_n, int(backward), int(accumulate), ptr(loss_scale),
                             ptr(eps), ptr(t_int), u0.shape[1] * u0.shape[2], int(mode), int(steps), current_stream()))
    return loss


@_timed
def res0_compose(wp, wd, weff):
    """weff [U+3,3] = [wp . wd ; wd]: the image-level residual projection folded into Dense(3) (see gct2_res0_compose)."""
    lib = _lib_for(wp)
    check(lib.gct2_res0_compose(ptr(wp), ptr(wd), ptr(weff), wp.shape[0], current_stream()))
    return weff


@_timed
def res0_decompose(dweff, wp, wd, dwp, dwd):
    """dwp = dweff[:U] . wd^T, dwd = wp^T . dweff[:U] + dweff[U:] (see gct2_res0_decompose)."""
    lib = _lib_for(wp)
    check(lib.gct2_res0_decompose(ptr(dweff), ptr(wp), ptr(wd), ptr(dwp), ptr(dwd), wp.shape[0], current_stream()))


@_timed
def adam_keras(w, m, v, g, w_bf16, iterations, hyper, base_lr: float, warmup_steps: int, beta1: float = 0.9,
               beta2: float = 0.999, eps: float = 1e-7, grad_scale: float = 1.0):
    """tf.keras.optimizers.Adam(WarmUp(base_lr, warmup_steps)) on flat fp32 buffers (train.py:50-65,75)."""
    lib = _lib_for(w)
    _policy(w_bf16)
    check(lib.gct2_adam_keras(ptr(w), ptr(m), ptr(v), ptr(g), ptr(w_bf16), w.numel(), ptr(iterations), ptr(hyper),
                              base_lr, warmup_steps, beta1, beta2, eps, grad_scale, current_stream()))


@_timed
def adam_prepare(iterations, hyper, base_lr: float, warmup_steps: int, beta1: float = 0.9, beta2: float = 0.999):
    """Once per step: alpha = lr(step)*sqrt(1-b2^t)/(1-b1^t) into hyper[0], then iterations += 1."""
    lib = _lib_for(hyper)
    check(lib.gct2_adam_prepare(ptr(iterations), ptr(hyper), base_lr, warmup_steps, beta1, beta2, current_stream()))


@_timed
def adam_apply(w, m, v, g, w_bf16, hyper, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-7,
               grad_scale: float = 1.0, iterations_inc=None, loss_scale_state=None):
    """Keras-Adam update of one contiguous range of the flat buffers (views starting on 16-byte boundaries).
    iterations_inc: the optimiser's iteration counter, incremented by this launch (pass it on the step's last range
    when the step was opened with step_begin)."""
    lib = _lib_for(w)
    _policy(w_bf16)
    if g.dtype == torch.bfloat16:  # summed data-parallel gradients
        check(lib.gct2_adam_apply_g16(ptr(w), ptr(m), ptr(v), ptr(g), ptr(w_bf16), w.numel(), ptr(hyper), beta1, beta2, eps,
                                      grad_scale, ptr(iterations_inc), current_stream()))
    else:
        check(lib.gct2_adam_apply(ptr(w), ptr(m), ptr(v), ptr(g), ptr(w_bf16), w.numel(), ptr(hyper), beta1, beta2, eps,
                                  grad_scale, ptr(iterations_inc), ptr(loss_scale_state), current_stream()))


@_timed
def step_begin(x, noised, iterations, hyper, gsmall, loss, seed: int, steps: int, base_lr: float, warmup_steps: int,
               beta1: float = 0.9, beta2: float = 0.999, eps_out=None, t_out=None):
    """Fused step prologue: device-side draws of t_int / eps (train.py:224-227), noising (train.py:231-234), zeroing of
    the atomically-accumulated gradients and the loss, Adam alpha of this step."""
    lib = _lib_for(x)
    B = x.shape[0]
    check(lib.gct2_step_begin(ptr(x), ptr(noised), ptr(eps_out), ptr(t_out), B, x.numel() // B, steps, seed,
                              ptr(iterations), ptr(hyper), base_lr, warmup_steps, beta1, beta2, ptr(gsmall),
                              gsmall.numel(), ptr(loss), current_stream()))


@_timed
def step_begin_u8(img, flip, x_out, noised, iterations, hyper, gsmall, loss, seed: int, steps: int, base_lr: float,
                  warmup_steps: int, beta1: float = 0.9, beta2: float = 0.999, eps_out=None, t_out=None):
    """step_begin on the uint8 batch of decode_file (train.py:285-293): x = img/128 - 1 (+ per-image left-right flip)
    is decoded on the fly, written to x_out for the loss, and noised."""
    lib = _lib_for(img)
    if img.dtype != torch.uint8 or not img.is_contiguous() or img.dim() != 4 or img.shape[3] != 3:
        raise ValueError("step_begin_u8 expects a contiguous uint8 [B,H,W,3] batch")
    if flip is not None and (flip.dtype != torch.uint8 or flip.numel() != img.shape[0]):
        raise ValueError("flip must be uint8 [B]")
    B = img.shape[0]
    check(lib.gct2_step_begin_u8(ptr(img), ptr(flip), ptr(x_out), img.shape[2], ptr(noised), ptr(eps_out), ptr(t_out), B,
                                 img.numel() // B, steps, seed, ptr(iterations), ptr(hyper), base_lr, warmup_steps,
                                 beta1, beta2, ptr(gsmall), gsmall.numel(), ptr(loss), current_stream()))


@_timed
def sample_update(pred, fake, x_theta, eps_theta, t: int, t_next: int, steps: int = 200, mode: int = 0):
    """log_sample's per-step arithmetic (train.py:365-398, 441-468, predict_x branch): x_theta / eps_theta from the
    prediction at step t, and the next Denoiser input (fake, in place) for step t_next.  pred=None: first mix only."""
    lib = _lib_for(fake)
    check(lib.gct2_sample_update(ptr(pred), ptr(fake), ptr(x_theta), ptr(eps_theta), t, t_next, steps, fake.numel(),
                                 int(mode), current_stream()))


@_timed
def cast_bf16(src, dst):
    """fp32 -> the 16-bit storage format of dst (bf16 or fp16)."""
    lib = _lib_for(src)
    _policy(dst)
    check(lib.gct2_cast_bf16(ptr(src), ptr(dst), src.numel(), current_stream()))
    return dst


@_timed
def loss_scale_check(g, ls):
    """Clears ls[2] when any gradient is inf / NaN (tf.keras.mixed_precision.LossScaleOptimizer, train.py:82-83)."""
    lib = _lib_for(g)
    check(lib.gct2_loss_scale_check(ptr(g), g.numel(), ptr(ls), current_stream()))


@_timed
def loss_scale_update(ls, growth_steps: int = 2000):
    """Dynamic loss-scale bookkeeping after the (possibly skipped) update; re-arms the finite flag."""
    lib = _lib_for(ls)
    check(lib.gct2_loss_scale_update(ptr(ls), int(growth_steps), current_stream()))


@_timed
def latent_edits(eps_theta, dictionary, out):
    """train.py:418-432: [epsilon_theta, pixelated, shifted, quantised] from the inverted latent [1,S,S,3] and the
    dictionary [S,S,K,3]; out fp32 [4,S,S,3]."""
    lib = _lib_for(eps_theta)
    S = eps_theta.shape[-2]
    if tuple(dictionary.shape[:2]) != (S, S) or dictionary.shape[3] != 3 or out.shape[0] != 4 or not out.is_contiguous():
        raise ValueError("latent_edits expects eps_theta [1,S,S,3], dictionary [S,S,K,3] and a contiguous out [4,S,S,3]")
    check(lib.gct2_latent_edits(ptr(eps_theta.contiguous()), ptr(dictionary.contiguous()), ptr(out), S, dictionary.shape[2],
                                current_stream()))
    return out


@_timed
def rmse(a, b, out):
    """train.py:357-361 'example loss': out[0] = sqrt(mean((a - b)**2))."""
    lib = _lib_for(a)
    check(lib.gct2_rmse(ptr(a.contiguous()), ptr(b.contiguous()), a.numel(), ptr(out), current_stream()))
    return out


@_timed
def adam_apply_p2p(w, m, v, g_ptrs, w16_ptrs, world: int, elem_offset: int, hyper, beta1: float = 0.9, beta2: float = 0.999,
                   eps: float = 1e-7, grad_scale: float = 1.0, write_all: bool = True, g_mc: int = 0, w16_mc: int = 0):
    """Data parallel, fused: gradient exchange + Keras-Adam + weight broadcast of this rank's slice over NVLink peer memory
    (gct2_adam_apply_p2p).  w, m, v: this rank's fp32 masters of the slice; g_ptrs / w16_ptrs: per-rank BASE device
    pointers (ints) of the bf16 gradient buffers and the 16-bit weight shadows; elem_offset: where the slice starts in
    them; g_mc / w16_mc: NVLS multicast addresses or 0."""
    import ctypes
    lib = _lib_for(w)
    ga = (ctypes.c_void_p * world)(*g_ptrs[:world])
    wa = (ctypes.c_void_p * world)(*w16_ptrs[:world])
    check(lib.gct2_adam_apply_p2p(ptr(w), ptr(m), ptr(v), ga, wa, g_mc or None, w16_mc or None, world, int(elem_offset),
                                  w.numel(), ptr(hyper), beta1, beta2, eps, grad_scale, int(write_all), current_stream()))


@_timed
def sum_peers_f32(src_ptrs, world: int, out_a, out_b):
    """out_a / out_b (fp32, consecutive in every rank's staging buffer) = the sum over ranks of the buffers at src_ptrs
    (per-rank base device pointers, ints), added in rank order (gct2_sum_peers_f32)."""
    import ctypes
    lib = _lib_for(out_a)
    sa = (ctypes.c_void_p * world)(*src_ptrs[:world])
    check(lib.gct2_sum_peers_f32(sa, world, ptr(out_a), out_a.numel(), ptr(out_b), out_b.numel(), current_stream()))
