"""CPU oracle: an fp32 PyTorch-CPU restatement of the training step of the reference's train.py.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (gan_class_transfer2_b200/) imports this module;
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may.

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures, TensorFlow is
not installable in this environment, and train.py cannot even be imported without a GPU and the author's
dataset (train.py:40,305,315).  This restatement therefore follows the TensorFlow/Keras semantics documented
in SURVEY.md Appendix A and is pinned only by its own self-consistency tests (tests/test_oracle.py): parameter
count 41 691 660, shape walk, conv-transpose == autograd-dgrad of the SAME-padded strided conv, the 4-phase
identity, Dense == 1x1 conv, a finite-difference gradient check and an Adam closed-form known answer -- and by
tests/test_oracle_direct.py against oracle/direct.c, a plain-C loop-level restatement written from the TensorFlow
definitions of the same ops (SAME padding from TF's rule, Conv2DTranspose as Conv2D's input-gradient scatter) that
shares no code with this file.

What follows what (reference = /root/reference/train.py):
  Config            :17-36   module-level hyper-parameters
  alpha_dash        :85-93
  WarmUp            :50-65   ; keras_adam_update: tf.keras.optimizers.Adam defaults (:75), Keras formula
  variable_specs    :175-204 construction recursion (Denoiser.__init__) + Keras variable layouts/order
  glorot_init       :134,149,162 (explicit glorot_uniform) and the Keras Dense default
  denoiser_forward  :97-121 (concat branch), :123-143 (block_depth=0 -> identity), :145-169, :206-215
  trainer_loss      :223-236,243-244,262-263,272 (RNG t_int/eps are *inputs*: TF never seeds them)
  identity          :171-173
"""
from __future__ import annotations

import dataclasses
import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


@dataclasses.dataclass(frozen=True)
class Config:
    """train.py:17-36 defaults."""
    size: int = 256
    pixel_size: int = 128
    max_size: int = 512
    block_depth: int = 0
    octaves: int = 6
    batch_size: int = 1
    steps: int = 200
    warm_up: int = 2000
    base_lr: float = 2e-5
    beta1: float = 0.9
    beta2: float = 0.999
    epsilon: float = 1e-7
    # train.py:29-32: what the network predicts / what the loss compares (all at the reference's defaults)
    predict_x: bool = True
    predict_scaled_epsilon: bool = False
    prediction_weighting: bool = False
    ordinary_differential_equation: bool = False
    test_step: int = 25   # train.py:95
    # train.py:26-27: how Residual joins its module and its input (reference: residual=False, concat=True)
    residual: bool = False
    concat: bool = True

    def mid_filters(self) -> int:  # train.py:179
        return min(self.pixel_size * 2 ** self.octaves, self.max_size)

    def level_in(self, i: int) -> int:
        """Channels of the tensor entering Residual level i."""
        if i == 0:
            return self.pixel_size if self.block_depth else 3
        return self.down_filters(i - 1)

    def res_out(self, i: int) -> int:
        """Channels leaving Residual level i (train.py:110-121)."""
        if self.residual:
            return self.level_in(i)
        return self.up_filters(i) + (self.level_in(i) if self.concat else 0)

    def down_filters(self, i: int) -> int:  # train.py:181
        return min(self.pixel_size * 2 ** i, self.max_size)

    def up_filters(self, i: int) -> int:  # train.py:188
        return min(self.pixel_size * 2 ** i // 2, self.max_size)


DEFAULT = Config()
#: shrunken config used by the fast tests (keeps a 4x4 bottleneck like the reference's comment at train.py:21)
TINY = Config(size=64, pixel_size=128, max_size=256, octaves=4)
#: BASELINE.json config 4 ("doubled-resolution / widened-channel variant"): 512 px, 7 octaves, up to 1024 channels,
#: 217 078 796 parameters, 2062 GFLOP per image per step; still a 4x4 bottleneck (train.py:21)
WIDE = Config(size=512, pixel_size=256, max_size=1024, octaves=7)


def alpha_dash(t, steps: int = 200):
    """train.py:85-93: (1 - t/(steps+1))**2 * 0.25 ; works on tensors and Python numbers."""
    t = t / (steps + 1)
    return (1 - t) ** 2 * 0.25


class WarmUp:
    """train.py:50-65. `step` is the 0-based optimizer iteration."""

    def __init__(self, base: float, warmup_steps: int):
        self.base = base
        self.warmup_steps = warmup_steps

    def __call__(self, step: int) -> float:
        if step < self.warmup_steps:
            return float(torch.tensor(self.base, dtype=torch.float32) * torch.tensor(float(step + 1), dtype=torch.float32)
                         / (self.warmup_steps + 1))
        return self.base


# --------------------------------------------------------------------------------------------- variables
def variable_specs(cfg: Config = DEFAULT) -> List[Tuple[str, Tuple[int, ...]]]:
    """Variable list in construction order (train.py:175-204), Keras layouts: Conv2D.kernel
    [k,k,Cin,Cout]; Conv2DTranspose.kernel [4,4,Cout,Cin]; Dense.kernel [Cin,Cout]; kernel then bias.

    Reference defaults (block_depth=0, residual=False, concat=True): down0..down{n-1}, up{n-1}..up0, dense.  A Block
    (train.py:123-143) adds block_depth Conv2D(filters, 3, 1, 'same') layers: block_in (:192), block_down{i} (:185),
    block_mid (:179), block_up{i} (:187), block_out (:194).  residual=True adds res{i}/dense/kernel, the bias-free
    Dense(input_channels) of train.py:106-108."""
    specs: List[Tuple[str, Tuple[int, ...]]] = []
    n, d = cfg.octaves, cfg.block_depth

    def block(prefix, cin, filters):
        for k in range(d):
            specs.append((f"{prefix}/conv{k}/kernel", (3, 3, cin, filters)))
            specs.append((f"{prefix}/conv{k}/bias", (filters,)))
            cin = filters
        return cin

    cin = block("block_in", 3, cfg.pixel_size)
    for i in range(n):
        co = cfg.down_filters(i)
        specs.append((f"down{i}/kernel", (4, 4, cin, co)))
        specs.append((f"down{i}/bias", (co,)))
        cin = block(f"block_down{i}", co, co)
    cin = block("block_mid", cin, cfg.mid_filters())
    # innermost: up_{n-1} consumes the innermost Block's output; up_i (i < n-1) what Residual_{i+1} returns
    for i in reversed(range(n)):
        ci = block(f"block_up{i}", cin if i == n - 1 else cfg.res_out(i + 1), cfg.down_filters(i))
        co = cfg.up_filters(i)
        specs.append((f"up{i}/kernel", (4, 4, co, ci)))
        specs.append((f"up{i}/bias", (co,)))
        if cfg.residual:
            specs.append((f"res{i}/dense/kernel", (co, cfg.level_in(i))))
    c = block("block_out", cfg.res_out(0), cfg.pixel_size)
    specs.append(("dense/kernel", (c, 3)))
    specs.append(("dense/bias", (3,)))
    return specs


def param_count(cfg: Config = DEFAULT) -> int:
    return sum(math.prod(s) for _, s in variable_specs(cfg))


def glorot_init(cfg: Config = DEFAULT, seed: int = 0) -> Dict[str, torch.Tensor]:
    """glorot_uniform kernels (limit sqrt(6/(fan_in+fan_out)), fans = receptive field x last two dims), zero biases."""
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    for name, shape in variable_specs(cfg):
        if name.endswith("bias"):
            out[name] = torch.zeros(shape, dtype=torch.float32)
            continue
        rf = math.prod(shape[:-2]) if len(shape) > 2 else 1
        fan_a, fan_b = shape[-2] * rf, shape[-1] * rf
        limit = math.sqrt(6.0 / (fan_a + fan_b))
        out[name] = (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * limit
    return out


def synthetic_batch(cfg: Config, batch: int, seed: int = 1):
    """x = k/128 - 1 with k ~ U{0..255} (decode_file's value grid, train.py:292),
    t_int ~ U{1..steps} (train.py:224-226), eps ~ N(0,1) (train.py:227)."""
    g = torch.Generator().manual_seed(seed)
    k = torch.randint(0, 256, (batch, cfg.size, cfg.size, 3), generator=g)
    x = k.to(torch.float32) / 128 - 1
    t_int = torch.randint(1, cfg.steps + 1, (batch,), generator=g, dtype=torch.int32)
    eps = torch.randn((batch, cfg.size, cfg.size, 3), generator=g, dtype=torch.float32)
    return x, t_int, eps


def decode_u8(img_u8: torch.Tensor, flip=None) -> torch.Tensor:
    """train.py:285-293 decode_file after the crop: (optional) left-right mirror per image
    (tf.image.random_flip_left_right, :290 -- the random draw is an input here, like every RNG of the oracle), then
    cast(image, float32) / 128 - 1 (:292).  img_u8 uint8 [B,H,W,3]; flip: iterable of B flags or None."""
    x = img_u8.to(torch.float32) / 128 - 1
    if flip is not None:
        x = x.clone()
        for b, f in enumerate(flip):
            if int(f):
                x[b] = torch.flip(x[b], dims=[1])
    return x


def sample_loop(weights, x_theta: torch.Tensor, eps_theta: torch.Tensor, t_values, cfg: "Config" = None,
                emulate_bf16: bool = False):
    """log_sample's diffusion loops, predict_x branch (train.py:365-398 with t ascending from 1, train.py:441-468 with
    t descending from steps):
        fake = alpha_dash(t)**0.5 * x_theta + (1 - alpha_dash(t))**0.5 * epsilon_theta        (:369-372, :445-448)
        prediction = denoiser((fake, [t]))                                                     (:374-377, :450-453)
        x_theta = prediction                                                                   (:394, :465)
        epsilon_theta = (fake - alpha_dash(t)**0.5 * x_theta) / (1 - alpha_dash(t))**0.5       (:395-397, :466-468)
    Returns the final (x_theta, epsilon_theta) and the list of x_theta after every step."""
    cfg = DEFAULT if cfg is None else cfg
    trace = []
    for t in t_values:
        a = alpha_dash(float(t), cfg.steps)
        fake = a ** 0.5 * x_theta + (1 - a) ** 0.5 * eps_theta
        pred = denoiser_forward(weights, fake, cfg, emulate_bf16=emulate_bf16)
        x_theta, eps_theta = sample_update(pred, fake, x_theta, eps_theta, t, cfg)
        trace.append(x_theta)
    return x_theta, eps_theta, trace


def sample_update(pred, fake, x_theta, eps_theta, t, cfg: "Config"):
    """What log_sample derives from one prediction (train.py:382-413 and :452-479, the same arithmetic in both loops)."""
    a = alpha_dash(float(t), cfg.steps)
    if cfg.ordinary_differential_equation:        # :382-391 / :452-461 -- epsilon_theta is left as it was
        a1 = alpha_dash(float(t - 1), cfg.steps)
        x_theta = (pred * (1 - a) ** 0.5 - fake * (1 - a1) ** 0.5) / (a1 ** 0.5 * (1 - a) ** 0.5 - a ** 0.5 * (1 - a1) ** 0.5)
        return x_theta, eps_theta
    if cfg.predict_x:                             # :394-398 / :464-468
        x_theta = pred
        return x_theta, (fake - a ** 0.5 * x_theta) / (1 - a) ** 0.5
    if cfg.predict_scaled_epsilon:                # :401-405 / :471-473
        eps_theta, scaled = pred / (1 - a) ** 0.5, pred
    else:                                         # :406-410 / :474-476
        eps_theta, scaled = pred, pred * (1 - a) ** 0.5
    return (fake - scaled) / a ** 0.5, eps_theta  # :411-413 / :477-479


def latent_edits(eps_theta: torch.Tensor, dictionary: torch.Tensor) -> torch.Tensor:
    """train.py:418-432: the four latents log_sample decodes again -- epsilon_theta itself, pixelated (4x4 average pooling
    + nearest up-sampling), shifted (tf.roll by one along both spatial axes) and quantised (per pixel the nearest of the
    2**bits_per_pixel entries of `dictionary` [S,S,K,3] in squared distance).  eps_theta [1,S,S,3] -> [4,S,S,3]."""
    e = eps_theta
    pooled = F.avg_pool2d(_nchw(e), 4, 4)
    pixelated = _nhwc(pooled.repeat_interleave(4, dim=2).repeat_interleave(4, dim=3))
    shifted = torch.roll(torch.roll(e, 1, 1), 1, 2)
    err = ((e[..., None, :] - dictionary[None]) ** 2).sum(-1)          # [1,S,S,K]
    idx = err.argmin(-1)                                               # first minimum, like tf.argmin
    quantised = torch.gather(dictionary[None], 3, idx[..., None, None].expand(-1, -1, -1, 1, 3)).squeeze(3)
    return torch.cat([e, pixelated, shifted, quantised], 0)


def example_denoise(weights, example_image, noise, cfg: "Config" = None, emulate_bf16=False):
    """train.py:325-361: one denoising of the fixed example at test_step and the scalar log_sample reports as
    'example loss': sqrt(mean((example - denoised)**2)).  example_image, noise: [1,S,S,3].  NOTE the reference's
    quirk, restated as written: image_factor = alpha_dash(test_step) and the mix uses image_factor**0.5 and
    (1 - image_factor)**0.5 (:325-332)."""
    cfg = DEFAULT if cfg is None else cfg
    f = alpha_dash(float(cfg.test_step), cfg.steps)
    if cfg.ordinary_differential_equation:
        f = alpha_dash(cfg.steps / 2, cfg.steps) ** 0.5          # :326-328
    noised = example_image * f ** 0.5 + noise * (1 - f) ** 0.5
    pred = denoiser_forward(weights, noised, cfg, emulate_bf16=emulate_bf16)
    if cfg.ordinary_differential_equation:                        # :338-347
        h, h1 = alpha_dash(cfg.steps / 2, cfg.steps), alpha_dash(cfg.steps / 2 - 1, cfg.steps)
        denoised = (pred * (1 - h) ** 0.5 - noised * (1 - h1) ** 0.5) / (h1 ** 0.5 * (1 - h) ** 0.5 - h ** 0.5 * (1 - h1) ** 0.5)
    elif cfg.predict_x:
        denoised = pred
    else:                                                         # :350-355
        p = pred if cfg.predict_scaled_epsilon else pred * (1 - f) ** 0.5
        denoised = (noised - p) / f ** 0.5
    return denoised, float(((example_image - denoised) ** 2).mean() ** 0.5)


# --------------------------------------------------------------------------------------------- layers
def _nchw(x):
    return x.permute(0, 3, 1, 2)


def _nhwc(x):
    return x.permute(0, 2, 3, 1)


def down_shuffle(x_nhwc, kernel_hwio, bias):
    """train.py:158-169: relu(Conv2D(filters, 4, 2, 'same')(x)); SAME for k4/s2/even N is pad (1,1)."""
    w = kernel_hwio.permute(3, 2, 0, 1)  # -> [Cout, Cin, kh, kw]
    return _nhwc(F.relu(F.conv2d(_nchw(x_nhwc), w, bias, stride=2, padding=1)))


def up_shuffle(x_nhwc, kernel_hwoi, bias):
    """train.py:145-156: relu(Conv2DTranspose(filters, 4, 2, 'same')(x)) == conv_transpose2d(k4,s2,p1), no flip
    (SURVEY A.2): out[2*iy-1+ky] += x[iy] * w[ky]."""
    w = kernel_hwoi.permute(3, 2, 0, 1)  # [4,4,Cout,Cin] -> [Cin, Cout, kh, kw]
    return _nhwc(F.relu(F.conv_transpose2d(_nchw(x_nhwc), w, bias, stride=2, padding=1)))


def conv3x3(x_nhwc, kernel_hwio, bias):
    """train.py:132-137: relu(Conv2D(filters, 3, 1, 'same')(x)); SAME for k3/s1 is pad (1,1)."""
    w = kernel_hwio.permute(3, 2, 0, 1)
    return _nhwc(F.relu(F.conv2d(_nchw(x_nhwc), w, bias, stride=1, padding=1)))


def dense(x_nhwc, kernel, bias):
    """train.py:198-202: Keras Dense contracts the last axis only; no activation."""
    return x_nhwc @ kernel + bias


class _RoundBF16(torch.autograd.Function):
    """Rounds to bf16 in the forward AND the backward direction (where the CUDA path stores bf16 tensors)."""

    @staticmethod
    def forward(ctx, t):
        return t.to(torch.bfloat16).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(torch.float32)


class _RoundF16(torch.autograd.Function):
    """The same for IEEE fp16 -- the storage format under the reference's mixed_float16 policy (train.py:43-45).
    Values beyond 65504 become inf and gradients below 6e-8 become 0, exactly as in fp16 storage: that is what the
    dynamic loss scale (train.py:82-83) exists for."""

    @staticmethod
    def forward(ctx, t):
        return t.to(torch.float16).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.float16).to(torch.float32)


def _rounding(emulate_bf16):
    """emulate_bf16: False / None (the reference's fp32 arithmetic), True / 'bf16', or 'f16'."""
    if emulate_bf16 in (False, None):
        return lambda t: t
    if emulate_bf16 in (True, "bf16"):
        return _RoundBF16.apply
    if emulate_bf16 == "f16":
        return _RoundF16.apply
    raise ValueError(f"unknown emulation {emulate_bf16!r}")


def denoiser_forward(weights: Dict[str, torch.Tensor], x_nhwc: torch.Tensor, cfg: Config = DEFAULT,
                     taps: Optional[Dict[str, torch.Tensor]] = None, emulate_bf16: bool = False,
                     force: Optional[Dict[str, torch.Tensor]] = None) -> torch.Tensor:
    """Denoiser.call (train.py:206-215): `t` is ignored by the reference; Block is identity at block_depth=0 (the
    reference's default) and a stack of 3x3 convolutions otherwise.

    Residual_i(h) = concat([Up_i(Residual_{i+1}(Down_i(h))), h], -1) with the module output FIRST (train.py:113-119).
    `taps`, when given, collects every layer output (post-ReLU) by name.

    emulate_bf16=False is the reference's arithmetic (fp32).  emulate_bf16=True additionally rounds to bf16 exactly
    where the CUDA path stores bf16 (tensor-core kernels' weights, every layer output and its gradient), which turns
    the loose fp32-vs-bf16 comparison into a tight one for the tests; it is not the reference's arithmetic.

    `force` (teacher forcing, a test aid): {layer name: tensor} -- the named layers' outputs are REPLACED by the given
    values (gradients still flow through the layer as computed), so that a backward pass can be compared with another
    implementation's on exactly the same activations and ReLU masks instead of through the mask flips that accumulated
    forward rounding causes.
    """
    rnd0 = _rounding(emulate_bf16)
    current = [None]

    def rnd(t):
        """Rounding of a layer output; under teacher forcing the output then takes the forced value."""
        t = rnd0(t)
        name, current[0] = current[0], None
        if force is not None and name is not None and name in force:
            t = t + (force[name].to(t.dtype) - t).detach()
        return t

    def named(name):
        current[0] = name

    def kern(name, cin):
        # the CUDA path runs the convolutions that read the 3-channel image on CUDA cores with the fp32 kernel
        return weights[name] if cin == 3 else rnd0(weights[name])

    def block(prefix: str, h):
        """train.py:123-143: block_depth x Conv2D(filters, 3, 1, 'same', relu)."""
        for k in range(cfg.block_depth):
            name = f"{prefix}/conv{k}"
            named(name)
            h = rnd(conv3x3(h, kern(f"{name}/kernel", h.shape[-1]), weights[f"{name}/bias"]))
            if taps is not None:
                taps[name] = h
        return h

    def residual(i: int, h):
        """train.py:97-121 around Sequential[DownShuffle, Block, middle, Block, UpShuffle] (train.py:182-190)."""
        named(f"down{i}")
        d = rnd(down_shuffle(h, kern(f"down{i}/kernel", h.shape[-1]), weights[f"down{i}/bias"]))
        if taps is not None:
            taps[f"down{i}"] = d
        d = block(f"block_down{i}", d)
        inner = residual(i + 1, d) if i + 1 < cfg.octaves else block("block_mid", d)
        inner = block(f"block_up{i}", inner)
        named(f"up{i}")
        u = rnd(up_shuffle(inner, rnd0(weights[f"up{i}/kernel"]), weights[f"up{i}/bias"]))
        if taps is not None:
            taps[f"up{i}"] = u
        if cfg.residual:                    # :110-111 input + Dense(input_channels, use_bias=False)(module(input))
            wres = weights[f"res{i}/dense/kernel"]
            if h.shape[-1] == 3:            # (CUDA path: folded into Dense(3) in fp32, the sum is never stored)
                return h + u @ wres
            named(f"res{i}/dense")
            r = rnd(h + u @ rnd0(wres))
            if taps is not None:
                taps[f"res{i}/dense"] = r
            return r
        if cfg.concat:                      # :112-119 module output FIRST
            return torch.cat([u, h], dim=-1)
        return u                            # :120-121

    top = block("block_out", residual(0, block("block_in", x_nhwc)))
    pred = dense(top, weights["dense/kernel"], weights["dense/bias"])
    if taps is not None:
        taps["pred"] = pred
    return pred


def noise_images(x, t_int, eps, cfg: Config = DEFAULT):
    """train.py:224-234. x [B,H,W,3] fp32, t_int [B] int32, eps like x."""
    t = t_int.to(x.dtype)[:, None, None, None]
    a = alpha_dash(t, cfg.steps)
    return x * a ** 0.5 + eps * (1 - a) ** 0.5


def loss_target(x, t_int, eps, pred, cfg: Config = DEFAULT):
    """train.py:238-252: the regression target (and the re-weighted prediction) for the four objective switches."""
    t = t_int.to(x.dtype)[:, None, None, None]
    if cfg.ordinary_differential_equation:                     # :238-242 the noised image of step t-1
        a1 = alpha_dash(t - 1, cfg.steps)
        return x * a1 ** 0.5 + eps * (1 - a1) ** 0.5, pred
    if cfg.predict_x:                                          # :243-244 (the reference's default)
        return x, pred
    target = eps                                               # :245-246
    s = (1 - alpha_dash(t, cfg.steps)) ** 0.5
    if cfg.predict_scaled_epsilon:                             # :247-248
        target = target * s
    if cfg.prediction_weighting:                               # :250-252
        target = target * s
        pred = pred * s
    return target, pred


def trainer_loss(weights, x, t_int, eps, cfg: Config = DEFAULT, taps=None, global_elems: Optional[int] = None,
                 emulate_bf16: bool = False, force=None):
    """Trainer.call (train.py:223-272) -> scalar mean squared error between the target selected by the objective
    switches (train.py:238-252; default predict_x: the clean image) and the prediction.

    global_elems overrides the mean's denominator (data-parallel shards of one global batch)."""
    noised = noise_images(x, t_int, eps, cfg)
    if taps is not None:
        taps["noised"] = noised
    pred = denoiser_forward(weights, noised, cfg, taps, emulate_bf16, force)
    target, pred = loss_target(x, t_int, eps, pred, cfg)
    sq = (target.to(torch.float32) - pred.to(torch.float32)) ** 2
    if global_elems is None:
        return sq.mean()
    return sq.sum() / global_elems


def identity(y_true, y_pred):
    """train.py:171-173."""
    return torch.mean(y_pred)


def loss_and_grads(weights, x, t_int, eps, cfg: Config = DEFAULT, want_taps: bool = False,
                   global_elems: Optional[int] = None, emulate_bf16=False, loss_scale: Optional[float] = None, force=None):
    """One forward+backward (what Keras train_step's GradientTape does, train.py:516): returns
    (loss, {name: grad}, taps) where taps also carries d(loss)/d(layer output) under 'd<name>' when requested.

    loss_scale (train.py:82-83, LossScaleOptimizer.get_scaled_loss / get_unscaled_gradients): the backward pass runs on
    loss * scale -- so every rounding of the emulated 16-bit storage sees the scaled gradient -- and the returned
    gradients (and 'd<name>' taps) have the scale divided out again; non-finite values stay non-finite."""
    ws = {k: v.detach().clone().requires_grad_(True) for k, v in weights.items()}
    taps: Optional[Dict[str, torch.Tensor]] = {} if want_taps else None
    loss = identity(None, trainer_loss(ws, x, t_int, eps, cfg, taps, global_elems, emulate_bf16, force))
    if want_taps:
        for v in taps.values():
            if v.requires_grad:
                v.retain_grad()
    scale = 1.0 if loss_scale is None else float(loss_scale)
    (loss * scale).backward()
    if emulate_bf16 == "f16":
        # under Keras' mixed_float16 policy every layer computes in fp16, so EVERY variable's gradient leaves its op as
        # an fp16 tensor (cast to fp32 afterwards): values beyond 65504 are inf -- the overflow LossScaleOptimizer detects
        grads = {k: v.grad.detach().to(torch.float16).to(torch.float32) / scale for k, v in ws.items()}
    else:
        grads = {k: v.grad.detach() / scale for k, v in ws.items()}
    out_taps: Dict[str, torch.Tensor] = {}
    if want_taps:
        for k, v in taps.items():
            out_taps[k] = v.detach()
            if v.grad is not None:
                out_taps["d" + k] = v.grad.detach() / scale
    return loss.detach(), grads, out_taps


# --------------------------------------------------------------------------------------------- optimizer
def keras_adam_update(w, m, v, g, iteration: int, cfg: Config = DEFAULT):
    """tf.keras.optimizers.Adam (OptimizerV2) dense update, SURVEY.md A.6, in place on fp32 tensors.

    t = iteration+1; lr = WarmUp(iteration); alpha = lr*sqrt(1-b2^t)/(1-b1^t);
    m += (g-m)(1-b1); v += (g^2-v)(1-b2); w -= alpha*m/(sqrt(v)+eps)   (epsilon on the un-corrected sqrt(v))."""
    lr = WarmUp(cfg.base_lr, cfg.warm_up)(iteration)
    t = float(iteration + 1)
    b1p = torch.tensor(cfg.beta1, dtype=torch.float32) ** t
    b2p = torch.tensor(cfg.beta2, dtype=torch.float32) ** t
    alpha = (torch.tensor(lr, dtype=torch.float32) * torch.sqrt(1 - b2p) / (1 - b1p)).item()
    # ResourceApplyAdam receives beta1/beta2 as scalars of the variable dtype and forms (1 - beta) in that dtype:
    # 1 - float32(0.999) = 0.00100004673, not 0.001 (4.7e-5 relative -- visible in v).
    one_minus_b1 = float(torch.tensor(1.0, dtype=w.dtype) - torch.tensor(cfg.beta1, dtype=w.dtype))
    one_minus_b2 = float(torch.tensor(1.0, dtype=w.dtype) - torch.tensor(cfg.beta2, dtype=w.dtype))
    m.add_((g - m) * one_minus_b1)
    v.add_((g * g - v) * one_minus_b2)
    w.sub_(alpha * m / (torch.sqrt(v) + cfg.epsilon))
    return alpha


class DynamicLossScale:
    """tf.keras.mixed_precision.LossScaleOptimizer's dynamic loss scale (train.py:82-83; Keras defaults initial_scale =
    2**15, dynamic_growth_steps = 2000): a step whose gradients are not all finite is skipped and halves the scale
    (never below 1); `growth_steps` consecutive finite steps double it."""

    def __init__(self, initial_scale: float = 2.0 ** 15, growth_steps: int = 2000):
        self.scale = float(initial_scale)
        self.growth_steps = int(growth_steps)
        self.good_steps = 0

    def update(self, finite: bool) -> bool:
        """Book-keeping after one step; returns whether the optimiser update is applied."""
        if finite:
            self.good_steps += 1
            if self.good_steps >= self.growth_steps:
                self.scale *= 2.0
                self.good_steps = 0
            return True
        self.scale = max(self.scale / 2.0, 1.0)
        self.good_steps = 0
        return False


class OracleTrainer:
    """Stateful restatement of `trainer.fit`'s per-step work (train.py:511-523): loss, grads, Adam.

    mixed_precision=True adds what train.py:34,43-45,82-83 switch on: fp16 storage (emulated by rounding where the
    CUDA path stores 16-bit values) and the dynamic LossScaleOptimizer -- skipped steps leave the variables, the Adam
    moments and the iteration count untouched."""

    def __init__(self, cfg: Config = DEFAULT, weights: Optional[Dict[str, torch.Tensor]] = None, seed: int = 0,
                 mixed_precision: bool = False, initial_scale: float = 2.0 ** 15, growth_steps: int = 2000):
        self.cfg = cfg
        self.weights = {k: v.clone() for k, v in (weights or glorot_init(cfg, seed)).items()}
        self.m = {k: torch.zeros_like(v) for k, v in self.weights.items()}
        self.v = {k: torch.zeros_like(v) for k, v in self.weights.items()}
        self.iterations = 0
        self.mixed_precision = mixed_precision
        self.loss_scale = DynamicLossScale(initial_scale, growth_steps) if mixed_precision else None

    def train_step(self, x, t_int, eps) -> float:
        if not self.mixed_precision:
            loss, grads, _ = loss_and_grads(self.weights, x, t_int, eps, self.cfg)
        else:
            loss, grads, _ = loss_and_grads(self.weights, x, t_int, eps, self.cfg, emulate_bf16="f16",
                                            loss_scale=self.loss_scale.scale)
            finite = all(bool(torch.isfinite(g).all()) for g in grads.values())
            if not self.loss_scale.update(finite):
                return float(loss)
        for k in self.weights:
            keras_adam_update(self.weights[k], self.m[k], self.v[k], grads[k], self.iterations, self.cfg)
        self.iterations += 1
        return float(loss)


# --------------------------------------------------------------------------------------------- bf16 emulation
def bf16_round(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def flops_per_image(cfg: Config = DEFAULT) -> Dict[str, float]:
    """FLOP convention of SURVEY.md 8(a): conv 2*Ho*Wo*k*k*Cin*Cout, convT 2*Hin*Win*16*Cin*Cout, dense 2*H*W*Cin*Cout."""
    fwd = 0.0
    no_dgrad = 0.0  # layers reading the image have no data gradient
    n = cfg.octaves

    def extent(name: str) -> int:
        """Output extent of a Conv2D / input extent of a Conv2DTranspose / extent of a Dense."""
        head = name.split("/")[0]
        if head in ("block_in", "block_out", "dense"):
            return cfg.size
        if head == "block_mid":
            return cfg.size >> n
        digits = int("".join(ch for ch in head if ch.isdigit()))
        if head.startswith("res"):
            return cfg.size >> digits
        return cfg.size >> (digits + 1)

    for name, shape in variable_specs(cfg):
        if not name.endswith("kernel"):
            continue
        h = extent(name)
        if len(shape) == 4:
            f = 2.0 * h * h * shape[0] * shape[1] * shape[2] * shape[3]
            if shape[2] == 3 and not name.startswith("up"):
                no_dgrad += f
        else:
            f = 2.0 * h * h * shape[0] * shape[1]
        fwd += f
    bwd = 2 * fwd - no_dgrad
    return {"fwd": fwd, "bwd": bwd, "step": fwd + bwd}
