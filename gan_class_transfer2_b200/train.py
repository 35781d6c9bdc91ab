"""Drop-in host surface for the reference's ``train.py``: same module-level hyper-parameters, the same class
names and constructor signatures (``WarmUp, Residual, Block, UpShuffle, DownShuffle, Denoiser, Trainer``), the
same construction recursion and the same step interface (``trainer(x) -> loss``, ``trainer.compile``,
``trainer.fit``, ``trainer.train_step``) -- with every tensor op routed to the sm_100a kernels behind the C ABI
(include/gct2_b200.h).  The Python loop stays Python; there is no TensorFlow here and no CPU fallback.

The Keras pieces the reference leans on (Layer / Sequential / Dense / Model / Adam / LambdaCallback) are
re-stated minimally in this file; only the behaviour train.py uses is provided.

Reference line numbers cited below are /root/reference/train.py.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Iterable, List, Optional, Sequence

import torch

from . import ops
from .engine import DataParallel, NetConfig, UNetEngine, glorot_uniform, make_engine

# ------------------------------------------------------------------------------------------------ train.py:17-36
size = 256
pixel_size = 128 * 1
max_size = 512 * 1
block_depth = 0
octaves = 6  # bottleneck = 4x4

batch_size = 1
steps = 200

residual = False
concat = True

predict_x = True
predict_scaled_epsilon = False
prediction_weighting = False
ordinary_differential_equation = False

#: train.py:34.  False (the reference's default, fp32 there): bf16 operands / fp32 accumulation and masters here.
#: True: the reference's own reduced-precision mode (Keras 'mixed_float16' policy, train.py:43-45, and the dynamic
#: LossScaleOptimizer, train.py:82-83): fp16 operands and activations, loss scaling with skipped steps on overflow.
mixed_precision = False

warm_up = 2_000
test_step = 25

#: use a CUDA graph for the training step (launch-latency bound at batch 1)
use_cuda_graph = True
#: data-parallel context applied to engines built after it is set (see engine.DataParallel)
data_parallel: Optional[DataParallel] = None


class WarmUp:
    """train.py:50-65: linear warm-up of the learning rate; `step` is the 0-based optimiser iteration."""

    def __init__(self, base, warmup_steps):
        self.base = base
        self.warmup_steps = warmup_steps

    def __call__(self, step):
        if step < self.warmup_steps:
            return self.base * float(step + 1) / (self.warmup_steps + 1)
        return self.base


class Adam:
    """tf.keras.optimizers.Adam as train.py:75 uses it (Keras defaults; epsilon on the un-corrected sqrt(v)).
    The update itself runs in gct2_adam_keras on the engine's flat buffers; this object carries the hyper-parameters."""

    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.learning_rate = learning_rate
        self.beta_1, self.beta_2, self.epsilon = beta_1, beta_2, epsilon

    def schedule(self):
        lr = self.learning_rate
        if isinstance(lr, WarmUp):
            return float(lr.base), int(lr.warmup_steps)
        if callable(lr):
            raise NotImplementedError("only WarmUp schedules (train.py:50-65) or constant learning rates are supported")
        return float(lr), 0


class LossScaleOptimizer:
    """tf.keras.mixed_precision.LossScaleOptimizer as train.py:82-83 uses it (dynamic scaling, Keras defaults): wraps the
    inner optimiser; the scaling itself runs on the device (gct2_loss_scale_check / _update, engine.NetConfig)."""

    def __init__(self, inner_optimizer, dynamic=True, initial_scale=2.0 ** 15, dynamic_growth_steps=2000):
        if not dynamic:
            raise NotImplementedError("train.py:83 uses the default dynamic loss scale")
        self.inner_optimizer = inner_optimizer
        self.initial_scale = float(initial_scale)
        self.dynamic_growth_steps = int(dynamic_growth_steps)


optimizer = Adam(WarmUp(2e-5, warm_up))
regularizer = None
if mixed_precision:  # train.py:82-83
    optimizer = LossScaleOptimizer(optimizer)


def alpha_dash(t):
    """train.py:85-93; works on tensors and Python numbers."""
    t = t / (steps + 1)
    return (1 - t) ** 2 * 0.25


# ------------------------------------------------------------------------------------------------ Keras-shaped shim
class Layer:
    def __init__(self):
        self.built = False

    def build(self, input_shape):
        pass

    def call(self, input):
        raise NotImplementedError

    def __call__(self, input, training=None):
        if not self.built:
            self.build(_shape_of(input))
            self.built = True
        return self.call(input)

    def sublayers(self) -> List["Layer"]:
        return []


def _shape_of(x):
    if isinstance(x, (tuple, list)):
        return [_shape_of(v) for v in x]
    return tuple(x.shape)


def preferred_type():
    """train.py:38: the 16-bit storage format of the active policy."""
    return torch.float16 if mixed_precision else torch.bfloat16


def _check_act(x: torch.Tensor) -> torch.Tensor:
    if not x.is_cuda:
        raise RuntimeError("this layer runs on the B200 only: pass a CUDA tensor (no CPU fallback)")
    return x if x.dtype == preferred_type() else x.to(preferred_type())


class Sequential(Layer):
    def __init__(self, layers: Sequence[Layer] = ()):
        super().__init__()
        self.layers = list(layers)

    def call(self, input):
        for layer in self.layers:
            input = layer(input)
        return input

    def sublayers(self):
        return self.layers


#: one advancing generator for layers that are built standalone (outside a Denoiser): Keras draws every glorot kernel
#: independently, so two layers of equal shape must not start from equal kernels
_init_generator: Optional[torch.Generator] = None


def _next_init_generator() -> torch.Generator:
    global _init_generator
    if _init_generator is None:
        _init_generator = torch.Generator().manual_seed(torch.initial_seed() % (2 ** 31))
    return _init_generator


class _ConvLayer(Layer):
    """Common part of the two stride-2 4x4 layers: lazily created glorot-uniform kernel in the Keras layout, zero bias
    (train.py:149,162), bf16 shadow for the tensor cores."""
    transposed = False
    ksize = 4

    def __init__(self, filters):
        super().__init__()
        self.filters = filters
        self.kernel: Optional[torch.Tensor] = None
        self.bias: Optional[torch.Tensor] = None
        self._k16: Optional[torch.Tensor] = None
        self._ws: Optional[ops.Workspace] = None

    def __call__(self, input, training=None):
        if not input.is_cuda:
            raise RuntimeError("this layer runs on the B200 only: pass a CUDA tensor (no CPU fallback)")
        return super().__call__(input)

    def build(self, input_shape):
        cin = input_shape[-1]
        k = self.ksize
        shape = (k, k, self.filters, cin) if self.transposed else (k, k, cin, self.filters)
        if self.kernel is None:
            self.kernel = glorot_uniform(shape, _next_init_generator()).cuda()
            self.bias = torch.zeros(self.filters, device="cuda")

    def _shadow(self):
        if self._k16 is None or self._k16.shape != self.kernel.shape or self._k16.dtype != preferred_type():
            self._k16 = torch.empty_like(self.kernel, dtype=preferred_type())
        ops.cast_bf16(self.kernel.reshape(-1), self._k16.reshape(-1))
        return self._k16

    def _workspace(self, nbytes, device):
        if self._ws is None or self._ws.nbytes < nbytes:
            self._ws = ops.Workspace(nbytes, device)
        return self._ws


class UpShuffle(_ConvLayer):
    """train.py:145-156: Conv2DTranspose(filters, 4, 2, 'same', glorot_uniform, relu)."""
    transposed = True

    def call(self, input):
        x = _check_act(input)
        B, H, W, _ = x.shape
        y = torch.empty(B, 2 * H, 2 * W, self.filters, dtype=preferred_type(), device=x.device)
        return ops.convT4s2_fprop(x, self._shadow(), self.bias, y, self._workspace(4 * y.numel(), x.device))


class DownShuffle(_ConvLayer):
    """train.py:158-169: Conv2D(filters, 4, 2, 'same', glorot_uniform, relu)."""

    def call(self, input):
        B, H, W, C = input.shape
        y = torch.empty(B, H // 2, W // 2, self.filters, dtype=preferred_type(), device=input.device)
        if C == 3:
            if not input.is_cuda:
                raise RuntimeError("this layer runs on the B200 only: pass a CUDA tensor (no CPU fallback)")
            return ops.conv4s2_c3_fprop(input.float().contiguous(), self.kernel, self.bias, y)
        return ops.conv4s2_fprop(_check_act(input), self._shadow(), self.bias, y,
                                 self._workspace(4 * y.numel(), input.device))


class Conv3x3(_ConvLayer):
    """train.py:132-137: Conv2D(filters, 3, 1, 'same', glorot_uniform, relu) -- the layer a Block stacks."""
    ksize = 3

    def call(self, input):
        B, H, W, C = input.shape
        y = torch.empty(B, H, W, self.filters, dtype=preferred_type(), device=input.device)
        if C == 3:
            return ops.conv3s1_c3_fprop(input.float().contiguous(), self.kernel, self.bias, y)
        return ops.conv3s1_fprop(_check_act(input), self._shadow(), self.bias, y, self._workspace(4 * y.numel(), input.device))


class Block(Layer):
    """train.py:123-143: block_depth x Conv2D(filters, 3, 1, 'same', relu).  block_depth is 0 in the reference
    (train.py:20), which makes this an empty Sequential: identity, no variables.  The depth is read when the Block is
    constructed (inside a Denoiser the engine runs the convolutions; standalone the layers run one launch each)."""

    def __init__(self, filters):
        super().__init__()
        self.filters = filters
        self.depth = int(block_depth)
        self.module = Sequential([Conv3x3(filters) for _ in range(self.depth)])

    def build(self, input_shape):
        pass  # the convolutions build themselves at first use (their input channels are inferred, train.py:139)

    def call(self, input):
        return self.module(input)

    def sublayers(self):
        return [self.module]


class Residual(Layer):
    """train.py:97-121 with the reference's flags (residual=False, concat=True): concat([module(x), highway(x)], -1),
    module output first."""

    def __init__(self, module, highway=lambda x: x):
        super().__init__()
        self.module = module
        self.highway = highway

    def build(self, input_shape):
        if residual:  # train.py:106-108
            self.dense = Dense(input_shape[-1], use_bias=False)

    def call(self, input):
        if residual:  # train.py:110-111 (standalone use; inside a Denoiser the engine fuses the add into the projection)
            out = self.module(input)
            if not self.dense.built:
                self.dense.build(tuple(out.shape))
                self.dense.built = True
            k = self.dense.kernel.to(out.dtype)
            y = torch.empty(*out.shape[:3], k.shape[1], dtype=out.dtype, device=out.device)
            return ops.conv3s1_fprop_add(out, k.view(1, 1, *k.shape), _check_act(input), y)
        if concat:
            out = self.module(input)
            return torch.cat([out, self.highway(input).to(out.dtype)], -1)
        return self.module(input)

    def sublayers(self):
        return [self.module]


class Dense(Layer):
    """tf.keras.layers.Dense as train.py:198-202 uses it: contraction of the last axis, bias, no activation."""

    def __init__(self, units, use_bias=True):
        super().__init__()
        self.units = units
        self.use_bias = use_bias
        self.kernel: Optional[torch.Tensor] = None
        self.bias: Optional[torch.Tensor] = None

    def build(self, input_shape):
        if self.kernel is None:
            self.kernel = glorot_uniform((input_shape[-1], self.units), _next_init_generator()).cuda()
            self.bias = torch.zeros(self.units, device="cuda") if self.use_bias else None


def identity(y_true, y_pred):
    """train.py:171-173."""
    return torch.mean(y_pred)


class LambdaCallback:
    def __init__(self, on_epoch_begin: Optional[Callable] = None, on_epoch_end: Optional[Callable] = None):
        self.on_epoch_begin = on_epoch_begin
        self.on_epoch_end = on_epoch_end


class Model(Layer):
    def compile(self, optimizer, loss):
        self.optimizer = optimizer
        self.compiled_loss = loss

    def train_step(self, data) -> Dict[str, torch.Tensor]:
        raise NotImplementedError

    def fit(self, dataset: Iterable, steps_per_epoch: int, epochs: int, callbacks: Sequence[LambdaCallback] = (),
            verbose: int = 0):
        """The Python loop of train.py:516-523: `dataset` yields (image, image); returns {'loss': [per-epoch mean]}."""
        it = iter(dataset)
        history: Dict[str, List[float]] = {"loss": []}
        logs: Dict[str, float] = {}
        for epoch in range(epochs):
            for cb in callbacks:
                if cb.on_epoch_begin:
                    cb.on_epoch_begin(epoch, logs)
            total = torch.zeros(1, device="cuda")
            for _ in range(steps_per_epoch):
                total += self.train_step(next(it))["loss"]
            logs = {"loss": float(total.item()) / steps_per_epoch}  # one host sync per epoch
            history["loss"].append(logs["loss"])
            if verbose:
                print(f"epoch {epoch + 1}/{epochs} - loss: {logs['loss']:.6f}")
            for cb in callbacks:
                if cb.on_epoch_end:
                    cb.on_epoch_end(epoch, logs)
        return history


# ------------------------------------------------------------------------------------------------ models
class Denoiser(Model):
    """train.py:175-215: the recursive concat-skip U-Net plus Dense(3).  Construction mirrors the reference line by
    line; execution goes through one UNetEngine per batch size (concat buffers, fused epilogues, flat variables)."""

    def __init__(self):
        super().__init__()
        self.middle = Block(min(pixel_size * 2 ** octaves, max_size))
        for i in reversed(range(octaves)):
            filters = min(pixel_size * 2 ** i, max_size)
            self.middle = Residual(
                Sequential([
                    DownShuffle(filters),
                    Block(filters),
                    self.middle,
                    Block(filters),
                    UpShuffle(min(pixel_size * 2 ** i // 2, max_size)),
                ])
            )
        self.middle = Sequential([
            Block(pixel_size),
            self.middle,
            Block(pixel_size),
            Dense(3),
        ])
        self._engines: Dict[tuple, UNetEngine] = {}
        self._seed = 0
        self._pending_weights: Optional[Dict[str, torch.Tensor]] = None
        self._optimizer: Optional[Adam] = None  # set by Trainer.compile; None -> the module-level `optimizer`

    # -- structure ---------------------------------------------------------------------------------------------
    def _walk(self):
        """Pattern-matches the layer tree against the one shape the fused engine implements and returns
        (down layers outer->inner, up layers outer->inner, dense)."""
        outer = self.middle.layers
        if not (len(outer) == 4 and isinstance(outer[0], Block) and isinstance(outer[1], Residual)
                and isinstance(outer[2], Block) and isinstance(outer[3], Dense) and outer[3].units == 3):
            raise NotImplementedError("Denoiser.middle must be Sequential[Block, Residual, Block, Dense(3)] (train.py:191-204)")
        downs, ups = [], []
        node = outer[1]
        while isinstance(node, Residual):
            seq = node.module.layers if isinstance(node.module, Sequential) else None
            if not (seq and len(seq) == 5 and isinstance(seq[0], DownShuffle) and isinstance(seq[1], Block)
                    and isinstance(seq[3], Block) and isinstance(seq[4], UpShuffle)):
                raise NotImplementedError("each Residual must wrap Sequential[DownShuffle, Block, inner, Block, UpShuffle] "
                                          "(train.py:182-190)")
            downs.append(seq[0])
            ups.append(seq[4])
            node = seq[2]
        if not isinstance(node, Block):
            raise NotImplementedError("the innermost module must be a Block (train.py:179)")
        return downs, ups, outer[3]

    def _blocks(self) -> Dict[str, Block]:
        """Every Block of the recursion by the variable prefix the engine uses for it."""
        outer = self.middle.layers
        blocks = {"block_in": outer[0], "block_out": outer[2]}
        node, i = outer[1], 0
        while isinstance(node, Residual):
            seq = node.module.layers
            blocks[f"block_down{i}"], blocks[f"block_up{i}"] = seq[1], seq[3]
            node, i = seq[2], i + 1
        blocks["block_mid"] = node
        return blocks

    def net_config(self, image_size: int) -> NetConfig:
        downs, ups, _ = self._walk()
        opt = self._optimizer if self._optimizer is not None else optimizer
        extra = {}
        if isinstance(opt, LossScaleOptimizer):
            extra = dict(loss_scale_init=opt.initial_scale, loss_scale_growth=opt.dynamic_growth_steps)
            opt = opt.inner_optimizer
        base_lr, warm = opt.schedule()
        blocks = self._blocks()
        depths = {b.depth for b in blocks.values()}
        if len(depths) != 1:
            raise NotImplementedError("all Blocks of a Denoiser must have the same depth (train.py:20 is one global)")
        depth = depths.pop()
        if depth:
            for i, d in enumerate(downs):
                if blocks[f"block_down{i}"].filters != d.filters or blocks[f"block_up{i}"].filters != d.filters:
                    raise NotImplementedError("the Blocks of a level must have the level's DownShuffle filters (train.py:184-187)")
            if blocks["block_in"].filters != blocks["block_out"].filters:
                raise NotImplementedError("the two outermost Blocks must have equal filters (train.py:192,194)")
            extra.update(block_depth=depth, mid_filters=blocks["block_mid"].filters, outer_filters=blocks["block_in"].filters)
        extra.update(concat=bool(concat), residual=bool(residual))
        return NetConfig(size=image_size, pixel_size=pixel_size, max_size=max_size, octaves=len(downs), steps=steps,
                         warm_up=warm, base_lr=base_lr, beta1=opt.beta_1, beta2=opt.beta_2,
                         epsilon=opt.epsilon, down_filters=tuple(d.filters for d in downs),
                         up_filters=tuple(u.filters for u in ups), mixed_precision=bool(mixed_precision),
                         target_mode=ops.target_mode(predict_x, predict_scaled_epsilon, prediction_weighting,
                                                     ordinary_differential_equation), **extra)

    def use_optimizer(self, opt: "Adam") -> None:
        """The optimiser the engines are built with (Trainer.compile hands over the one it was given, train.py:511-514)."""
        inner = opt.inner_optimizer if isinstance(opt, LossScaleOptimizer) else opt
        if not isinstance(inner, Adam):
            raise NotImplementedError("only tf.keras.optimizers.Adam (train.py:75), optionally inside a LossScaleOptimizer "
                                      "(train.py:83), is implemented on the device")
        if self._engines:
            cur = self._optimizer if self._optimizer is not None else optimizer
            cur = cur.inner_optimizer if isinstance(cur, LossScaleOptimizer) else cur
            same = (cur.schedule() == inner.schedule() and (cur.beta_1, cur.beta_2, cur.epsilon) ==
                    (inner.beta_1, inner.beta_2, inner.epsilon))
            if not same:
                raise RuntimeError("the training engine has already been built with another optimiser; compile() before "
                                   "the first step")
        self._optimizer = opt

    def engine(self, batch: int, image_size: int) -> UNetEngine:
        key = (batch, image_size)
        if key not in self._engines:
            first = next(iter(self._engines.values()), None)
            eng = make_engine(self.net_config(image_size), batch, dp=data_parallel, use_graph=use_cuda_graph,
                              share_params_with=first)
            if first is None:
                if self._pending_weights is not None:
                    eng.load_weights(self._pending_weights)
                    self._pending_weights = None
                else:
                    eng.init_glorot(self._seed)
                self._bind_variables(eng)
            self._engines[key] = eng
        return self._engines[key]

    def sample(self, x_theta, epsilon_theta, t_values):
        """The diffusion loops of log_sample (train.py:365-413: `reversed(range(steps, 0, -1))`, batch 1;
        train.py:439-479: `range(steps, 0, -1)`, batch 6) for the objective the module's switches select (predict_x by
        default; predicted / scaled noise and the ODE form as in train.py:382-413): every iteration mixes `fake`, calls
        the denoiser and re-derives (x_theta, epsilon_theta).  Returns the final pair as fp32 NHWC device tensors.  One
        CUDA graph per (batch, schedule)."""
        B, H, W, C = x_theta.shape
        if C != 3 or H != W or tuple(epsilon_theta.shape) != tuple(x_theta.shape):
            raise ValueError("sample expects two square NHWC tensors [B,S,S,3] of equal shape")
        xt, et = self.engine(B, H).sample(x_theta, epsilon_theta, t_values)
        return xt.clone(), et.clone()

    def _bind_variables(self, eng: UNetEngine) -> None:
        """Layers own their variables in Keras; here they are views into the engine's flat fp32 buffer."""
        downs, ups, dense = self._walk()
        for i, layer in enumerate(downs):
            layer.kernel, layer.bias = eng.view(eng.w, f"down{i}/kernel"), eng.view(eng.w, f"down{i}/bias")
            layer.built = True
        for i, layer in enumerate(ups):
            layer.kernel, layer.bias = eng.view(eng.w, f"up{i}/kernel"), eng.view(eng.w, f"up{i}/bias")
            layer.built = True
        dense.kernel, dense.bias = eng.view(eng.w, "dense/kernel"), eng.view(eng.w, "dense/bias")
        dense.built = True
        node, i = self.middle.layers[1], 0
        while isinstance(node, Residual):  # train.py:106-108: the projections of residual = True
            if eng.cfg.residual:
                node.dense = Dense(eng.cfg.level_in(i), use_bias=False)
                node.dense.kernel, node.dense.built = eng.view(eng.w, f"res{i}/dense/kernel"), True
            node.built = True
            node, i = node.module.layers[2], i + 1
        for prefix, block in self._blocks().items():
            for k, conv in enumerate(block.module.layers):
                conv.kernel, conv.bias = eng.view(eng.w, f"{prefix}/conv{k}/kernel"), eng.view(eng.w, f"{prefix}/conv{k}/bias")
                conv.built = True
            block.built = True

    # -- Keras-like surface ------------------------------------------------------------------------------------
    def set_seed(self, seed: int) -> None:
        self._seed = seed

    def set_weights(self, weights: Dict[str, torch.Tensor]) -> None:
        """Weights by Keras variable name ('down0/kernel', ..., 'dense/bias') in the Keras layouts."""
        if self._engines:
            next(iter(self._engines.values())).load_weights(weights)
            for e in self._engines.values():
                e._graph = None  # captured graphs stay valid (same buffers); dropped only to be safe
        else:
            self._pending_weights = {k: v.clone() for k, v in weights.items()}

    @property
    def trainable_variables(self) -> List[torch.Tensor]:
        eng = next(iter(self._engines.values()))
        return [eng.view(eng.w, name) for name, _ in eng.specs]

    def call(self, input):
        x, t = input  # t is ignored by the reference as well (train.py:207-210)
        B, H, W, C = x.shape
        if C != 3 or H != W:
            raise ValueError("Denoiser expects square NHWC images with 3 channels")
        # a fresh tensor per call, like Keras (the engine's own buffer is overwritten by the next call / graph replay)
        return self.engine(B, H).denoise(x.to("cuda", torch.float32)).clone()

    def __call__(self, input, training=None):
        return self.call(input)


class Trainer(Model):
    """train.py:217-280: the diffusion training objective around a Denoiser."""

    def __init__(self, denoiser):
        super().__init__()
        self.denoiser = denoiser
        self.optimizer = optimizer
        self.compiled_loss = identity

    def compile(self, optimizer, loss):
        """train.py:511-514.  The optimiser given here -- not the module-level one -- parameterises the device update."""
        super().compile(optimizer, loss)
        self.denoiser.use_optimizer(optimizer)

    def _engine_for(self, x) -> UNetEngine:
        B, H, W, C = x.shape
        if C != 3 or H != W:
            raise ValueError("Trainer expects square NHWC images with 3 channels")
        return self.denoiser.engine(B, H)

    def call(self, x, t_int=None, epsilon=None):
        """Scalar loss of one forward pass (train.py:223-272).  t_int / epsilon may be injected for parity tests;
        by default they are drawn on the device exactly where the reference draws them (train.py:224-227)."""
        eng = self._engine_for(x)
        eng.set_batch(x, t_int, epsilon)
        if t_int is None:
            eng.t_int.random_(1, eng.cfg.steps + 1)
            eng.eps.normal_()
        eng.loss.zero_()
        ops.noise_images(eng.x, eng.eps, eng.t_int, eng.noised, eng.cfg.steps)  # forward-only call: torch's draws
        eng._forward(want_pred=True, backward=False, inv_n=1.0 / (eng.global_batch * eng.cfg.size ** 2 * 3))
        return eng.loss[0].clone()

    def __call__(self, x, training=None, **kw):
        return self.call(x, **kw)

    def train_step(self, data, t_int=None, epsilon=None):
        """What Keras' Model.train_step does for train.py:516: loss, gradients, Adam update.  Returns {'loss': device
        scalar}; nothing synchronises with the host."""
        x = data[0] if isinstance(data, (tuple, list)) else data
        eng = self._engine_for(x)
        if x.dtype == torch.uint8:
            # the dataset stops one op short of decode_file's cast (train.py:292): bytes in, /128 - 1 on the device
            if t_int is not None or epsilon is not None:
                raise ValueError("injected t_int / epsilon need the float32 batch")
            loss = eng.train_step_u8(x)[0]
        else:
            loss = eng.train_step(x, t_int, epsilon)[0]
        # a fresh scalar per step, like Keras: the engine's loss buffer is overwritten by the next replay (one 4-byte
        # device-to-device copy; a list of per-step losses keeps its values)
        loss = loss.clone()
        # identity (train.py:171-173) is the mean of an already-scalar loss: skip the extra launch
        return {"loss": loss if self.compiled_loss is identity else self.compiled_loss(None, loss)}


# ------------------------------------------------------------------------------------------------ log_sample
#: train.py:305-311: the fixed evaluation tensors.  The reference reads `example_image` from disk at import; here the
#: caller assigns them (fp32 device tensors): example_image [1,S,S,3] in [-1,1), example [1,2,S,S,3] ~ N(0,1),
#: dictionary [S,S,2**bits_per_pixel,3] ~ N(0,1).  None -> log_sample draws synthetic ones once.
example_image: Optional[torch.Tensor] = None
example: Optional[torch.Tensor] = None
dictionary: Optional[torch.Tensor] = None
bits_per_pixel = 3


def log_sample(epochs, logs, denoiser=None):
    """train.py:323-496 without the TensorBoard writer: returns what the reference logs -- {'example loss': float,
    'denoised', 'step_1', 'step_0.25', 'step_0.5', 'step_0.75', 'fake': images (x*0.5+0.5)} -- computed on the device:
      :325-361  one denoising of the example at test_step and its RMSE (gct2_rmse);
      :365-413  the forward (inversion) loop t = 1..steps at batch 1 (Denoiser.sample: one CUDA graph);
      :418-432  the latent edits (gct2_latent_edits) + the two fixed noise samples -> batch 6;
      :439-496  the backward loop t = steps..1 at batch 6, with x_theta snapshots at t = steps, 3/4, 1/2, 1/4 steps."""
    global example_image, example, dictionary
    den = denoiser if denoiser is not None else __getattr__("denoiser")
    S = size
    dev = torch.device("cuda", torch.cuda.current_device())
    if example_image is None or example is None or dictionary is None:
        g = torch.Generator().manual_seed(1234)
        example_image = (torch.randint(0, 256, (1, S, S, 3), generator=g).float() / 128 - 1).to(dev)
        example = torch.randn(1, 2, S, S, 3, generator=g).to(dev)
        dictionary = torch.randn(S, S, 2 ** bits_per_pixel, 3, generator=g).to(dev)
    out: Dict[str, object] = {}
    # ---- :325-361
    f = alpha_dash(float(test_step))
    if ordinary_differential_equation:
        f = alpha_dash(steps / 2) ** 0.5
    noised = example_image * f ** 0.5 + example[0, :1] * (1 - f) ** 0.5
    pred = den((noised, None))
    if ordinary_differential_equation:
        h, h1 = alpha_dash(steps / 2), alpha_dash(steps / 2 - 1)
        denoised = (pred * (1 - h) ** 0.5 - noised * (1 - h1) ** 0.5) / (h1 ** 0.5 * (1 - h) ** 0.5 - h ** 0.5 * (1 - h1) ** 0.5)
    elif predict_x:
        denoised = pred
    else:
        p = pred if predict_scaled_epsilon else pred * (1 - f) ** 0.5
        denoised = (noised - p) / f ** 0.5
    loss_t = torch.zeros(1, device=dev)
    ops.rmse(example_image, denoised, loss_t)
    out["denoised"] = denoised * 0.5 + 0.5
    # ---- :365-413 forward diffusion (inversion): x_theta = epsilon_theta = the example, t ascending
    _, eps_theta = den.sample(example_image, example_image, list(range(1, steps + 1)))
    # ---- :418-434
    edits = torch.empty(4, S, S, 3, device=dev)
    ops.latent_edits(eps_theta, dictionary, edits)
    fake = torch.cat([example[0], edits], 0)   # [6,S,S,3]
    # ---- :439-496 backward diffusion at batch 6, snapshots of x_theta
    marks = {steps: "step_1", 3 * steps // 4: "step_0.75", 2 * steps // 4: "step_0.5", steps // 4: "step_0.25"}
    x_theta, e_theta = fake, fake
    sched = list(range(steps, 0, -1))
    cuts = sorted({sched.index(t) + 1 for t in marks if t in sched} | {len(sched)})
    start = 0
    for cut in cuts:   # one captured graph per segment; a snapshot is taken right after the step it belongs to
        x_theta, e_theta = den.sample(x_theta, e_theta, sched[start:cut])
        t_done = sched[cut - 1]
        if t_done in marks:
            out[marks[t_done]] = x_theta * 0.5 + 0.5
        start = cut
    out["fake"] = x_theta * 0.5 + 0.5
    out["example loss"] = float(loss_t)
    return out


_singletons: Dict[str, object] = {}


def __getattr__(name):
    # train.py:282-283 creates `denoiser` and `trainer` at import; here they are created on first access so that
    # importing this module never touches the GPU.
    if name in ("denoiser", "trainer"):
        if "denoiser" not in _singletons:
            _singletons["denoiser"] = Denoiser()
            _singletons["trainer"] = Trainer(_singletons["denoiser"])
        return _singletons[name]
    raise AttributeError(name)
