"""Step engine for the reference's dormant wiring switches.

``train.py:20`` ``block_depth``, ``train.py:26`` ``residual`` and ``train.py:27`` ``concat`` change what
``Denoiser.__init__`` (train.py:175-204) builds: with ``residual = True`` a ``Residual`` returns ``input +
Dense(input_channels, use_bias=False)(module(input))`` (train.py:106-111; a 1x1 case of the stride-1 map with the add fused
into its epilogue, and at the image level -- 3 channels -- folded into the Dense(3) kernel); with ``block_depth = d > 0`` every ``Block`` of the recursion is a stack of d ``Conv2D(filters, 3, 1, 'same',
relu)`` (train.py:131-139) -- one on the image, one behind every DownShuffle, one in the middle, one in front of
every UpShuffle, one in front of Dense(3) --, and with ``concat = False`` a ``Residual`` is just its module (no skip).

The tuned ``UNetEngine`` is specialised for the defaults (fixed launch sequence, three chains, buckets).  This engine
runs any of the other combinations from a *layer list* built by the same recursion as the reference's constructor:

    forward    every layer's fprop in list order (bias + ReLU fused, output written into its channel slice)
    backward   the list in reverse: weight gradient, then data gradient (ReLU mask of the producer and the add of a
               skip gradient fused, exactly as in the default engine); all bias gradients in one launch; Keras-Adam
               over the flat buffers in one launch

with the same kernels: the 4x4 / stride-2 tensor-core family, its stride-1 index maps (gct2_conv3s1_*), the CUDA-core
convolutions on the 3-channel image, the fused Dense(3)+MSE.  Concatenations still never exist: ``cat[j]`` is one buffer
per level whose two channel ranges are written by their producers.  Everything public (train_step, run_step with CUDA
graphs, loss_and_grads, sample, denoise, weights / grads, mixed precision) is inherited from ``UNetEngine``; data
parallelism is not offered here.

Gradient convention (as in engine.py): a gradient buffer holds d(loss)/d(pre-activation) of the layer that produced
the activation, i.e. the ReLU mask is applied by whoever *writes* the gradient -- a consumer's dgrad epilogue masks by
the activation it reads as input.  A tensor with two consumers (the skip: DownShuffle and the concat reader) gets the
reader's part stored raw first and is completed (+=, mask) by the DownShuffle's dgrad.
"""
from __future__ import annotations

import dataclasses
from typing import Callable, List, Optional

import torch

from . import ops
from .engine import NetConfig, UNetEngine


@dataclasses.dataclass
class _Layer:
    kind: str                       # "image3" | "s1" | "down" | "down_image" | "up" | "proj"
    name: str                       # variable prefix ("<name>/kernel", "<name>/bias")
    x: torch.Tensor                 # input view (fp32 image for the two image kinds)
    y: torch.Tensor                 # output view (post-ReLU)
    gy: torch.Tensor                # gradient w.r.t. this layer's pre-activation (same geometry as y)
    gx: Optional[torch.Tensor]      # where the input's gradient goes (None: the input is the image)
    mask: int = 0                   # leading channels of gx that are ReLU-masked by x (the rest is a raw skip part)
    add_old: bool = False           # gx already holds the other consumer's part
    res: Optional[torch.Tensor] = None  # "proj": the Residual's input, added in the epilogue (train.py:110-111)


def build_layer_list(cfg: NetConfig, image, pair: Callable, alloc_like: Callable):
    """The layer list of a configuration, by the recursion of the reference's constructor (train.py:175-204).

    Works on anything buffer-shaped: `image` stands for the fp32 noised image, `pair(C, H)` returns a new 16-bit
    activation buffer [B,H,H,C] and its gradient twin, `alloc_like(buf)` one more buffer of buf's geometry; buffers are
    sliced as buf[..., a:b] and asked for .shape and .dtype.  The engine passes CUDA tensors, the CPU tests symbolic
    buffers (tests/test_block_wiring.py checks the forward and backward data flow of every switch combination).
    Returns (layers, cat, gcat, dense_in, gdense_in)."""
    n, d, S = cfg.octaves, cfg.block_depth, cfg.size
    layers: List[_Layer] = []
    joined = cfg.concat and not cfg.residual  # the skip travels as a channel slice of the level's buffer
    # one buffer per Residual level: [up_j output | skip] (train.py:113-119), or the up output alone
    cat, gcat = {}, {}
    for j in range(n):
        if (j == 0 and d == 0) or not joined:
            C = cfg.up_c(j)  # (at the image level the skip is not 16-bit data: Dense reads the image separately)
        else:
            C = cfg.res_out(j)
        cat[j], gcat[j] = pair(C, S >> j)

    def is_image(t) -> bool:
        return t.dtype == torch.float32

    def skip_view(j: int):
        """Where the tensor entering level j lives (and its gradient)."""
        if joined and not (j == 0 and d == 0):
            return cat[j][..., cfg.up_c(j):], gcat[j][..., cfg.up_c(j):]
        return pair(cfg.level_in(j), S >> j)

    def add(kind, name, x, gx, y, gy, mask=None, add_old=False, res=None):
        layers.append(_Layer(kind, name, x, y, gy, gx, x.shape[3] if mask is None else mask, add_old, res))

    def block(prefix, x, gx, filters, H, last=None, first_mask=None):
        """d stride-1 convs; the last one writes into `last` (a view pair) when given.  Returns the output pair."""
        for k in range(d):
            y, gy = last if (k == d - 1 and last is not None) else pair(filters, H)
            if is_image(x):
                add("image3", f"{prefix}/conv{k}", x, None, y, gy)
            else:
                add("s1", f"{prefix}/conv{k}", x, gx, y, gy, mask=first_mask if k == 0 else None)
            x, gx = y, gy
        return x, gx

    def level(i: int, h, gh):
        """Residual level i on the tensor h (train.py:180-190); returns the level's output pair."""
        H = S >> (i + 1)
        nxt = skip_view(i + 1) if i + 1 < n else pair(cfg.down_c(i), H)
        dy, gdy = nxt if d == 0 else pair(cfg.down_c(i), H)
        if is_image(h):
            add("down_image", f"down{i}", h, None, dy, gdy)
        else:
            # with a skip (concat) or identity (residual) path h has a second consumer whose raw part is already in gh
            add("down", f"down{i}", h, gh, dy, gdy, add_old=cfg.concat or cfg.residual)
        hy, ghy = block(f"block_down{i}", dy, gdy, cfg.down_c(i), H, last=nxt)
        if i + 1 < n:
            inner, ginner = level(i + 1, hy, ghy)
            # what the consumer of level i+1's output may ReLU-mask: the up part of a concat buffer (its skip part is
            # stored raw), nothing of a residual sum (not a ReLU output), everything otherwise
            inner_mask = 0 if cfg.residual else (cfg.up_c(i + 1) if cfg.concat else None)
        else:
            inner, ginner = block("block_mid", hy, ghy, cfg.mid_c(), H)
            inner_mask = None
        if d:
            u, gu = block(f"block_up{i}", inner, ginner, cfg.down_c(i), H, first_mask=inner_mask)
            inner_mask = None
        else:
            u, gu = inner, ginner
        out, gout = cat[i][..., :cfg.up_c(i)], gcat[i][..., :cfg.up_c(i)]
        add("up", f"up{i}", u, gu, out, gout, mask=inner_mask)
        if cfg.residual:
            if is_image(h):
                return out, gout  # image level: the projection onto 3 channels is folded into Dense(3) (_forward)
            # r = h + Dense(C, use_bias=False)(up_i output).  The gradient of r IS the identity path's part of h's
            # gradient: r's gradient buffer is h's, and down_i's dgrad completes it in place (add_old)
            r = alloc_like(h)
            add("proj", f"res{i}/dense", out, gout, r, gh, res=h)
            return r, gh
        return cat[i], gcat[i]

    if d:
        h0, gh0 = block("block_in", image, None, cfg.outer_c(), S, last=skip_view(0))
    else:
        h0, gh0 = image, None
    top, gtop = level(0, h0, gh0)
    if d:
        top_mask = 0 if cfg.residual else (cfg.up_c(0) if cfg.concat else None)
        top, gtop = block("block_out", top, gtop, cfg.outer_c(), S, first_mask=top_mask)
    return layers, cat, gcat, top, gtop


class BlockUNetEngine(UNetEngine):
    def __init__(self, cfg: NetConfig, batch: int, device=None, dp=None, use_graph: bool = False,
                 share_params_with: Optional[UNetEngine] = None):
        if dp is not None and dp.world > 1:
            raise NotImplementedError("the layer-list engine (block_depth > 0 / concat = False) is single-GPU")
        super().__init__(cfg, batch, device=device, dp=None, use_graph=use_graph, share_params_with=share_params_with)
        # the optimiser runs on the main stream here: a conv launch's prologue could overlap its tail (programmatic
        # dependent launch), so the kernels may not be fetched ahead of the dependency (GCT2_WEIGHTS_STABLE)
        self.weights_stable = False

    # ------------------------------------------------------------------------------------------ buffers + layer list
    def _alloc_activations(self) -> None:
        cfg, dev, B = self.cfg, self.device, self.B
        self._bufs: List[torch.Tensor] = []

        def pair(C: int, H: int):
            a = torch.zeros(B, H, H, C, dtype=self.half, device=dev)
            g = torch.zeros_like(a)
            self._bufs += [a, g]
            return a, g

        def alloc_like(t):
            r = torch.zeros_like(t)
            self._bufs.append(r)
            return r

        self.layers, self.cat, self.gcat, self.dense_in, self.gdense_in = build_layer_list(cfg, self.noised, pair, alloc_like)
        biggest = max(t.numel() for t in self._bufs)
        self.ws = ops.Workspace(max(4 * 4 * biggest, 64 << 20), dev)
        self.ws_w = self.ws  # one stream: the weight gradients share the scratch
        self._buckets = []
        # BiasAddGrad of every layer: gct2_bias_grad_multi takes up to 16 tensors per launch
        biased = [l for l in self.layers if l.kind != "proj"]
        self._bias_plans = [ops.BiasGradPlan([l.gy for l in chunk], [self.view(self.g, f"{l.name}/bias") for l in chunk])
                            for chunk in (biased[i:i + 16] for i in range(0, len(biased), 16))]
        self._res0 = cfg.residual and cfg.block_depth == 0
        if self._res0:
            U = cfg.up_c(0)
            self._weff = torch.zeros(U + 3, 3, dtype=torch.float32, device=dev)
            self._dweff = torch.zeros(U + 3, 3, dtype=torch.float32, device=dev)

    # ------------------------------------------------------------------------------------------ the step's two halves
    def plan_keys(self) -> List[str]:
        return []  # the plan table of tools/tune_plans.py belongs to the default wiring

    def _proj_kernel(self, l: _Layer, buf: torch.Tensor) -> torch.Tensor:
        """Dense kernel [Cin, Cout] of a residual projection as the [1,1,Cin,Cout] kernel of the stride-1 map."""
        k = self.view(buf, f"{l.name}/kernel")
        return k.view(1, 1, *k.shape)

    def _fprop(self, l: _Layer) -> None:
        if l.kind == "proj":
            ops.conv3s1_fprop_add(l.x, self._proj_kernel(l, self.w16), l.res, l.y, self.weights_stable)
            return
        bias = self.view(self.w, f"{l.name}/bias")
        if l.kind == "image3":
            ops.conv3s1_c3_fprop(l.x, self.view(self.w, f"{l.name}/kernel"), bias, l.y)
        elif l.kind == "down_image":
            ops.conv4s2_c3_fprop(l.x, self.view(self.w, f"{l.name}/kernel"), bias, l.y)
        elif l.kind == "s1":
            ops.conv3s1_fprop(l.x, self.view(self.w16, f"{l.name}/kernel"), bias, l.y, self.ws, self.weights_stable)
        elif l.kind == "down":
            ops.conv4s2_fprop(l.x, self.view(self.w16, f"{l.name}/kernel"), bias, l.y, self.ws, self.weights_stable)
        else:
            ops.convT4s2_fprop(l.x, self.view(self.w16, f"{l.name}/kernel"), bias, l.y, self.ws, self.weights_stable)

    def _forward(self, want_pred: bool, backward: bool, inv_n: float) -> None:
        cfg = self.cfg
        for l in self.layers:
            self._fprop(l)
        # the default wiring (runnable here as a cross-check of the tuned engine) hands Dense the image channels too
        image = self.noised if (cfg.block_depth == 0 and cfg.concat and not cfg.residual) else None
        wd, dwd = self.view(self.w, "dense/kernel"), self.view(self.g, "dense/kernel")
        if self._res0:
            # image-level residual: pred = (noised + up0 . Wp) . Wd + bd through the effective kernel [Wp Wd ; Wd]
            image = self.noised
            wd = ops.res0_compose(self.view(self.w, "res0/dense/kernel"), wd, self._weff)
            dwd = self._dweff
            if backward:
                self._dweff.zero_()
        ops.dense_mse(self.dense_in, image, self.x, wd, self.view(self.w, "dense/bias"),
                      self.loss, inv_n, pred=self.pred if want_pred else None,
                      du0=self.gdense_in if backward else None,
                      dwd=dwd if backward else None,
                      dbd=self.view(self.g, "dense/bias") if backward else None, accumulate=True,
                      loss_scale=self.ls if (backward and cfg.mixed_precision) else None,
                      eps=self.eps, t_int=self.t_int, mode=cfg.target_mode, steps=cfg.steps)
        if self._res0 and backward:
            ops.res0_decompose(self._dweff, self.view(self.w, "res0/dense/kernel"), self.view(self.w, "dense/kernel"),
                               self.view(self.g, "res0/dense/kernel"), self.view(self.g, "dense/kernel"))

    def _backward(self, apply_adam: bool, inc_iterations: bool = False) -> None:
        cfg = self.cfg
        for l in reversed(self.layers):
            if l.kind == "proj":
                # d(up_i output) = relu'(.) (g_r . Wp^T); dWp = up_i output^T . g_r; the identity path needs no work: g_r
                # already sits in the buffer down_i's dgrad completes
                ops.conv3s1_wgrad(l.x, l.gy, self._proj_kernel(l, self.g), self.ws_w)
                ops.conv3s1_dgrad(l.gy, self._proj_kernel(l, self.w16), l.gx, l.x, l.mask, False, self.ws,
                                  self.weights_stable)
                continue
            dw = self.view(self.g, f"{l.name}/kernel")
            w16 = self.view(self.w16, f"{l.name}/kernel")
            if l.kind == "image3":
                ops.conv3s1_c3_wgrad(l.x, l.gy, dw, accumulate=True)
            elif l.kind == "down_image":
                ops.conv4s2_c3_wgrad(l.x, l.gy, dw, None, accumulate=True)
            elif l.kind == "s1":
                ops.conv3s1_wgrad(l.x, l.gy, dw, self.ws_w)
                ops.conv3s1_dgrad(l.gy, w16, l.gx, l.x, l.mask, l.add_old, self.ws, self.weights_stable)
            elif l.kind == "down":
                ops.conv4s2_wgrad(l.x, l.gy, dw, self.ws_w)
                ops.conv4s2_dgrad(l.gy, w16, l.gx, l.x, l.add_old, self.ws, self.weights_stable)
            else:
                ops.convT4s2_wgrad(l.x, l.gy, dw, self.ws_w)
                ops.convT4s2_dgrad(l.gy, w16, l.gx, l.x, l.mask, self.ws, self.weights_stable)
        for plan in self._bias_plans:
            ops.bias_grad_multi(plan, accumulate=True)
        if apply_adam:
            ops.adam_apply(self.w, self.m, self.v, self.g, self.w16, self.hyper, cfg.beta1, cfg.beta2, cfg.epsilon, 1.0,
                           iterations_inc=self.iterations if inc_iterations else None)

    def conv_family_pass(self) -> None:
        raise NotImplementedError("bench.py's roofline pass belongs to the default wiring")
