"""Whole-step parity: the UNetEngine (CUDA, through the C ABI) against the CPU oracle on the same seeded inputs.

Two comparisons per quantity:
  * "emu": against the oracle with bf16 rounding injected exactly where the CUDA path stores bf16: activations
    <= 6e-3 rel-L2, gradients <= 1.5 % + 1.9 % per U-Net level (tol_emu_grad);
  * "f32": against the oracle in the reference's own fp32 arithmetic (the tolerance north_star asks to be *stated*):
    activations <= 1e-2 rel-L2, loss <= 1e-3 relative, gradients <= 6 % + 2.2 % per level (tol_f32_grad; 10-25 % above the floor rounding alone sets): bf16
    rounding compounds through the ReLU masks of up to 12 layers; SURVEY.md Appendix D measured the same growth.
  Indexing / fusion bugs show up as O(1) errors; the per-kernel checks (tests/kernel_checks.py) are the tight ones.
"""
from __future__ import annotations

from typing import Dict

import torch

from oracle import oracle as O


def rel(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def level_of(layer: str, octaves: int = 0) -> int:
    """U-Net level of 'down3/kernel', 'up2', 'block_up1/conv0/kernel', ... (0 = full resolution; the innermost Block
    sits at `octaves`)."""
    head = layer.split("/")[0]
    if head in ("block_in", "block_out"):
        return 0
    if head == "block_mid":
        return octaves
    return int("".join(ch for ch in head if ch.isdigit()))


def depth_factor(cfg: O.Config) -> float:
    """Free-running comparisons with block_depth > 0: a level holds 1 + block_depth layers on each side, every layer's
    rounding moves the ReLU masks of all later ones, and the flips compound faster than linearly (measured on B200, tiny
    model, activation gradients of level 0 against the same-rounding oracle: 1.2 % / 3.3 % / 6.9 % at block_depth 0 / 1 /
    2).  Stated: the default wiring's tolerances x (1 + block_depth)^1.5.  The tight statement about the backward pass
    of these networks is teacher_forced_parity, which removes the mask flips from the comparison."""
    f = (1.0 + cfg.block_depth) ** 1.5
    # residual = True: the outer levels' gradients come through the projections of the residual sums instead of straight
    # from the concat's skip slice: measured 2.6 % instead of 1.2 % at level 0 (same-rounding oracle), equal further down
    return f * (1.8 if cfg.residual else 1.0)


def no_skip_tol(cfg: O.Config, layer: str, flavour: str):
    """concat = False (train.py:120-121) removes the skip connections: everything an outer layer sees has gone through
    the whole chain, so the forward rounding noise at the last layers is ~10x that of the default wiring (measured on
    B200, tiny model: activations 1.5e-5 -> 4.5e-3 against the same-rounding oracle) and every ReLU mask downstream of it
    flips for a fraction f of its elements -- a relative L2 error of sqrt(f) in the gradients whatever the arithmetic
    does: measured 3.5 - 9.5 % (same-rounding oracle) / 6 - 15.5 % (fp32 oracle), flat over the levels.  Stated: 12 % /
    20 % (x depth factor), Dense 1 %.  An indexing bug is an O(1) error; the tight checks are the per-kernel ones."""
    if layer.startswith("dense") or layer == "pred":
        return 1e-2 * depth_factor(cfg)
    return (0.12 if flavour == "emu" else 0.20) * depth_factor(cfg)


def tol_f32_grad(cfg: O.Config, layer: str, mixed_precision: bool = False) -> float:
    """Stated 16-bit-vs-fp32 tolerance (rel-L2) for the gradients of `layer` (weight gradients and activation
    gradients): bf16 storage 6 % + 2.2 % per U-Net level, fp16 storage (the reference's mixed_float16 policy, three more
    mantissa bits) 2.5 % + 1.2 % per level.  These sit 10-25 % above what rounding to the storage format ALONE does to
    the gradients (tests/test_oracle.py::test_stated_gradient_tolerances_bracket_the_inherent_bf16_error; measured on
    B200, default model, bf16: 4.1 / 7.2 / 9.0 / 10.6 / 13.3 / 14.8 % for the activation gradients of levels 0..5,
    profiles/r2_parity_tables.jsonl) -- the CUDA path adds nothing measurable to that floor.  The Dense layer sits
    directly under the fp32 loss: 0.5 %."""
    if layer.startswith("dense") or layer == "pred":
        return 5e-3
    if mixed_precision:
        return (0.025 + 0.012 * level_of(layer, cfg.octaves)) * depth_factor(cfg)
    return (0.06 + 0.022 * level_of(layer, cfg.octaves)) * depth_factor(cfg)


def tol_emu_grad(cfg: O.Config, layer: str, mixed_precision: bool = False) -> float:
    """Tolerance against the oracle that rounds to the storage format at the same points: what remains is fp32
    summation order, which flips roundings and ReLU masks downstream: bf16 1.5 % + 1.9 % per level (measured 0.2 % ..
    10 % at level 6), fp16 1 % + 1 % per level (measured up to 4.1 %)."""
    if layer.startswith("dense") or layer == "pred":
        return 1e-3
    if mixed_precision:
        return (0.01 + 0.01 * level_of(layer, cfg.octaves)) * depth_factor(cfg)
    return (0.015 + 0.019 * level_of(layer, cfg.octaves)) * depth_factor(cfg)


def make_engine(cfg: O.Config, batch: int, seed: int = 0, use_graph: bool = False, layer_list: bool = False, **net_kw):
    """layer_list=True runs the configuration through block_engine.BlockUNetEngine even when the tuned UNetEngine could
    (the default wiring): an independent schedule over the same kernels."""
    from gan_class_transfer2_b200 import ops
    from gan_class_transfer2_b200 import engine as EN
    from gan_class_transfer2_b200.engine import NetConfig
    net_kw.setdefault("residual", cfg.residual)
    net_kw.setdefault("block_depth", cfg.block_depth)
    net_kw.setdefault("concat", cfg.concat)
    net_kw.setdefault("target_mode", ops.target_mode(cfg.predict_x, cfg.predict_scaled_epsilon, cfg.prediction_weighting,
                                                     cfg.ordinary_differential_equation))
    ncfg = NetConfig(size=cfg.size, pixel_size=cfg.pixel_size, max_size=cfg.max_size, octaves=cfg.octaves,
                     steps=cfg.steps, warm_up=cfg.warm_up, base_lr=cfg.base_lr, beta1=cfg.beta1, beta2=cfg.beta2,
                     epsilon=cfg.epsilon, **net_kw)
    if layer_list:
        from gan_class_transfer2_b200.block_engine import BlockUNetEngine
        eng = BlockUNetEngine(ncfg, batch, use_graph=use_graph)
    else:
        eng = EN.make_engine(ncfg, batch, use_graph=use_graph)
    weights = O.glorot_init(cfg, seed)
    eng.load_weights(weights)
    return eng, weights


def engine_taps(eng) -> Dict[str, torch.Tensor]:
    """The engine's buffers under the oracle's tap names (activations post-ReLU; 'd<name>' = gradient w.r.t. the
    pre-activation, i.e. the oracle's d(loss)/d(output) times the ReLU mask)."""
    n = eng.cfg.octaves
    taps = {"noised": eng.noised, "pred": eng.pred}
    if hasattr(eng, "layers"):  # the layer-list engine: every layer by the oracle's tap name
        for l in eng.layers:
            taps[l.name] = l.y
            if l.kind != "proj":  # (a residual sum's gradient buffer is completed in place into its input's gradient)
                taps["d" + l.name] = l.gy
        return taps
    for i in range(n):
        taps[f"down{i}"] = eng.down_out(i)
        taps[f"up{i}"] = eng.up_out(i)
        taps[f"ddown{i}"] = eng.gdown_out(i)
        taps[f"dup{i}"] = eng.gup_out(i)
    return taps


def _is_dact(name: str) -> bool:
    return name.startswith(("ddown", "dup", "dblock"))


def step_parity(cfg: O.Config, batch: int, seed: int = 0, mixed_precision: bool = False,
                layer_list: bool = False) -> Dict[str, Dict[str, float]]:
    """One forward+backward of the engine vs the oracle (both flavours). Returns {quantity: {"emu": err, "f32": err}}.
    mixed_precision: the reference's fp16 policy with the loss scaled by 2^15 (train.py:34,43-45,82-83); the engine's
    stored activation gradients carry the scale and are compared after dividing it out."""
    eng, weights = make_engine(cfg, batch, seed, mixed_precision=mixed_precision, layer_list=layer_list)
    x, t, e = O.synthetic_batch(cfg, batch, seed + 1)
    loss = eng.loss_and_grads(x.cuda(), t.cuda(), e.cuda())
    torch.cuda.synchronize()
    got_taps = engine_taps(eng)
    scale = float(eng.ls[0]) if mixed_precision else None
    if mixed_precision:
        got_taps = {k: (v.float() / scale if _is_dact(k) else v) for k, v in got_taps.items()}
    got_grads = eng.grads()
    out: Dict[str, Dict[str, float]] = {}
    for flavour, emulate in (("emu", "f16" if mixed_precision else True), ("f32", False)):
        rl, rg, rt = O.loss_and_grads(weights, x, t, e, cfg, want_taps=True, emulate_bf16=emulate,
                                      loss_scale=scale if emulate else None)
        out.setdefault("loss", {})[flavour] = abs(float(loss) - float(rl)) / abs(float(rl))
        for name, ref in rt.items():
            if name == "dpred" or name not in got_taps:
                continue
            if _is_dact(name):
                ref = ref * (rt[name[1:]] > 0)
            out.setdefault("act/" + name, {})[flavour] = rel(got_taps[name], ref)
        for name, ref in rg.items():
            out.setdefault("grad/" + name, {})[flavour] = rel(got_grads[name], ref)
    return out


def check_step_parity(cfg: O.Config, batch: int, seed: int = 0, mixed_precision: bool = False, layer_list: bool = False):
    """Returns (results, failures) with the tolerances stated in this module's docstring."""
    res = step_parity(cfg, batch, seed, mixed_precision, layer_list)
    bad = []
    for name, errs in res.items():
        if not cfg.concat and name != "loss" and (name.startswith("grad/") or _is_dact(name[4:])):
            lim = {f: no_skip_tol(cfg, name[5:], f) for f in ("emu", "f32")}
        elif name == "loss":
            lim = {"emu": 1e-3, "f32": 1e-3}
        elif name.startswith("act/") and _is_dact(name[4:]):
            lim = {"emu": tol_emu_grad(cfg, name[5:], mixed_precision), "f32": tol_f32_grad(cfg, name[5:], mixed_precision)}
        elif name.startswith("act/"):
            lim = {"emu": 6e-3 * depth_factor(cfg), "f32": 1e-2 * depth_factor(cfg)}
        else:
            lim = {"emu": tol_emu_grad(cfg, name[5:], mixed_precision), "f32": tol_f32_grad(cfg, name[5:], mixed_precision)}
        for flavour, err in errs.items():
            if not err <= lim[flavour]:
                bad.append((name, flavour, err, lim[flavour]))
    return res, bad


def loss_curve_parity(cfg: O.Config, batch: int, steps: int, seed: int = 0, use_graph: bool = False, **kw):
    """`steps` training steps (loss, backward, Keras-Adam) on both sides with the same per-step batches and RNG draws;
    returns (engine losses, oracle losses)."""
    eng, weights = make_engine(cfg, batch, seed, use_graph=use_graph, **kw)
    tr = O.OracleTrainer(cfg, weights=weights)
    got, ref = [], []
    for s in range(steps):
        x, t, e = O.synthetic_batch(cfg, batch, 1000 + s)
        got.append(eng.train_step(x.cuda(), t.cuda(), e.cuda()).clone())
        ref.append(tr.train_step(x, t, e))
    torch.cuda.synchronize()
    return [float(g) for g in got], ref, eng, tr


#: teacher_forced_parity: what is left when both sides run backward on the SAME activations is the rounding of the
#: stored 16-bit gradients (2^-9 per layer, random-walking over the layers of the chain) and fp32 summation order
TOL_FORCED = 0.02  # measured on B200: 0.4 - 0.95 % over every configuration of profiles/r2_parity_forced_activations.jsonl


def teacher_forced_parity(cfg: O.Config, batch: int, seed: int = 0, layer_list: bool = False):
    """Backward pass of the engine against the oracle run on the engine's own activations (oracle.denoiser_forward's
    `force`): same inputs and same ReLU masks in every layer, so every activation gradient, weight gradient and bias
    gradient must agree to rounding -- whatever the depth of the network.  A wrong tap, slice, mask or skip-gradient
    add is an O(1) error here, and unlike in the free-running comparison nothing else is.
    Returns ({quantity: rel-L2 error}, failures)."""
    eng, weights = make_engine(cfg, batch, seed, layer_list=layer_list)
    x, t, e = O.synthetic_batch(cfg, batch, seed + 1)
    loss = eng.loss_and_grads(x.cuda(), t.cuda(), e.cuda())
    torch.cuda.synchronize()
    got_taps = engine_taps(eng)
    force = {k: v.detach().float().cpu() for k, v in got_taps.items() if not _is_dact(k) and k not in ("noised", "pred")}
    rl, rg, rt = O.loss_and_grads(weights, x, t, e, cfg, want_taps=True, emulate_bf16=True, force=force)
    res = {"loss": abs(float(loss) - float(rl)) / abs(float(rl))}
    for name, ref in rt.items():
        if _is_dact(name) and name in got_taps:
            res["act/" + name] = rel(got_taps[name], ref * (rt[name[1:]] > 0))
    got_grads = eng.grads()
    for name, ref in rg.items():
        res["grad/" + name] = rel(got_grads[name], ref)
    bad = [(k, v) for k, v in res.items() if not v <= (1e-3 if k == "loss" else TOL_FORCED)]
    return res, bad
