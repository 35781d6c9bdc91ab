"""CPU: the layer list block_engine.build_layer_list produces for every combination of train.py's wiring switches
(block_depth, residual, concat: train.py:20,26-27), checked on SYMBOLIC buffers -- no GPU, no kernels:

  * shapes: every layer reads / writes the channel counts its variables have (engine.variable_specs), at the right extent;
  * forward data flow: a layer only reads what an earlier layer (or the image) has written, completely;
  * backward data flow, simulated in the engine's order (reverse list: weight gradient, then data gradient): every
    gradient a layer consumes is COMPLETE when it is read -- all its consumers have contributed and the ReLU mask of its
    producer has been applied exactly once, or never for a residual sum, which is not a ReLU output -- and no data
    gradient overwrites a part that another consumer has already stored (only `add_old` may touch it again).

The GPU tests check the numbers; this one checks the bookkeeping that decides which numbers end up where."""
import dataclasses
import itertools

import pytest
import torch

from gan_class_transfer2_b200 import engine as E
from gan_class_transfer2_b200.block_engine import build_layer_list


class Buf:
    """A [B,H,H,C] buffer or a channel slice of one; identity = (root id, channel range)."""
    _ids = itertools.count()

    def __init__(self, C, H, dtype, root=None, lo=0):
        self.C, self.H, self.dtype = C, H, dtype
        self.root = next(Buf._ids) if root is None else root
        self.lo = lo
        self.shape = (1, H, H, C)

    def __getitem__(self, key):
        assert key[0] is Ellipsis and isinstance(key[1], slice) and key[1].step is None
        lo, hi, _ = key[1].indices(self.C)
        return Buf(hi - lo, self.H, self.dtype, self.root, self.lo + lo)

    def channels(self):
        return {(self.root, c) for c in range(self.lo, self.lo + self.C)}


def plan(cfg):
    pairs = []

    def pair(C, H):
        a, g = Buf(C, H, torch.bfloat16), Buf(C, H, torch.bfloat16)
        pairs.append((a, g))
        return a, g

    image = Buf(3, cfg.size, torch.float32)
    layers, cat, gcat, dense_in, gdense_in = build_layer_list(cfg, image, pair, lambda t: Buf(t.C, t.H, t.dtype))
    return image, layers, dense_in, gdense_in


CONFIGS = [dict(block_depth=d, concat=c, residual=r) for d in (0, 1, 2) for c in (True, False) for r in (False, True)]


@pytest.mark.parametrize("kw", CONFIGS, ids=lambda kw: "-".join(f"{k}{int(v)}" for k, v in kw.items()))
@pytest.mark.parametrize("octaves", [2, 4])
def test_layer_list_shapes_and_data_flow(kw, octaves):
    cfg = E.NetConfig(size=64, pixel_size=128, max_size=256, octaves=octaves, **kw)
    cfg.validate()
    specs = dict(E.variable_specs(cfg))
    image, layers, dense_in, gdense_in = plan(cfg)
    # every variable of the network belongs to exactly one layer (+ Dense(3), + the image-level projection that is folded
    # into Dense(3) when residual = True and block_depth = 0)
    kernels = {n[:-len("/kernel")] for n in specs if n.endswith("/kernel")}
    folded = {"res0/dense"} if (cfg.residual and cfg.block_depth == 0) else set()
    assert {l.name for l in layers} | {"dense"} | folded == kernels
    assert len({l.name for l in layers}) == len(layers)

    # ---- shapes + forward data flow
    written = set(image.channels())
    relu_output = set()            # channels holding a ReLU output (what a consumer's dgrad may mask by)
    for l in layers:
        k = specs[f"{l.name}/kernel"]
        cin, cout = (k[3], k[2]) if l.kind == "up" else (k[-2], k[-1])
        assert (l.x.C, l.y.C) == (cin, cout), (l.name, l.x.C, l.y.C, k)
        scale = {"down": 0.5, "down_image": 0.5, "up": 2}.get(l.kind, 1)
        assert l.y.H == l.x.H * scale and l.gy.H == l.y.H and l.gy.C == l.y.C, l.name
        assert (l.gx is None) == (l.kind in ("image3", "down_image")), l.name
        assert l.x.channels() <= written, f"{l.name} reads channels nobody has written"
        assert not (l.y.channels() & written), f"{l.name} overwrites an activation that is still needed"
        written |= l.y.channels()
        if l.kind == "proj":
            assert l.res is not None and l.res.channels() <= written and l.res.C == l.y.C
        else:
            relu_output |= l.y.channels()
        # the ReLU mask a dgrad applies must cover exactly the leading channels of x that ARE ReLU outputs of one producer
        if l.gx is not None and not l.add_old:
            masked = l.x[..., :l.mask].channels()
            assert masked <= relu_output, l.name
    # Dense(3) reads the 16-bit channels (+ the 3 image channels separately in the default wiring); with the image-level
    # residual folded into it, its 16-bit input is the projection's input
    want = specs["res0/dense/kernel"][0] if folded else specs["dense/kernel"][0] - (3 if cfg.fused_default else 0)
    assert dense_in.channels() <= written and dense_in.C == want

    # ---- backward data flow.  state per gradient channel: number of contributions stored, and whether it is masked
    grad_of = {}                   # activation channel -> gradient channel (through the layers' (y, gy) pairs)
    for l in layers:
        for a, g in zip(sorted(l.y.channels()), sorted(l.gy.channels())):
            grad_of[a] = g
    consumers = {}                 # activation channel -> how many layers (or Dense / a residual identity) read it
    for l in layers:
        for a in l.x.channels():
            consumers[a] = consumers.get(a, 0) + 1
        if l.kind == "proj":
            for a in l.res.channels():
                consumers[a] = consumers.get(a, 0) + 1   # the identity path of the residual sum
    for a in dense_in.channels():
        consumers[a] = consumers.get(a, 0) + 1
    contributions, masked = {}, set()
    for g in gdense_in.channels():  # Dense(3)+MSE writes the gradient of its input, masked by it
        contributions[g] = 1
        masked.add(g)
    for l in reversed(layers):
        # what this layer's weight / bias / data gradient read: complete and masked exactly when y is a ReLU output
        for a, g in zip(sorted(l.y.channels()), sorted(l.gy.channels())):
            assert contributions.get(g, 0) == consumers.get(a, 0) >= 1, f"{l.name}: gradient of its output is incomplete"
            assert (g in masked) == (l.kind != "proj"), f"{l.name}: ReLU mask of its output gradient"
        if l.kind == "proj":
            # the identity path: the gradient of the sum IS a contribution to the gradient of the residual input (the
            # buffers are the same, so nothing is launched for it)
            assert sorted(l.gy.channels()) == sorted(grad_of[a] for a in sorted(l.res.channels())), l.name
        if l.gx is None:
            continue
        for idx, (a, g) in enumerate(zip(sorted(l.x.channels()), sorted(l.gx.channels()))):
            assert grad_of[a] == g, f"{l.name}: its data gradient goes to the wrong buffer"
            if l.add_old:
                assert contributions.get(g, 0) >= 1 and g not in masked, f"{l.name}: add_old on an empty or finished gradient"
            else:
                assert contributions.get(g, 0) == 0, f"{l.name}: overwrites a gradient another consumer has stored"
            contributions[g] = contributions.get(g, 0) + 1
            if idx < l.mask:
                assert contributions[g] == consumers[a], f"{l.name}: masks a gradient before its last contribution"
                masked.add(g)


def test_default_wiring_matches_the_tuned_engine_layout():
    """block_depth = 0, concat = True through the layer list: the launch order of UNetEngine (down0..down{n-1},
    up{n-1}..up0) and the same slice layout (up_j output first, skip second)."""
    cfg = E.NetConfig(size=64, pixel_size=128, max_size=256, octaves=4)
    _, layers, dense_in, _ = plan(cfg)
    assert [l.name for l in layers] == [f"down{i}" for i in range(4)] + [f"up{i}" for i in reversed(range(4))]
    up1 = next(l for l in layers if l.name == "up1")
    down0 = next(l for l in layers if l.name == "down0")
    assert up1.y.root == down0.y.root and up1.y.lo == 0 and down0.y.lo == cfg.up_c(1)   # cat[1] = [up1 | down0]
    assert dense_in.C == cfg.up_c(0)
    down1 = next(l for l in layers if l.name == "down1")
    assert down1.add_old and down1.mask == cfg.down_c(0)
    up0 = next(l for l in layers if l.name == "up0")
    assert not up0.add_old and up0.mask == cfg.up_c(1) and up0.x.C == cfg.up_c(1) + cfg.down_c(0)
