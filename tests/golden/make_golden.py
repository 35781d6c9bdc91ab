"""Generates tests/golden/tiny_step.npz (and tiny_step_depth1.npz / tiny_step_residual.npz: train.py:20,26 flipped) from the
oracle (run once; re-run only when the oracle is deliberately changed).

The reference itself cannot produce vectors here: it needs TensorFlow (not installed, no network), a GPU at
train.py:40 and the author's dataset at train.py:305,315.  These vectors therefore pin the oracle against drift
and give the CUDA path a committed fixture; they do not pin the oracle to TensorFlow ("parity unpinned").

    python tests/golden/make_golden.py [depth1 residual]      # no argument: all three
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402


VARIANTS = {"": {}, "depth1": dict(block_depth=1), "residual": dict(residual=True)}


def make(name: str, kw: dict):
    """One fixture: tiny_step.npz (train.py's default wiring) or tiny_step_<variant>.npz (a dormant switch flipped)."""
    import dataclasses
    cfg = dataclasses.replace(O.TINY, **kw)
    tr = O.OracleTrainer(cfg, seed=0)
    x, t, e = O.synthetic_batch(cfg, 2, 1)
    loss, grads, taps = O.loss_and_grads(tr.weights, x, t, e, cfg, want_taps=True)
    out = {"loss": np.float32(loss), "pred_corner": taps["pred"][0, :4, :4].numpy()}
    # gradients as TENSORS (a norm cannot see a permuted or transposed gradient): small ones in full, large ones at
    # 2048 fixed random positions (the indices are stored beside the values)
    import torch
    gen = torch.Generator().manual_seed(1234)
    for k, g in grads.items():
        out["gnorm/" + k] = np.float32(g.norm())
        flat = g.reshape(-1)
        if flat.numel() <= 4096:
            idx = torch.arange(flat.numel())
        else:
            idx = torch.randperm(flat.numel(), generator=gen)[:2048].sort().values
        out["gidx/" + k] = idx.numpy().astype(np.int64)
        out["gval/" + k] = flat[idx].numpy()
    out["pred"] = taps["pred"].numpy()
    for k in ("down0", "down3", "up3", "up0", "ddown1", "dup2"):
        v = taps[k]
        if k.startswith(("ddown", "dup")):  # gradient w.r.t. the PRE-activation (what the CUDA path stores): ReLU mask applied
            v = v * (taps[k[1:]] > 0)
        flat = v.reshape(-1)
        idx = torch.randperm(flat.numel(), generator=gen)[:2048].sort().values
        out["aidx/" + k] = idx.numpy().astype(np.int64)
        out["aval/" + k] = flat[idx].numpy()
    for k in ("down0", "down3", "up3", "up0"):
        out["anorm/" + k] = np.float32(taps[k].norm())
    out["losses3"] = np.array([tr.train_step(*O.synthetic_batch(cfg, 2, 100 + s)) for s in range(3)], dtype=np.float32)
    out["w_after3/dense/kernel"] = tr.weights["dense/kernel"].numpy()
    fname = "tiny_step.npz" if not name else f"tiny_step_{name}.npz"
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), fname), **out)
    print(fname, {k: (v.tolist() if v.size < 4 else v.shape) for k, v in out.items() if not k.startswith(("gidx", "gval", "gnorm"))})


def main():
    which = sys.argv[1:] or list(VARIANTS)
    for name in which:
        make(name, VARIANTS[name])


if __name__ == "__main__":
    main()
