"""Prints the whole-step parity table (CUDA engine vs the CPU oracle, both flavours) for a configuration: the measured
relative L2 errors behind the tolerances stated in tests/engine_checks.py.

    python tools/parity_table.py --config default --batch 1 [--mixed-precision] [--block-depth D] [--no-concat] [--residual] [--forced]

--forced prints the teacher-forced table instead (tests/engine_checks.py:teacher_forced_parity): the backward pass
against the oracle running on the engine's own activations.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="default", choices=["default", "tiny", "wide"])
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--mixed-precision", action="store_true")
    ap.add_argument("--block-depth", type=int, default=0)
    ap.add_argument("--no-concat", action="store_true")
    ap.add_argument("--residual", action="store_true")
    ap.add_argument("--forced", action="store_true")
    a = ap.parse_args()
    import dataclasses
    from oracle import oracle as O
    from tests import engine_checks as E
    cfg = {"default": O.DEFAULT, "tiny": O.TINY, "wide": O.WIDE}[a.config]
    cfg = dataclasses.replace(cfg, block_depth=a.block_depth, concat=not a.no_concat, residual=a.residual)
    tag = {"config": a.config, "batch": a.batch, "mp": a.mixed_precision, "block_depth": a.block_depth, "concat": cfg.concat,
           "residual": cfg.residual}
    if a.forced:
        res, _ = E.teacher_forced_parity(cfg, a.batch, a.seed)
        for name, err in res.items():
            print(json.dumps({**tag, "forced": True, "quantity": name, "err": round(err, 6), "tol": E.TOL_FORCED}), flush=True)
        return
    res = E.step_parity(cfg, a.batch, a.seed, a.mixed_precision)
    for name, errs in res.items():
        lim = {}
        if name.startswith("grad/") or name.startswith(("act/ddown", "act/dup", "act/dblock")):
            key = name[5:]
            lim = {"emu": E.tol_emu_grad(cfg, key, a.mixed_precision), "f32": E.tol_f32_grad(cfg, key, a.mixed_precision)}
        print(json.dumps({**tag, "quantity": name,
                          **{k: round(v, 5) for k, v in errs.items()},
                          **{"tol_" + k: v for k, v in lim.items()}}), flush=True)


if __name__ == "__main__":
    main()
