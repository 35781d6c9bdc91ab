"""Runs every per-kernel parity check and prints one JSON line per case (GPU box only).

    python tools/diag_kernels.py              # all groups, each in its own subprocess (a trap poisons a context)
    python tools/diag_kernels.py --group NAME # one group in this process
"""
import argparse
import json
import os
import subprocess
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def groups():
    from tests import kernel_checks as K
    out = {}
    for fn, kw in K.CONV_CASES + K.EW_CASES:
        out.setdefault(fn.__name__, []).append((fn, kw))
    for fn, kw, force in K.FORCED_CASES:
        out.setdefault("forced_" + fn.__name__, []).append((lambda fn=fn, force=force, **k: K.forced(fn, **force, **k), kw))
    return out


def run_group(name, debug):
    from tests import kernel_checks as K
    from gan_class_transfer2_b200 import _lib
    lib = _lib.init(0)
    for key, val in debug:
        lib.gct2_debug_set(key, val)
    ok = True
    for fn, kw in groups()[name]:
        try:
            m = fn(**kw)
            m["pass"] = K.passed(m)
        except Exception as e:  # noqa: BLE001
            m = {"name": f"{fn.__name__} {kw}", "pass": False, "exception": repr(e)[:300]}
            traceback.print_exc()
        ok &= m["pass"]
        print(json.dumps(m), flush=True)
    return ok


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--group")
    ap.add_argument("--debug", action="append", default=[], help="key=value for gct2_debug_set")
    ap.add_argument("--only", default="", help="comma-separated group-name substrings")
    a = ap.parse_args()
    debug = [tuple(int(v) for v in d.split("=")) for d in a.debug]
    if a.group:
        sys.exit(0 if run_group(a.group, debug) else 1)
    rc = 0
    for name in groups():
        if a.only and not any(s in name for s in a.only.split(",")):
            continue
        cmd = ["timeout", "300", sys.executable, os.path.abspath(__file__), "--group", name]
        for d in a.debug:
            cmd += ["--debug", d]
        print(f"=== {name}", flush=True)
        r = subprocess.run(cmd)
        if r.returncode != 0:
            rc = 1
            print(f"=== {name} exit {r.returncode}", flush=True)
    sys.exit(rc)


if __name__ == "__main__":
    main()
