"""Opcode histogram per kernel of libgct2_b200.so (cuobjdump -sass): the evidence that the conv family is tcgen05 + TMEM +
TMA (UTCHMMA / LDTM / UTMALDG / UTCBAR; no HMMA) and what the CUDA-core kernels are made of.  Runs on a CPU box.

    python tools/sass_summary.py > profiles/sass_summary.txt
"""
import collections
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gan_class_transfer2_b200", "libgct2_b200.so")
CSRC = os.path.join(ROOT, "gan_class_transfer2_b200", "csrc")
KEY = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "UTMALDG", "UTMASTG", "LDTM", "STTM", "SYNCS", "HMMA", "FFMA", "LDG", "STG", "LDS",
       "STS", "RED", "ATOM", "SHFL", "MUFU", "BAR", "STL", "LDL"]


def kernel_sources_sha() -> str:
    """sha256 over the kernel sources: recorded beside every profile so that a stale capture can be recognised."""
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)):
        if name.endswith((".cu", ".cuh")):
            with open(os.path.join(CSRC, name), "rb") as f:
                h.update(name.encode())
                h.update(f.read())
    return h.hexdigest()[:16]


def demangle(names):
    try:
        out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.split("\n")
        return [o.strip() or n for o, n in zip(out, names)]
    except OSError:
        return names


def main():
    exe = "cuobjdump" if subprocess.run(["which", "cuobjdump"], capture_output=True).returncode == 0 else "/usr/local/cuda/bin/cuobjdump"
    sass = subprocess.run([exe, "-sass", LIB], capture_output=True, text=True).stdout
    funcs = collections.OrderedDict()
    cur = None
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            op = m.group(1)
            funcs[cur]["total"] += 1
            base = op.split(".")[0]
            funcs[cur][base] += 1
            if op.startswith("UTCHMMA.2CTA"):
                funcs[cur]["UTCHMMA.2CTA"] += 1
    names = demangle(list(funcs))
    print(f"# libgct2_b200.so, kernel sources sha {kernel_sources_sha()} (tools/sass_summary.py)")
    print("# columns: " + " ".join(["total"] + KEY))
    for (mangled, c), name in zip(funcs.items(), names):
        short = name.replace("(int)", "").replace("(bool)", "").replace("gct2::", "").replace("void ", "")
        short = re.sub(r">\(.*", ">", short) if "<" in short else re.sub(r"\(.*", "", short)
        cells = [f"total={c['total']}"] + [f"{k}={c[k]}" for k in KEY if c[k]]
        print(f"{short:60s} " + " ".join(cells))


if __name__ == "__main__":
    sys.exit(main())
