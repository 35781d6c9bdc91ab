"""CPU: the C-ABI library builds/loads and exports exactly what include/gct2_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gct2_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gct2_\w+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from gan_class_transfer2_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.load()


def test_header_declares_the_hot_path_entry_points():
    syms = declared_symbols()
    for must in ["gct2_init", "gct2_noise_images", "gct2_conv4s2_fprop", "gct2_conv4s2_dgrad", "gct2_conv4s2_wgrad",
                 "gct2_convT4s2_fprop", "gct2_convT4s2_dgrad", "gct2_convT4s2_wgrad", "gct2_dense_mse",
                 "gct2_adam_keras", "gct2_bias_grad", "gct2_conv4s2_c3_fprop", "gct2_conv4s2_c3_wgrad"]:
        assert must in syms


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/gct2_b200.h but not exported"


def test_binding_covers_every_declared_symbol():
    from gan_class_transfer2_b200 import _lib
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared_symbols()


def test_abi_version_and_error_string(lib):
    assert lib.gct2_abi_version() == 2
    assert isinstance(lib.gct2_last_error(), bytes)


def test_no_torch_types_in_signatures():
    text = open(HEADER).read()
    assert "torch" not in text.lower().replace("pytorch", "") and "at::" not in text and "std::" not in text
    assert 'extern "C"' in text


def test_init_fails_loudly_without_a_b200(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert lib.gct2_init(0) != 0
    assert len(lib.gct2_last_error()) > 0


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    """The conv family must be tcgen05 + TMA (UTC*MMA / UTMALDG in SASS), not mma.sync (HMMA)."""
    import shutil
    import subprocess
    from gan_class_transfer2_b200 import _lib
    if shutil.which("cuobjdump") is None and not os.path.exists("/usr/local/cuda/bin/cuobjdump"):
        pytest.skip("cuobjdump not available")
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    sass = subprocess.run([exe, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass or "UTCMMA" in sass
    assert "UTMALDG" in sass
    assert "LDTM" in sass
    assert "HMMA." not in sass.replace("UTCHMMA", "")


def test_every_conv_instantiation_is_a_tensor_core_kernel_within_its_register_budget():
    """profiles/sass_summary.txt is generated from the shipped library (tools/sass_summary.py): all 30 instantiations of
    conv_umma_kernel -- the 4x4 / stride-2 maps, their cta_group::2 and cluster split-K variants and the stride-1 maps of
    Block / Residual (modes 3-5) -- issue UTCHMMA and UTMALDG, read their accumulators with LDTM, and contain no HMMA; the
    S / P maps stage split-K partials with LDGSTS."""
    import re
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "sass_summary.py")], capture_output=True, text=True).stdout
    rows = [l for l in out.split("\n") if l.startswith("conv_umma_kernel<")]
    if not rows:
        pytest.skip("cuobjdump produced no SASS")
    inst = {}
    for l in rows:
        name = l.split(">")[0] + ">"
        inst[name] = dict(kv.split("=") for kv in l[len(name):].split() if "=" in kv)
    modes = {int(re.match(r"conv_umma_kernel<(\d)", n).group(1)) for n in inst}
    assert modes == {0, 1, 2, 3, 4, 5} and len(inst) == 30, sorted(inst)
    for name, ops in inst.items():
        assert int(ops.get("UTCHMMA", 0)) > 0 and int(ops.get("UTMALDG", 0)) > 0 and int(ops.get("LDTM", 0)) > 0, name
        assert "HMMA" not in ops, name
    log = os.path.join(root, "gan_class_transfer2_b200", "csrc", "conv_umma.ptxas.log")
    if not os.path.exists(log):
        return  # (a build artefact of csrc/Makefile: absent when the library was built some other way)
    with open(log) as f:
        regs = [int(m) for m in re.findall(r"Used (\d+) registers", f.read())]
    assert regs and max(r for r in regs if r > 60) <= 128  # 384 threads x 128 registers: three quarters of an SM's file
