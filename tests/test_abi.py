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
