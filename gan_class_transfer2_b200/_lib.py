"""ctypes binding of libgct2_b200.so (the C ABI declared in include/gct2_b200.h).

PyTorch is used only for device memory and streams: every entry point receives raw device pointers
(``tensor.data_ptr()``) and the current CUDA stream handle.  There is no CPU fallback: if the shared
library is missing or the device is not sm_100 every call raises.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import c_char_p, c_float, c_int, c_longlong, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GCT2_LIB") or os.path.join(_HERE, "libgct2_b200.so")  # GCT2_LIB: test hook (A/B builds)
CSRC_DIR = os.path.join(_HERE, "csrc")

_lib = None
_inited_devices: set[int] = set()


class Gct2Error(RuntimeError):
    pass


def build(force: bool = False) -> str:
    """Compile the CUDA sources for sm_100a into libgct2_b200.so (in-tree, via csrc/Makefile)."""
    if force:
        subprocess.run(["make", "-C", CSRC_DIR, "clean"], check=True, capture_output=True)
    proc = subprocess.run(["make", "-C", CSRC_DIR, "-j4"], capture_output=True, text=True)
    if proc.returncode != 0:
        raise Gct2Error("building libgct2_b200.so failed:\n" + proc.stdout[-4000:] + proc.stderr[-4000:])
    return LIB_PATH


_P = c_void_p
_PROTOS = {
    "gct2_abi_version": (c_int, []),
    "gct2_last_error": (c_char_p, []),
    "gct2_init": (c_int, [c_int]),
    "gct2_num_sms": (c_int, []),
    "gct2_launch_count": (c_longlong, []),
    "gct2_debug_set": (None, [c_int, c_int]),
    "gct2_debug_timeline": (c_int, [_P, c_int]),
    "gct2_debug_last_plan": (None, [_P]),
    "gct2_set_sm_budget": (None, [c_int]),
    "gct2_set_adam_sms": (None, [c_int]),
    "gct2_set_policy": (None, [c_int]),
    "gct2_get_policy": (c_int, []),
    "gct2_loss_scale_check": (c_int, [_P, c_longlong, _P, _P]),
    "gct2_loss_scale_update": (c_int, [_P, c_int, _P]),
    "gct2_debug_trace": (c_int, [_P, c_int]),
    "gct2_noise_images": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, _P]),
    "gct2_conv4s2_c3_fprop": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "gct2_conv4s2_c3_wgrad": (c_int, [_P, _P, c_int, _P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "gct2_conv4s2_fprop": (c_int, [_P, c_int, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, c_size_t, c_int,
                                   _P]),
    "gct2_conv4s2_dgrad": (c_int, [_P, c_int, _P, _P, c_int, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                   _P, c_size_t, c_int, _P]),
    "gct2_conv4s2_wgrad": (c_int, [_P, c_int, _P, c_int, _P, c_int, c_int, c_int, c_int, c_int, _P, c_size_t, _P]),
    "gct2_convT4s2_fprop": (c_int, [_P, c_int, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, c_size_t, c_int,
                                    _P]),
    "gct2_convT4s2_dgrad": (c_int, [_P, c_int, _P, _P, c_int, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                    _P, c_size_t, c_int, _P]),
    "gct2_convT4s2_wgrad": (c_int, [_P, c_int, _P, c_int, _P, c_int, c_int, c_int, c_int, c_int, _P, c_size_t, _P]),
    "gct2_conv3s1_fprop": (c_int, [_P, c_int, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P, c_size_t,
                                   c_int, _P]),
    "gct2_conv3s1_fprop_add": (c_int, [_P, c_int, _P, _P, c_int, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                       _P]),
    "gct2_conv3s1_dgrad": (c_int, [_P, c_int, _P, _P, c_int, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                   c_int, _P, c_size_t, c_int, _P]),
    "gct2_conv3s1_wgrad": (c_int, [_P, c_int, _P, c_int, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, c_size_t, _P]),
    "gct2_conv3s1_c3_fprop": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "gct2_conv3s1_c3_wgrad": (c_int, [_P, _P, c_int, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "gct2_bias_grad": (c_int, [_P, c_int, c_longlong, c_int, _P, _P]),
    "gct2_bias_grad_multi": (c_int, [c_int, _P, _P, _P, _P, _P, c_int, _P]),
    "gct2_dense_mse": (c_int, [_P, c_int, _P, _P, _P, _P, _P, _P, _P, c_int, _P, _P, c_longlong, c_int, c_float,
                               c_int, c_int, _P, _P, _P, c_longlong, c_int, c_int, _P]),
    "gct2_res0_compose": (c_int, [_P, _P, _P, c_int, _P]),
    "gct2_res0_decompose": (c_int, [_P, _P, _P, _P, _P, c_int, _P]),
    "gct2_adam_keras": (c_int, [_P, _P, _P, _P, _P, c_longlong, _P, _P, c_float, c_int, c_float, c_float, c_float,
                                c_float, _P]),
    "gct2_adam_prepare": (c_int, [_P, _P, c_float, c_int, c_float, c_float, _P]),
    "gct2_adam_apply": (c_int, [_P, _P, _P, _P, _P, c_longlong, _P, c_float, c_float, c_float, c_float, _P, _P, _P]),
    "gct2_adam_apply_g16": (c_int, [_P, _P, _P, _P, _P, c_longlong, _P, c_float, c_float, c_float, c_float, _P, _P]),
    "gct2_adam_apply_p2p": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int, c_longlong, c_longlong, _P, c_float, c_float, c_float,
                                    c_float, c_int, _P]),
    "gct2_sum_peers_f32": (c_int, [_P, c_int, _P, c_longlong, _P, c_longlong, _P]),
    "gct2_step_begin": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, ctypes.c_ulonglong, _P, _P, c_float, c_int, c_float,
                                c_float, _P, c_longlong, _P, _P]),
    "gct2_step_begin_u8": (c_int, [_P, _P, _P, c_int, _P, _P, _P, c_int, c_int, c_int, ctypes.c_ulonglong, _P, _P, c_float,
                                   c_int, c_float, c_float, _P, c_longlong, _P, _P]),
    "gct2_sample_update": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_longlong, c_int, _P]),
    "gct2_latent_edits": (c_int, [_P, _P, _P, c_int, c_int, _P]),
    "gct2_rmse": (c_int, [_P, _P, c_longlong, _P, _P]),
    "gct2_cast_bf16": (c_int, [_P, _P, c_longlong, _P]),
}
EXPORTED_SYMBOLS = tuple(_PROTOS)


def load() -> ctypes.CDLL:
    """dlopen the library and attach prototypes. Raises Gct2Error when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Gct2Error(
            f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the training step)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.gct2_abi_version() != 2:
        raise Gct2Error("libgct2_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def init(device: int = 0) -> ctypes.CDLL:
    lib = load()
    if device not in _inited_devices:
        if lib.gct2_init(device) != 0:
            raise Gct2Error(lib.gct2_last_error().decode())
        _inited_devices.add(device)
        # test hook: GCT2_DEBUG="key=value,key=value" -> gct2_debug_set (see include/gct2_b200.h)
        for item in filter(None, os.environ.get("GCT2_DEBUG", "").split(",")):
            key, value = item.split("=")
            lib.gct2_debug_set(int(key), int(value))
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise Gct2Error(load().gct2_last_error().decode())


def ptr(t) -> int:
    """Device pointer of a torch tensor (or 0 for None)."""
    return 0 if t is None else t.data_ptr()


def current_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
