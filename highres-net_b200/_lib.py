"""ctypes binding of libhrn_b200.so (C ABI in include/hrn_b200.h).

There is no CPU or PyTorch fallback: if the shared library is missing and cannot
be built with nvcc, importing the package fails loudly."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int32, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# HRN_B200_LIB lets tools/ab.py load another build of the same ABI for same-box A/B timing
_LIB_PATH = os.environ.get("HRN_B200_LIB") or os.path.join(_HERE, "csrc", "libhrn_b200.so")


class HrnConfig(ctypes.Structure):
    _fields_ = [(n, c_int32) for n in (
        "enc_in_channels", "enc_num_layers", "enc_kernel_size", "enc_channels",
        "rec_alpha_residual", "rec_in_channels", "rec_kernel_size",
        "dec_in_channels", "dec_kernel_size", "dec_stride", "dec_out_channels",
        "fin_in_channels", "fin_kernel_size", "fin_out_channels")]


# every symbol include/hrn_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "hrn_abi_version": (c_int32, []),
    "hrn_last_error": (c_char_p, []),
    "hrn_create": (c_int32, [POINTER(HrnConfig), c_int32, POINTER(c_void_p)]),
    "hrn_destroy": (None, [c_void_p]),
    "hrn_set_weight": (c_int32, [c_void_p, c_char_p, c_void_p, POINTER(c_int64), c_int32]),
    "hrn_missing_weights": (c_int32, [c_void_p]),
    "hrn_forward": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "hrn_reserve": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_int32]),
    "hrn_forward_host": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p,
                                   c_void_p]),
    "hrn_forward_host_u16": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p,
                                       c_void_p]),
    "hrn_forward_host_submit": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p,
                                          c_void_p, POINTER(c_int64)]),
    "hrn_forward_host_wait": (c_int32, [c_void_p, c_int64]),
    "hrn_u16_to_unit_float": (c_int32, [c_void_p, c_int64, c_void_p, c_void_p]),
    "hrn_unit_float_to_u16": (c_int32, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "hrn_collate": (c_int32, [c_void_p, c_int32, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "hrn_png_info": (c_int32, [c_char_p, POINTER(c_int32), POINTER(c_int32), POINTER(c_int32), POINTER(c_int32)]),
    "hrn_png_read_gray_u16": (c_int32, [POINTER(c_char_p), c_int32, c_int32, c_int32, c_void_p, c_int32]),
    "hrn_png_write_gray_u16": (c_int32, [POINTER(c_char_p), c_int32, c_int32, c_int32, c_void_p, c_int32]),
    "hrn_clearance_scores": (c_int32, [POINTER(c_char_p), c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "hrn_clearance_order": (c_int32, [c_void_p, c_int32, c_void_p]),
    "hrn_zip_store": (c_int32, [c_char_p, POINTER(c_char_p), POINTER(c_char_p), c_int32]),
    "hrn_lanczos_shift": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                                    c_int32, c_void_p, c_void_p]),
    "hrn_lanczos_taps": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "hrn_shift_cpsnr": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                                  c_void_p, c_void_p, c_void_p, c_void_p]),
    "hrn_clear_loss": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "hrn_forward_dump": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p,
                                   c_int32, c_void_p, c_void_p]),
    "hrn_profile_begin": (c_int32, [c_void_p]),
    "hrn_profile_end": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "hrn_debug_set": (c_int32, [c_void_p, c_char_p, c_int32]),
    "hrn_scoring_debug_set": (c_int32, [c_char_p, c_int32]),
    "hrn_kernel_launch_count": (c_int64, []),
    "hrn_shiftnet_create": (c_int32, [c_int32, POINTER(c_void_p)]),
    "hrn_shiftnet_destroy": (None, [c_void_p]),
    "hrn_shiftnet_set_weight": (c_int32, [c_void_p, c_char_p, c_void_p, POINTER(c_int64), c_int32]),
    "hrn_shiftnet_missing_weights": (c_int32, [c_void_p]),
    "hrn_shiftnet_debug_set": (c_int32, [c_void_p, c_char_p, c_int32]),
    "hrn_shiftnet_forward": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
}

_lib = None


def library_path() -> str:
    return _LIB_PATH


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        # in-tree build (nvcc cross-compiles sm_100a); raises if nvcc is unavailable
        import importlib.util
        spec = importlib.util.spec_from_file_location("_hrn_build", os.path.join(_HERE, "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    try:
        lib = ctypes.CDLL(_LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise ImportError(f"cannot load {_LIB_PATH}: {e}; the CUDA extension is mandatory (no CPU fallback)") from e
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.hrn_abi_version() != 1:
        raise ImportError("libhrn_b200.so ABI version mismatch; rebuild with highres-net_b200/build.py --force")
    _lib = lib
    return lib


def last_error() -> str:
    return load().hrn_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed: {last_error()}")


def kernel_launch_count() -> int:
    return int(load().hrn_kernel_launch_count())


def require_cuda_tensor(t, name: str):
    import torch
    if not torch.is_tensor(t) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the B200 path has no CPU fallback")
    return t


def current_stream_ptr(device) -> int:
    import torch
    return int(torch.cuda.current_stream(device).cuda_stream)
