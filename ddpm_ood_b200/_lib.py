"""ctypes binding of libddpm_ood_b200.so (the C ABI declared in include/ddpm_ood_b200.h).

The product path has no CPU fallback: if the shared library is missing, loading raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_CSRC = Path(__file__).resolve().parent / "csrc"
LIB_PATH = _CSRC / "libddpm_ood_b200.so"

_lib = None


class ConvArgs(C.Structure):
    _fields_ = [
        ("spatial_dims", C.c_int),
        ("N", C.c_int), ("D", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("stride", C.c_int),
        ("n_seg", C.c_int),
        ("seg_ptr", C.c_void_p * 3),
        ("seg_channels", C.c_int * 3),
        ("seg_ksize", C.c_int * 3),
        ("weights", C.c_void_p),
        ("w_rows", C.c_int),
        ("Cout", C.c_int),
        ("b_rows_per_mtile", C.c_int),
        ("mode", C.c_int),
        ("bias", C.c_void_p),
        ("chan_add", C.c_void_p),
        ("residual", C.c_void_p),
        ("out", C.c_void_p),
        ("scale", C.c_float),
        ("group", C.c_int),
        ("vt_col0", C.c_int),
        ("out_vt", C.c_void_p),
    ]


class DdpmError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load the shared library (building it first if the sources are newer and nvcc is present)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        from .csrc.build import build

        build(verbose=False)
    if not LIB_PATH.exists():
        raise DdpmError(f"{LIB_PATH} is missing: build it with `python -m ddpm_ood_b200.csrc.build`")
    L = C.CDLL(str(LIB_PATH))
    L.ddpm_last_error.restype = C.c_char_p
    L.ddpm_abi_version.restype = C.c_int
    L.ddpm_conv_forward.argtypes = [C.POINTER(ConvArgs), C.c_void_p]
    L.ddpm_conv_forward.restype = C.c_int
    L.ddpm_pack_conv_weight.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_longlong,
                                        C.c_longlong, C.c_void_p]
    L.ddpm_pack_conv_weight.restype = C.c_int
    _lib = L
    return L


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().ddpm_last_error()
        raise DdpmError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def current_stream_ptr() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream
