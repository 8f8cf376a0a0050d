"""ctypes binding of libddpm_ood_b200.so (the C ABI declared in include/ddpm_ood_b200.h).

The product path has no CPU fallback: if the shared library is missing, loading raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_CSRC = Path(__file__).resolve().parent / "csrc"
LIB_PATH = _CSRC / "libddpm_ood_b200.so"
MAX_LEVELS = 8

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
        ("upsample2", C.c_int),
        ("impl", C.c_int),
        ("stats_out", C.c_void_p),
        ("gn_scale_shift", C.c_void_p),
        ("gn_channels", C.c_int),
        ("gn_st0", C.c_void_p), ("gn_parts0", C.c_int), ("gn_c0", C.c_int),
        ("gn_st1", C.c_void_p), ("gn_parts1", C.c_int), ("gn_c1", C.c_int),
        ("gn_gamma", C.c_void_p), ("gn_beta", C.c_void_p), ("gn_groups", C.c_int), ("gn_eps", C.c_float),
        ("gn_no_act", C.c_int),
        ("concat3x3", C.c_int),
    ]


class UNetConfig(C.Structure):
    _fields_ = [
        ("spatial_dims", C.c_int),
        ("in_channels", C.c_int), ("out_channels", C.c_int),
        ("num_levels", C.c_int),
        ("num_channels", C.c_int * MAX_LEVELS),
        ("attention_levels", C.c_int * MAX_LEVELS),
        ("num_res_blocks", C.c_int * MAX_LEVELS),
        ("num_head_channels", C.c_int * MAX_LEVELS),
        ("norm_num_groups", C.c_int),
        ("norm_eps", C.c_float),
    ]


class VqVaeConfig(C.Structure):
    _fields_ = [
        ("spatial_dims", C.c_int),
        ("in_channels", C.c_int), ("out_channels", C.c_int),
        ("num_levels", C.c_int),
        ("num_res_layers", C.c_int),
        ("num_channels", C.c_int * MAX_LEVELS),
        ("num_res_channels", C.c_int * MAX_LEVELS),
        ("num_embeddings", C.c_int), ("embedding_dim", C.c_int),
        ("precise_encode", C.c_int),
    ]


class PlmsStep(C.Structure):
    _fields_ = [
        ("c", C.c_float * 4),
        ("vA", C.c_float), ("vB", C.c_float),
        ("A", C.c_float), ("Bc", C.c_float),
        ("use_stash", C.c_int),
        ("write_stash", C.c_int),
        ("push", C.c_int),
        ("slot_new", C.c_int),
        ("slot", C.c_int * 3),
    ]


NUM_OP_TYPES = 9
OP_TYPE_NAMES = ("conv_in_small", "conv_in_gemm", "groupnorm_silu", "conv_gemm", "attention_core", "upsample",
                 "conv_out_small_plms", "conv_out_gemm", "time_embed")


class OpProfile(C.Structure):
    _fields_ = [
        ("ms", C.c_double * NUM_OP_TYPES),
        ("flops", C.c_double * NUM_OP_TYPES),
        ("bytes", C.c_double * NUM_OP_TYPES),
        ("launches", C.c_longlong * NUM_OP_TYPES),
        ("forwards", C.c_longlong),
        ("forward_ms", C.c_double),
    ]


class DdpmError(RuntimeError):
    pass


# name -> (restype, argtypes); every symbol include/ddpm_ood_b200.h declares.
SIGNATURES = {
    "ddpm_last_error": (C.c_char_p, []),
    "ddpm_abi_version": (C.c_int, []),
    "ddpm_struct_sizes": (None, [C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ddpm_conv_forward": (C.c_int, [C.POINTER(ConvArgs), C.c_void_p]),
    "ddpm_conv_stats_parts": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "ddpm_conv_halo_stats_parts": (C.c_int, [C.c_int, C.c_int]),
    "ddpm_conv_halo_stats_parts3": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "ddpm_gn_finalize": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]),
    "ddpm_gn_silu": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                               C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p]),
    "ddpm_gn_apply": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
                                C.c_void_p]),
    "ddpm_pack_upconv_weight": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ddpm_attention": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
                                 C.c_void_p]),
    "ddpm_attention_block": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                       C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p]),
    "ddpm_attention_block_stats_parts": (C.c_int, [C.c_int]),
    "ddpm_pack_conv_weight": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_longlong,
                                        C.c_longlong, C.c_void_p]),
    "ddpm_unet_create": (C.c_int, [C.POINTER(UNetConfig), C.POINTER(C.c_void_p)]),
    "ddpm_unet_destroy": (None, [C.c_void_p]),
    "ddpm_unet_set_param": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_longlong, C.c_void_p]),
    "ddpm_unet_finalize": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ddpm_unet_workspace_bytes": (C.c_longlong, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "ddpm_unet_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_void_p, C.c_longlong, C.c_void_p]),
    "ddpm_unet_launch_count": (C.c_longlong, [C.c_void_p]),
    "ddpm_unet_set_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "ddpm_unet_read_profile": (C.c_int, [C.c_void_p, C.POINTER(OpProfile), C.c_int]),
    "ddpm_add_noise": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p,
                                 C.c_int, C.c_longlong, C.c_void_p]),
    "ddpm_plms_update": (C.c_int, [C.c_void_p, C.POINTER(PlmsStep), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_longlong, C.c_void_p]),
    "ddpm_unet_run_chain": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(PlmsStep), C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                      C.c_longlong, C.c_void_p]),
    "ddpm_clamp_mse": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_longlong,
                                 C.c_void_p]),
    "ddpm_val_stats": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ddpm_mean_z": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ddpm_auc_counts": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "ddpm_conv_in_stats_parts": (C.c_int, [C.c_int] * 6),
    "ddpm_conv_in": (C.c_int, [C.c_void_p] * 4 + [C.c_int] * 7 + [C.c_void_p, C.c_void_p]),
    "ddpm_out_norm_conv": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_float, C.c_void_p]),
    "ddpm_simplex_noise": (C.c_int, [C.c_void_p] * 4 + [C.c_int] * 5 + [C.c_double, C.c_double, C.c_void_p]),
    "ddpm_scale_intensity": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_longlong, C.c_void_p]),
    "ddpm_vqvae_create": (C.c_int, [C.POINTER(VqVaeConfig), C.POINTER(C.c_void_p)]),
    "ddpm_vqvae_destroy": (None, [C.c_void_p]),
    "ddpm_vqvae_set_param": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_longlong, C.c_void_p]),
    "ddpm_vqvae_finalize": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ddpm_vqvae_workspace_bytes": (C.c_longlong, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "ddpm_vqvae_encode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_void_p, C.c_longlong, C.c_void_p]),
    "ddpm_vqvae_decode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_void_p, C.c_longlong, C.c_void_p]),
    "ddpm_vqvae_launch_count": (C.c_longlong, [C.c_void_p]),
    "ddpm_lpips_create": (C.c_int, [C.POINTER(C.c_void_p)]),
    "ddpm_lpips_destroy": (None, [C.c_void_p]),
    "ddpm_lpips_set_param": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_longlong, C.c_void_p]),
    "ddpm_lpips_finalize": (C.c_int, [C.c_void_p]),
    "ddpm_lpips_workspace_bytes": (C.c_longlong, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "ddpm_lpips_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                     C.c_int, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p]),
    "ddpm_lpips_launch_count": (C.c_longlong, [C.c_void_p]),
}


ABI_VERSION = 7  # must equal ddpm_abi_version() of the loaded library (include/ddpm_ood_b200.h DDPM_ABI_VERSION)


def _verify(L: C.CDLL) -> None:
    """A stale or foreign .so must not load silently: every declared symbol resolves, the ABI version matches and the
    struct layouts the library was compiled with equal the ctypes mirrors above."""
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(L, name)
        except AttributeError as e:
            raise DdpmError(f"{LIB_PATH} does not export {name}: the library is stale, rebuild it with "
                            f"`python -m ddpm_ood_b200.csrc.build --force`") from e
        fn.restype = res
        fn.argtypes = args
    got = L.ddpm_abi_version()
    if got != ABI_VERSION:
        raise DdpmError(f"{LIB_PATH} has ABI version {got}, this package needs {ABI_VERSION}: rebuild it with "
                        f"`python -m ddpm_ood_b200.csrc.build --force`")
    sizes = [C.c_int(0) for _ in range(4)]
    L.ddpm_struct_sizes(*[C.byref(v) for v in sizes])
    want = [C.sizeof(ConvArgs), C.sizeof(UNetConfig), C.sizeof(PlmsStep), C.sizeof(OpProfile)]
    if [v.value for v in sizes] != want:
        raise DdpmError(f"{LIB_PATH}: struct layouts {[v.value for v in sizes]} differ from the Python mirrors {want} "
                        f"(ConvArgs, UNetConfig, PlmsStep, OpProfile): rebuild the library")


def lib() -> C.CDLL:
    """Load the shared library. If it is missing or older than its sources and nvcc is present it is (re)built first,
    under an exclusive file lock and with an atomic rename, so concurrent ranks of one torchrun launch neither build
    twice nor load a half-written file."""
    global _lib
    if _lib is not None:
        return _lib
    from .csrc import build as _build

    variant = os.environ.get("DDPM_LIB_VARIANT")  # kernel A/B experiments only: an alternately compiled library
    if variant:
        L = C.CDLL(variant)
        _verify(L)
        _lib = L
        return L
    if _build.needs_build() and _build.have_nvcc():
        _build.build(verbose=False)
    if not LIB_PATH.exists():
        raise DdpmError(f"{LIB_PATH} is missing and cannot be built here (no nvcc): build it with "
                        f"`python -m ddpm_ood_b200.csrc.build`")
    L = C.CDLL(str(LIB_PATH))
    _verify(L)
    _lib = L
    return L


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().ddpm_last_error()
        raise DdpmError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def current_stream_ptr() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream
