"""Thin Python wrappers over the C ABI building blocks (torch tensors in, torch tensors out).

torch is used only to own device memory and name the stream; all arithmetic happens in libddpm_ood_b200.so.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from ._lib import ConvArgs, check, current_stream_ptr, lib

EPI_STORE, EPI_SOFTMAX_BD, EPI_STORE_VT = 0, 1, 2


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def pack_conv_weight(w: torch.Tensor, dst: torch.Tensor, koff: int = 0) -> None:
    """w: fp32 [Cout, Cin, *k] on device; dst: fp16 [rows, Ktot] packed matrix (written at column offset koff)."""
    assert w.is_cuda and w.dtype == torch.float32 and w.is_contiguous()
    assert dst.dtype == torch.float16 and dst.is_contiguous()
    cout, cin = w.shape[0], w.shape[1]
    taps = 1
    for k in w.shape[2:]:
        taps *= k
    check(lib().ddpm_pack_conv_weight(w.data_ptr(), cout, cin, taps, dst.data_ptr(), dst.shape[1], koff,
                                      current_stream_ptr()), "ddpm_pack_conv_weight")


def conv_forward(
    segs: Sequence[torch.Tensor],
    ksizes: Sequence[int],
    weights: torch.Tensor,
    cout: int,
    *,
    stride: int = 1,
    bias: Optional[torch.Tensor] = None,
    chan_add: Optional[torch.Tensor] = None,
    residual: Optional[torch.Tensor] = None,
    out: Optional[torch.Tensor] = None,
    mode: int = EPI_STORE,
    b_rows_per_mtile: int = 0,
    scale: float = 1.0,
    group: int = 0,
    vt_col0: int = 0,
    out_vt: Optional[torch.Tensor] = None,
    stats_out: Optional[torch.Tensor] = None,
    upsample2: bool = False,
    impl: int = 0,
    gn_scale_shift: Optional[torch.Tensor] = None,
    concat3x3: bool = False,
    gn_no_act: bool = False,
    gn_stats: Optional[Sequence[torch.Tensor]] = None,
    gn_affine: Optional[Sequence[torch.Tensor]] = None,
    gn_groups: int = 32,
    gn_eps: float = 1e-6,
) -> torch.Tensor:
    """segs: fp16 channels-last tensors [N, (D,) H, W, C]; weights: packed fp16 [rows, Ktot].
    impl 3 = halo-tile kernel; gn_scale_shift (fp32 [N, C3x3, 2], impl 3 only) normalises the 3x3 segments on the fly."""
    x0 = segs[0]
    sd = x0.dim() - 2
    assert sd in (2, 3)
    if sd == 2:
        n, h, w = x0.shape[0], x0.shape[1], x0.shape[2]
        d = 1
    else:
        n, d, h, w = x0.shape[0], x0.shape[1], x0.shape[2], x0.shape[3]
    a = ConvArgs()
    a.spatial_dims = sd
    a.N, a.D, a.H, a.W = n, d, h, w
    a.stride = stride
    a.n_seg = len(segs)
    for i, s in enumerate(segs):
        assert s.dtype == torch.float16 and s.is_contiguous() and s.is_cuda
        a.seg_ptr[i] = s.data_ptr()
        a.seg_channels[i] = s.shape[-1]
        a.seg_ksize[i] = ksizes[i]
    a.weights = weights.data_ptr()
    a.w_rows = weights.shape[0]
    a.Cout = cout
    a.b_rows_per_mtile = b_rows_per_mtile
    a.mode = mode
    a.bias = _ptr(bias)
    a.chan_add = _ptr(chan_add)
    a.residual = _ptr(residual)
    a.upsample2 = int(upsample2)
    a.impl = impl
    so = (lambda v: 2 * v) if upsample2 else (lambda v: (v + stride - 1) // stride)  # noqa: E731
    if out is None:
        shape = (n, so(h), so(w), cout) if sd == 2 else (n, so(d), so(h), so(w), cout)
        out = torch.empty(shape, dtype=torch.float16, device=x0.device)
    a.out = out.data_ptr()
    a.scale = scale
    a.group = group
    a.vt_col0 = vt_col0
    a.out_vt = _ptr(out_vt)
    a.stats_out = _ptr(stats_out)
    a.gn_scale_shift = _ptr(gn_scale_shift)
    a.gn_channels = 0 if gn_scale_shift is None else gn_scale_shift.shape[1]
    a.concat3x3 = int(concat3x3)
    a.gn_no_act = int(gn_no_act)
    if gn_stats is not None:  # statistics [N, parts, C/4, 2] of one or two tensors + (gamma, beta): table built in-kernel
        a.gn_st0, a.gn_parts0, a.gn_c0 = gn_stats[0].data_ptr(), gn_stats[0].shape[1], gn_stats[0].shape[2] * 4
        if len(gn_stats) > 1:
            a.gn_st1, a.gn_parts1, a.gn_c1 = gn_stats[1].data_ptr(), gn_stats[1].shape[1], gn_stats[1].shape[2] * 4
        a.gn_gamma, a.gn_beta = gn_affine[0].data_ptr(), gn_affine[1].data_ptr()
        a.gn_groups, a.gn_eps = gn_groups, gn_eps
    check(lib().ddpm_conv_forward(C.byref(a), current_stream_ptr()), "ddpm_conv_forward")
    return out


def conv_stats_parts(spatial_dims: int, d: int, h: int, w: int) -> int:
    """GroupNorm-statistics parts per image that conv_forward(stats_out=...) emits for an output of this geometry."""
    return int(lib().ddpm_conv_stats_parts(spatial_dims, d, h, w))


def conv_halo_stats_parts(h: int, w: int, d: int = 1) -> int:
    """Parts per image emitted by conv_forward(impl=3, stats_out=...) for a (d x) h x w output."""
    if d > 1:
        return int(lib().ddpm_conv_halo_stats_parts3(d, h, w))
    return int(lib().ddpm_conv_halo_stats_parts(h, w))


def gn_finalize(st0: torch.Tensor, st1: Optional[torch.Tensor], gamma: torch.Tensor, beta: torch.Tensor, spatial: int,
                groups: int, eps: float) -> torch.Tensor:
    """Partial statistics st* [N, parts, C*/4, 2] -> GroupNorm (scale, shift) table fp32 [N, C0+C1, 2]."""
    n = st0.shape[0]
    c0 = st0.shape[2] * 4
    c1 = 0 if st1 is None else st1.shape[2] * 4
    ab = torch.empty((n, c0 + c1, 2), dtype=torch.float32, device=st0.device)
    check(lib().ddpm_gn_finalize(c0, st0.data_ptr(), st0.shape[1], c1, _ptr(st1), 0 if st1 is None else st1.shape[1],
                                 gamma.data_ptr(), beta.data_ptr(), ab.data_ptr(), n, spatial, groups, eps,
                                 current_stream_ptr()), "ddpm_gn_finalize")
    return ab


def gn_silu(src0: torch.Tensor, src1: Optional[torch.Tensor], gamma: torch.Tensor, beta: torch.Tensor, groups: int,
            eps: float, silu: bool = True) -> torch.Tensor:
    """Two-pass GroupNorm(+SiLU) over cat(src0, src1) (channels-last fp16 [N, ..., C])."""
    n = src0.shape[0]
    c0 = src0.shape[-1]
    c1 = 0 if src1 is None else src1.shape[-1]
    s = src0.numel() // (n * c0)
    out = torch.empty(src0.shape[:-1] + (c0 + c1,), dtype=torch.float16, device=src0.device)
    check(lib().ddpm_gn_silu(src0.data_ptr(), c0, _ptr(src1), c1, gamma.data_ptr(), beta.data_ptr(), out.data_ptr(), n, s,
                             groups, eps, int(silu), current_stream_ptr()), "ddpm_gn_silu")
    return out


def gn_apply(src0: torch.Tensor, st0: torch.Tensor, src1: Optional[torch.Tensor], st1: Optional[torch.Tensor],
             gamma: torch.Tensor, beta: torch.Tensor, groups: int, eps: float, silu: bool = True) -> torch.Tensor:
    """One-pass GroupNorm(+SiLU) from producer-side partial statistics st* [N, parts, C/4, 2] fp32."""
    n = src0.shape[0]
    c0 = src0.shape[-1]
    c1 = 0 if src1 is None else src1.shape[-1]
    s = src0.numel() // (n * c0)
    out = torch.empty(src0.shape[:-1] + (c0 + c1,), dtype=torch.float16, device=src0.device)
    check(lib().ddpm_gn_apply(src0.data_ptr(), c0, st0.data_ptr(), st0.shape[1], _ptr(src1), c1, _ptr(st1),
                              0 if st1 is None else st1.shape[1], gamma.data_ptr(), beta.data_ptr(), out.data_ptr(), n, s,
                              groups, eps, int(silu), current_stream_ptr()), "ddpm_gn_apply")
    return out


def conv_in(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, with_stats: bool = True):
    """The UNet's first conv: x fp32 [N, Cin, (D,) H, W], w fp32 [Cout, Cin, 3(,3),3] -> (fp16 [N, (D,) H, W, Cout],
    statistics partials [N, parts, Cout/4, 2] or None)."""
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and w.is_contiguous()
    sd = x.dim() - 2
    n, cin = x.shape[:2]
    d = x.shape[2] if sd == 3 else 1
    h, wd = x.shape[-2:]
    cout = w.shape[0]
    out = torch.empty((n,) + tuple(x.shape[2:]) + (cout,), dtype=torch.float16, device=x.device)
    parts = lib().ddpm_conv_in_stats_parts(cin, cout, sd, d, h, wd) if with_stats else 0
    st = torch.full((n, parts, cout // 4, 2), float("nan"), dtype=torch.float32, device=x.device) if parts else None
    check(lib().ddpm_conv_in(x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), n, cin, d, h, wd, cout, sd,
                             _ptr(st), current_stream_ptr()), "ddpm_conv_in")
    return out, st


def out_norm_conv(src: torch.Tensor, st: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, w: torch.Tensor,
                  b: torch.Tensor, groups: int, eps: float) -> torch.Tensor:
    """The UNet tail GroupNorm -> SiLU -> 3x3 conv to Cout in {1, 3} channels without the normalised tensor in HBM.
    src fp16 [N, H, W, C] with statistics partials st [N, parts, C/4, 2]; w fp32 [Cout, C, 3, 3]; returns fp32
    [N, Cout, H, W]."""
    assert src.dtype == torch.float16 and src.is_contiguous() and src.dim() == 4
    n, h, wd, c = src.shape
    cout = w.shape[0]
    ws = torch.empty((n, h * wd, 9 * cout), dtype=torch.float32, device=src.device)
    out = torch.empty((n, cout, h, wd), dtype=torch.float32, device=src.device)
    check(lib().ddpm_out_norm_conv(src.data_ptr(), st.data_ptr(), st.shape[1], gamma.data_ptr(), beta.data_ptr(),
                                   w.data_ptr(), b.data_ptr(), ws.data_ptr(), out.data_ptr(), n, c, h, wd, cout, groups,
                                   eps, current_stream_ptr()), "ddpm_out_norm_conv")
    return out


def attention(qkv: torch.Tensor, n: int, t: int, heads: int, scale: float, impl: int = 0) -> torch.Tensor:
    """qkv: fp16 [n*t, 3C] (q | k | v); returns fp16 [n*t, C] = softmax(q k^T * scale) v per (image, head)."""
    assert qkv.dtype == torch.float16 and qkv.is_contiguous() and qkv.shape[0] == n * t
    c = qkv.shape[1] // 3
    out = torch.empty((n * t, c), dtype=torch.float16, device=qkv.device)
    check(lib().ddpm_attention(qkv.data_ptr(), out.data_ptr(), n, t, c, heads, scale, impl, current_stream_ptr()),
          "ddpm_attention")
    return out


def attention_block(h: torch.Tensor, n: int, t: int, gamma: torch.Tensor, beta: torch.Tensor, wqkv: torch.Tensor,
                    bqkv: torch.Tensor, wproj: torch.Tensor, bproj: torch.Tensor, eps: float = 1e-6, groups: int = 32,
                    with_stats: bool = False):
    """The whole AttentionBlock in one launch. h: fp16 [n*t, C]; wqkv fp16 [3C, C]; wproj fp16 [C, C]; fp32 norm affine and
    biases. Returns out fp16 [n*t, C] (and the GroupNorm partial statistics of out when with_stats)."""
    assert h.dtype == torch.float16 and h.is_contiguous() and h.shape[0] == n * t
    c = h.shape[1]
    out = torch.empty_like(h)
    stats = None
    if with_stats:
        parts = lib().ddpm_attention_block_stats_parts(t)
        assert parts > 0
        stats = torch.zeros((n, parts, c // 4, 2), dtype=torch.float32, device=h.device)
    check(lib().ddpm_attention_block(h.data_ptr(), out.data_ptr(), n, t, c, 1, groups, eps, 1.0 / c ** 0.5,
                                     gamma.data_ptr(), beta.data_ptr(), wqkv.data_ptr(), bqkv.data_ptr(),
                                     wproj.data_ptr(), bproj.data_ptr(), _ptr(stats), current_stream_ptr()),
          "ddpm_attention_block")
    return (out, stats) if with_stats else out


def pack_upconv_weight(w: torch.Tensor) -> torch.Tensor:
    """w: fp32 [Cout, Cin, 3(,3),3] -> fp16 [2^d * Cout, 2^d * Cin] sub-pixel phase weights for conv_forward(upsample2=True)."""
    assert w.is_cuda and w.dtype == torch.float32 and w.is_contiguous()
    sd = w.dim() - 2
    cout, cin = w.shape[0], w.shape[1]
    dst = torch.empty(((1 << sd) * cout, (1 << sd) * cin), dtype=torch.float16, device=w.device)
    check(lib().ddpm_pack_upconv_weight(w.data_ptr(), cout, cin, sd, dst.data_ptr(), current_stream_ptr()),
          "ddpm_pack_upconv_weight")
    return dst
