"""GPU parity of the tcgen05 implicit-GEMM conv kernel against plain PyTorch fp32 convs of the same fp16-rounded
operands (kernel-level numerics; the end-to-end parity against oracle/ lives in test_unet_gpu.py)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _nhwc(x):  # [N,C,*sp] fp32 -> channels-last fp16 contiguous [N,*sp,C]
    perm = (0,) + tuple(range(2, x.dim())) + (1,)
    return x.permute(*perm).contiguous().half()


def _to_ncx(y):  # channels-last -> [N,C,*sp] fp32
    perm = (0, y.dim() - 1) + tuple(range(1, y.dim() - 1))
    return y.permute(*perm).float()


def _run_case(sd, n, sp, segs, cout, stride=1, use_bias=True, use_cadd=False, use_res=False, seed=0, impl=0, gn=False,
              stats=False):
    """segs: list of (channels, ksize). Returns (rel_l2, max_abs_err/max_ref).
    gn (impl 3): the 3x3 segments are consumed as silu(x * scale[n, c] + shift[n, c]) rounded to fp16, applied on the fly.
    stats (impl 3): also check the epilogue's GroupNorm partial statistics."""
    from ddpm_ood_b200 import ops

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(seed)
    dev = "cuda"
    conv = F.conv2d if sd == 2 else F.conv3d
    xs, ws = [], []
    ktot = sum(c * (k ** sd) for c, k in segs)
    wp = torch.zeros(cout, ktot, dtype=torch.float16, device=dev)
    koff = 0
    ref = None
    abs_ = []
    for c, k in segs:
        x = torch.randn((n, c) + tuple(sp), generator=g, device=dev)
        w = torch.randn((cout, c) + (k,) * sd, generator=g, device=dev) * (1.0 / (c * k ** sd) ** 0.5)
        x16 = _nhwc(x)
        ops.pack_conv_weight(w.contiguous(), wp, koff)
        koff += c * k ** sd
        xs.append(x16)
        xin = _to_ncx(x16)
        if gn and k == 3:
            ab = torch.stack([0.5 + torch.rand((n, c), generator=g, device=dev),
                              torch.randn((n, c), generator=g, device=dev)], dim=-1)
            abs_.append(ab)
            bshape = (n, c) + (1,) * sd
            xin = F.silu(xin * ab[..., 0].view(bshape) + ab[..., 1].view(bshape)).half().float()
        r = conv(xin, w.half().float(), stride=stride, padding=k // 2)
        ref = r if ref is None else ref + r
    bias = torch.randn(cout, generator=g, device=dev) if use_bias else None
    cadd = torch.randn(n, cout, generator=g, device=dev) if use_cadd else None
    if bias is not None:
        ref = ref + bias.view(1, -1, *([1] * sd))
    if cadd is not None:
        ref = ref + cadd.view(n, -1, *([1] * sd))
    res16 = None
    if use_res:
        res = torch.randn(ref.shape, generator=g, device=dev)
        res16 = _nhwc(res)
        ref = ref + _to_ncx(res16)
    st = None
    if stats:
        parts = ops.conv_halo_stats_parts(sp[0], sp[1]) if sd == 2 else ops.conv_halo_stats_parts(sp[1], sp[2], sp[0])
        st = torch.full((n, parts, cout // 4, 2), float("nan"), device=dev)
    out = ops.conv_forward(xs, [k for _, k in segs], wp, cout, stride=stride, bias=bias, chan_add=cadd,
                           residual=res16, impl=impl, stats_out=st,
                           gn_scale_shift=torch.cat(abs_, dim=1).contiguous() if abs_ else None)
    torch.cuda.synchronize()
    if st is not None:
        assert torch.isfinite(st).all()
        o = out.float().reshape(n, -1, cout // 4, 4)
        assert torch.allclose(st.sum(1)[..., 0], o.sum(dim=(1, 3)), rtol=1e-5, atol=1e-2)
        assert torch.allclose(st.sum(1)[..., 1], (o * o).sum(dim=(1, 3)), rtol=1e-5, atol=1e-2)
    got = _to_ncx(out)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    err = (got - ref)
    rel = (err.norm() / ref.norm()).item()
    mx = (err.abs().max() / ref.abs().max()).item()
    return rel, mx


CASES = {
    "3x3_32px_128to128": dict(sd=2, n=4, sp=(32, 32), segs=[(128, 3)], cout=128),
    "3x3_16px_128to256_bias_temb_res": dict(sd=2, n=4, sp=(16, 16), segs=[(128, 3)], cout=256, use_cadd=True,
                                            use_res=True),
    "3x3_8px_256to256_oddN": dict(sd=2, n=3, sp=(8, 8), segs=[(256, 3)], cout=256),
    "3x3_stride2_32to16": dict(sd=2, n=4, sp=(32, 32), segs=[(128, 3)], cout=128, stride=2),
    "3x3_stride2_16to8_256": dict(sd=2, n=5, sp=(16, 16), segs=[(256, 3)], cout=256, stride=2),
    "resnet_conv2_plus_1x1_skip_concat": dict(sd=2, n=4, sp=(16, 16), segs=[(256, 3), (256, 1), (128, 1)], cout=256),
    "linear_qkv": dict(sd=2, n=1, sp=(1, 512), segs=[(256, 1)], cout=768),
    "3x3_28px": dict(sd=2, n=3, sp=(28, 28), segs=[(128, 3)], cout=128),
    "3x3_7px": dict(sd=2, n=5, sp=(7, 7), segs=[(256, 3)], cout=256),
    "3x3_64px_384to128": dict(sd=2, n=2, sp=(64, 64), segs=[(384, 3)], cout=128),
    "3x3x3_8vox_128to128": dict(sd=3, n=2, sp=(8, 8, 8), segs=[(128, 3)], cout=128),
    "3x3x3_stride2_8to4": dict(sd=3, n=3, sp=(8, 8, 8), segs=[(128, 3)], cout=256, stride=2),
    "3x3x3_2vox_256": dict(sd=3, n=20, sp=(2, 2, 2), segs=[(256, 3)], cout=256),
    "many_tiles_persistent": dict(sd=2, n=64, sp=(32, 32), segs=[(128, 3)], cout=128, use_res=True),
    # >= 2 tiles per SM with Cout = 128: CTAs take PAIRS of M tiles sharing each weight tile (conv_gemm_kernel<128, 2>);
    # 43 images x 7 tiles = 301 tiles, so the last pair is half empty
    "paired_m_tiles_odd_count_28px": dict(sd=2, n=43, sp=(28, 28), segs=[(128, 3)], cout=128, use_cadd=True),
    "paired_m_tiles_concat_skip": dict(sd=2, n=40, sp=(32, 32), segs=[(128, 3), (128, 1), (128, 1)], cout=128,
                                       use_res=False),
    # >= 4 tiles per SM with Cout = 128: CTA pairs with two M tiles per CTA (conv_gemm_2cta_kernel<128, 2>), ragged end
    "cta_pair_two_m_tiles_odd": dict(sd=2, n=85, sp=(28, 28), segs=[(128, 3)], cout=128, use_res=True),
    "cta_pair_256_many": dict(sd=2, n=70, sp=(16, 16), segs=[(256, 3), (128, 1)], cout=256, use_cadd=True),
}


@pytest.mark.parametrize("impl", [0, 1], ids=["auto_cta_pair", "single_cta"])
@pytest.mark.parametrize("name", list(CASES))
def test_conv_case(name, impl):
    """impl 0 picks the CTA-pair kernel (tcgen05 cta_group::2) whenever two 128-pixel tiles exist; impl 1 forces the
    single-CTA kernel."""
    rel, mx = _run_case(**CASES[name], impl=impl)
    # fp16 operands are shared with the reference; what remains is fp32 accumulation order + the fp16 output rounding
    # (2^-11 relative).
    assert rel < 6e-4, (name, rel, mx)
    assert mx < 3e-3, (name, rel, mx)


# Halo-tile kernel (impl 3): the haloed input of an 8 x 16 pixel tile is staged once per 64 channels and the 9 taps are
# shifted UMMA-descriptor views of it; optional on-the-fly GroupNorm scale/shift + SiLU of the 3x3 segments' input.
HALO_CASES = {
    "halo_32px_128to128": dict(sd=2, n=4, sp=(32, 32), segs=[(128, 3)], cout=128),
    "halo_16px_128to256_bias_temb_res": dict(sd=2, n=4, sp=(16, 16), segs=[(128, 3)], cout=256, use_cadd=True,
                                             use_res=True),
    "halo_conv2_plus_1x1_skip_concat": dict(sd=2, n=4, sp=(16, 16), segs=[(256, 3), (256, 1), (128, 1)], cout=256),
    "halo_concat_3x3_inputs": dict(sd=2, n=3, sp=(32, 32), segs=[(256, 3), (128, 3)], cout=256, use_cadd=True),
    "halo_28px_ragged_tiles": dict(sd=2, n=3, sp=(28, 28), segs=[(128, 3)], cout=128, use_res=True),
    "halo_64px_384to128": dict(sd=2, n=2, sp=(64, 64), segs=[(384, 3)], cout=128),
    "halo_odd_tile_count_16px_256": dict(sd=2, n=5, sp=(16, 24), segs=[(256, 3)], cout=256),
    "halo_many_tiles_persistent_128": dict(sd=2, n=85, sp=(32, 32), segs=[(128, 3)], cout=128, use_res=True),
    "halo_many_tiles_persistent_256": dict(sd=2, n=170, sp=(16, 16), segs=[(256, 3), (128, 1)], cout=256,
                                           use_cadd=True),
    "halo_512_out_channels": dict(sd=2, n=2, sp=(16, 16), segs=[(128, 3)], cout=512),
    # images of up to 8 x 8 pixels: a tile is TWO whole images with their rows interleaved in shared memory
    "halo_pair_8px_256to256_odd_n": dict(sd=2, n=5, sp=(8, 8), segs=[(256, 3)], cout=256, use_cadd=True, use_res=True),
    "halo_pair_8px_concat_plus_skip": dict(sd=2, n=6, sp=(8, 8), segs=[(256, 3), (256, 1), (256, 1)], cout=256),
    "halo_pair_7px": dict(sd=2, n=4, sp=(7, 7), segs=[(256, 3)], cout=256, use_res=True),
    "halo_pair_4px_512out": dict(sd=2, n=9, sp=(4, 4), segs=[(128, 3), (128, 3)], cout=512),
    "halo_pair_many_items": dict(sd=2, n=333, sp=(8, 8), segs=[(256, 3)], cout=256, use_cadd=True),
    "halo_14px_one_partial_tile_row": dict(sd=2, n=3, sp=(14, 14), segs=[(256, 3)], cout=256),
}


# 3-D volumes of 8 x 8 slabs on the halo kernel: a pair tile is two consecutive depth slabs of one image, a 3x3x3 segment
# runs as three depth-tap stages per 64 channels (slabs outside the volume are TMA zero fill and stay zero under the
# on-the-fly GroupNorm). 128 output channels: conv_halo_kernel<128, 2, true> (four slabs per CTA, D % 4 == 0);
# 256: <256, 1, true> (D % 2 == 0).
HALO_CASES.update({
    "halo3d_8vox_128to128": dict(sd=3, n=2, sp=(8, 8, 8), segs=[(128, 3)], cout=128),
    "halo3d_8vox_odd_n_temb_res": dict(sd=3, n=3, sp=(8, 8, 8), segs=[(128, 3)], cout=128, use_cadd=True, use_res=True),
    "halo3d_two_3x3x3_inputs_384to128": dict(sd=3, n=2, sp=(8, 8, 8), segs=[(256, 3), (128, 3)], cout=128, use_cadd=True),
    "halo3d_conv2_plus_1x1_skip": dict(sd=3, n=3, sp=(8, 8, 8), segs=[(128, 3), (256, 1), (128, 1)], cout=128),
    "halo3d_depth4": dict(sd=3, n=5, sp=(4, 8, 8), segs=[(128, 3)], cout=128, use_res=True),
    "halo3d_256_out_depth6": dict(sd=3, n=3, sp=(6, 8, 8), segs=[(128, 3)], cout=256, use_cadd=True),
    "halo3d_many_items": dict(sd=3, n=85, sp=(8, 8, 8), segs=[(128, 3)], cout=128, use_cadd=True),
    # slabs larger than 8 x 8: 8 x 16 region tiles of one depth slab (the VQ-VAE's residual units, larger latents)
    "halo3d_region_16px_slabs": dict(sd=3, n=2, sp=(4, 16, 16), segs=[(128, 3)], cout=256, use_cadd=True, use_res=True),
    "halo3d_region_ragged_20x12": dict(sd=3, n=3, sp=(3, 20, 12), segs=[(128, 3), (64, 3)], cout=128),
    "halo3d_region_plus_1x1_skip": dict(sd=3, n=2, sp=(5, 16, 16), segs=[(128, 3), (128, 1)], cout=128, use_res=True),
})


@pytest.mark.parametrize("gn", [False, True], ids=["raw", "gn_silu_on_the_fly"])
@pytest.mark.parametrize("name", list(HALO_CASES))
def test_conv_halo_case(name, gn):
    rel, mx = _run_case(**HALO_CASES[name], impl=3, gn=gn, stats=True)
    assert rel < 6e-4, (name, rel, mx)
    assert mx < 3e-3, (name, rel, mx)


# Small-batch tilings of the halo kernel: with few work items the host halves the items' N extent (128-, then 64-wide N
# tiles of one M tile per CTA) so that more clusters share a launch. DDPM_HALO_FINE pins the finest level allowed, so the
# same small problems run on every instantiation (0: <256,1> / <128,2>; 1: <128,1>; 2: <64,1>).
FINE_CASES = ["halo_32px_128to128", "halo_16px_128to256_bias_temb_res", "halo_conv2_plus_1x1_skip_concat",
              "halo_28px_ragged_tiles", "halo_512_out_channels", "halo_pair_8px_256to256_odd_n",
              "halo_pair_8px_concat_plus_skip", "halo_pair_7px"]


@pytest.mark.parametrize("fine", ["0", "1", "2"], ids=["default_tiles", "n128_tiles", "n64_tiles"])
@pytest.mark.parametrize("name", FINE_CASES)
def test_conv_halo_small_batch_tilings(name, fine, monkeypatch):
    monkeypatch.setenv("DDPM_HALO_FINE", fine)
    rel, mx = _run_case(**HALO_CASES[name], impl=3, gn=True, stats=True)
    assert rel < 6e-4, (name, fine, rel, mx)
    assert mx < 3e-3, (name, fine, rel, mx)


@pytest.mark.parametrize("gn", [False, True], ids=["raw", "gn_silu_on_the_fly"])
def test_conv_halo_concat_inputs_one_weight(gn):
    """conv3x3(cat(a, b)) + 1x1 skip conv over a third tensor with ONE torch-layout weight for the 3x3 part (K ordered
    tap, then channel over the concatenation) - ResnetBlock.conv1 of the UNet's up path."""
    from ddpm_ood_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(11)
    n, ca, cb, cs, cout, hw = 5, 256, 128, 128, 256, 16
    xa = _nhwc(torch.randn((n, ca, hw, hw), generator=g, device="cuda"))
    xb = _nhwc(torch.randn((n, cb, hw, hw), generator=g, device="cuda"))
    xs = _nhwc(torch.randn((n, cs, hw, hw), generator=g, device="cuda"))
    w = torch.randn((cout, ca + cb, 3, 3), generator=g, device="cuda") / ((ca + cb) * 9) ** 0.5
    ws = torch.randn((cout, cs, 1, 1), generator=g, device="cuda") / cs ** 0.5
    wp = torch.zeros(cout, (ca + cb) * 9 + cs, dtype=torch.float16, device="cuda")
    ops.pack_conv_weight(w.contiguous(), wp, 0)
    ops.pack_conv_weight(ws.contiguous(), wp, (ca + cb) * 9)
    xin = torch.cat([_to_ncx(xa), _to_ncx(xb)], dim=1)
    ab = None
    if gn:
        ab = torch.stack([0.5 + torch.rand((n, ca + cb), generator=g, device="cuda"),
                          torch.randn((n, ca + cb), generator=g, device="cuda")], dim=-1).contiguous()
        xin = F.silu(xin * ab[..., 0].view(n, -1, 1, 1) + ab[..., 1].view(n, -1, 1, 1)).half().float()
    ref = F.conv2d(xin, w.half().float(), padding=1) + F.conv2d(_to_ncx(xs), ws.half().float())
    out = ops.conv_forward([xa, xb, xs], [3, 3, 1], wp, cout, impl=3, gn_scale_shift=ab, concat3x3=True)
    torch.cuda.synchronize()
    got = _to_ncx(out)
    rel = ((got - ref).norm() / ref.norm()).item()
    assert rel < 6e-4, rel


@pytest.mark.parametrize("hw", [8, 16])
def test_conv_halo_norm_without_activation_on_1x1(hw):
    """AttentionBlock: GroupNorm (no SiLU) applied to the input of the q/k/v projection (a 1x1 conv, N = 768)."""
    from ddpm_ood_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(13)
    n, c, cout = 7, 256, 768
    x16 = _nhwc(torch.randn((n, c, hw, hw), generator=g, device="cuda"))
    w = torch.randn((cout, c, 1, 1), generator=g, device="cuda") / c ** 0.5
    bias = torch.randn(cout, generator=g, device="cuda")
    wp = torch.zeros(cout, c, dtype=torch.float16, device="cuda")
    ops.pack_conv_weight(w.contiguous(), wp, 0)
    ab = torch.stack([0.5 + torch.rand((n, c), generator=g, device="cuda"),
                      torch.randn((n, c), generator=g, device="cuda")], dim=-1).contiguous()
    xin = (_to_ncx(x16) * ab[..., 0].view(n, c, 1, 1) + ab[..., 1].view(n, c, 1, 1)).half().float()
    ref = F.conv2d(xin, w.half().float(), bias=bias)
    out = ops.conv_forward([x16], [1], wp, cout, bias=bias, impl=3, gn_scale_shift=ab, gn_no_act=True)
    torch.cuda.synchronize()
    got = _to_ncx(out)
    rel = ((got - ref).norm() / ref.norm()).item()
    assert rel < 6e-4, rel


@pytest.mark.parametrize("hw", [8, 16, 32])
def test_conv_halo_scale_shift_from_statistics_in_kernel(hw):
    """The kernel's own statistics -> scale/shift step (per work item, in the transform warps) runs the same reduction
    as ddpm_gn_finalize: identical bits to the table path, for one tensor and for a channel concatenation of two."""
    from ddpm_ood_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(17)
    n, ca, cb, cout = 5, 256, 128, 256
    xa = _nhwc(torch.randn((n, ca, hw, hw), generator=g, device="cuda"))
    xb = _nhwc(torch.randn((n, cb, hw, hw), generator=g, device="cuda"))
    w = torch.randn((cout, ca + cb, 3, 3), generator=g, device="cuda") / ((ca + cb) * 9) ** 0.5
    wp = torch.zeros(cout, (ca + cb) * 9, dtype=torch.float16, device="cuda")
    ops.pack_conv_weight(w.contiguous(), wp, 0)
    gamma = torch.randn(ca + cb, generator=g, device="cuda")
    beta = torch.randn(ca + cb, generator=g, device="cuda")

    def stats_of(x, parts):  # what a producer epilogue leaves: per-part partial sums (here: pixel slices)
        q = x.float().reshape(n, parts, -1, x.shape[-1] // 4, 4)
        return torch.stack([q.sum(dim=(2, 4)), (q * q).sum(dim=(2, 4))], dim=-1).contiguous()

    sta, stb = stats_of(xa, 4), stats_of(xb, 8)
    ab = ops.gn_finalize(sta, stb, gamma, beta, hw * hw, 32, 1e-6)
    want = ops.conv_forward([xa, xb], [3, 3], wp, cout, impl=3, gn_scale_shift=ab, concat3x3=True)
    got = ops.conv_forward([xa, xb], [3, 3], wp, cout, impl=3, concat3x3=True, gn_stats=[sta, stb],
                           gn_affine=[gamma, beta], gn_groups=32, gn_eps=1e-6)
    torch.cuda.synchronize()
    assert torch.equal(want, got)
    # and against torch's GroupNorm + SiLU + conv on the same fp16 inputs
    xin = torch.cat([_to_ncx(xa), _to_ncx(xb)], dim=1)
    z = F.silu(F.group_norm(xin, 32, gamma, beta, eps=1e-6)).half().float()
    ref = F.conv2d(z, w.half().float(), padding=1)
    rel = ((_to_ncx(got) - ref).norm() / ref.norm()).item()
    assert rel < 1e-3, rel


def test_conv_halo_matches_im2col_kernel_bitwise():
    """Same operands, same K order inside a 64-channel chunk per tap but a different tap/chunk interleave: fp32
    accumulation order differs, so compare within one fp16 ulp; and the fused GroupNorm path against gn_apply + conv."""
    from ddpm_ood_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(5)
    n, c, cout, hw = 6, 256, 256, 16
    x16 = _nhwc(torch.randn((n, c, hw, hw), generator=g, device="cuda"))
    w = torch.randn((cout, c, 3, 3), generator=g, device="cuda") / (c * 9) ** 0.5
    wp = torch.zeros(cout, c * 9, dtype=torch.float16, device="cuda")
    ops.pack_conv_weight(w.contiguous(), wp, 0)
    gamma = torch.randn(c, generator=g, device="cuda")
    beta = torch.randn(c, generator=g, device="cuda")
    # producer statistics of x16 as a conv epilogue would leave them (one part per image here)
    xq = x16.float().reshape(n, -1, c // 4, 4)
    st = torch.stack([xq.sum(dim=(1, 3)), (xq * xq).sum(dim=(1, 3))], dim=-1).reshape(n, 1, c // 4, 2).contiguous()
    z = ops.gn_apply(x16, st, None, None, gamma, beta, 32, 1e-6, silu=True)
    a = ops.conv_forward([z], [3], wp, cout, impl=0)
    ab = ops.gn_finalize(st, None, gamma, beta, hw * hw, 32, 1e-6)
    b = ops.conv_forward([x16], [3], wp, cout, impl=3, gn_scale_shift=ab)
    torch.cuda.synchronize()
    d = (a.float() - b.float()).abs()
    assert (d <= 2.0 ** -10 * a.float().abs().clamp_min(1.0)).all(), d.max().item()


if __name__ == "__main__":
    for name, kw in list(CASES.items()) + [(k, dict(v, impl=3, gn=True, stats=True)) for k, v in HALO_CASES.items()]:
        try:
            rel, mx = _run_case(**kw)
            print(f"{name:40s} rel_l2={rel:.3e} max={mx:.3e}", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"{name:40s} EXC {type(e).__name__}: {e}", flush=True)


UP_CASES = {
    "up2d_8to16_256": dict(sd=2, n=5, sp=(8, 8), c=256),
    "up2d_16to32_256_paired": dict(sd=2, n=6, sp=(16, 16), c=256),
    "up2d_14to28_128": dict(sd=2, n=3, sp=(14, 14), c=128),
    "up2d_7to14_256": dict(sd=2, n=3, sp=(7, 7), c=256),
    "up2d_16to32_128_many": dict(sd=2, n=40, sp=(16, 16), c=128),
    "up3d_4to8_128": dict(sd=3, n=2, sp=(4, 4, 4), c=128),
}


@pytest.mark.parametrize("impl", [0, 3], ids=["im2col_tiles", "halo_tiles"])
@pytest.mark.parametrize("name", list(UP_CASES))
def test_upsample_conv_as_subpixel_phases(name, impl):
    """conv3x3(nearest_upsample_x2(x)) (the Upsample block of the UNet) computed as 2^d sub-pixel 2x2 convs over the
    low-res input with pre-summed weights, against PyTorch's interpolate + conv on the same fp16 input. The phase
    weights are sums of fp32 weights rounded once to fp16, so the comparison uses the fp32 weights on the torch side."""
    from ddpm_ood_b200 import ops

    case = UP_CASES[name]
    sd, n, sp, c = case["sd"], case["n"], case["sp"], case["c"]
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn((n, c) + tuple(sp), generator=g, device="cuda")
    w = torch.randn((c, c) + (3,) * sd, generator=g, device="cuda") * (1.0 / (c * 3 ** sd) ** 0.5)
    bias = torch.randn(c, generator=g, device="cuda")
    if impl == 3 and sd == 3:
        pytest.skip("the halo-tile kernel is 2-D")
    x16 = _nhwc(x)
    wp = ops.pack_upconv_weight(w.contiguous())
    if impl == 3:
        parts = ops.conv_halo_stats_parts(sp[-2], sp[-1]) * 4
    else:
        parts = ops.conv_stats_parts(sd, 1 if sd == 2 else sp[0], sp[-2], sp[-1]) * (1 << sd)
    st = torch.full((n, parts, c // 4, 2), float("nan"), device="cuda") if parts else None
    out = ops.conv_forward([x16], [2], wp, c, bias=bias, upsample2=True, stats_out=st, impl=impl)
    torch.cuda.synchronize()
    conv = F.conv2d if sd == 2 else F.conv3d
    ref = conv(F.interpolate(_to_ncx(x16), scale_factor=2, mode="nearest"), w, bias=bias, padding=1)
    got = _to_ncx(out)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    rel = ((got - ref).norm() / ref.norm()).item()
    mx = ((got - ref).abs().max() / ref.abs().max()).item()
    assert rel < 8e-4, (name, rel, mx)   # fp16 rounding of the summed weights + fp16 output
    assert mx < 4e-3, (name, rel, mx)
    if st is not None:
        assert torch.isfinite(st).all()
        o = out.float().reshape(n, -1, c // 4, 4)
        assert torch.allclose(st.sum(1)[..., 0], o.sum(dim=(1, 3)), rtol=1e-5, atol=1e-2)
        assert torch.allclose(st.sum(1)[..., 1], (o * o).sum(dim=(1, 3)), rtol=1e-5, atol=1e-2)


@pytest.mark.parametrize("fine", ["0", "1"], ids=["default_tiles", "n128_tiles"])
@pytest.mark.parametrize("name", ["up2d_8to16_256", "up2d_16to32_256_paired", "up2d_7to14_256"])
def test_upsample_conv_halo_small_batch_tilings(name, fine, monkeypatch):
    """The sub-pixel upsample conv on the coarser tilings too (the default, 64-wide N tiles, is the case above)."""
    monkeypatch.setenv("DDPM_HALO_FINE", fine)
    test_upsample_conv_as_subpixel_phases(name, 3)
