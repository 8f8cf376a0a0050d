"""GPU parity of the GroupNorm kernels against torch.nn.functional.group_norm (fp32) on the same fp16 inputs:
the two-pass kernel (ddpm_gn_silu), the statistics the conv epilogue emits (ddpm_conv_args.stats_out) and the one-pass
kernel that consumes them (ddpm_gn_apply). GroupNorm(32, eps=1e-6)+SiLU sits in front of every conv of the reference's
DiffusionModelUNet (built at src/trainers/base.py:66-75)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ref_gn(x16_list, gamma, beta, groups, eps, silu):
    x = torch.cat([t.float() for t in x16_list], dim=-1)  # [N, ..., C]
    perm = (0, x.dim() - 1) + tuple(range(1, x.dim() - 1))
    y = F.group_norm(x.permute(*perm), groups, gamma, beta, eps)
    if silu:
        y = F.silu(y)
    inv = (0,) + tuple(range(2, x.dim())) + (1,)
    return y.permute(*inv)


def _stats_from(out16, parts):
    n, c = out16.shape[0], out16.shape[-1]
    x = out16.float().reshape(n, -1, c // 4, 4)
    return x.sum(dim=(1, 3)), (x * x).sum(dim=(1, 3))


CASES = [
    dict(n=3, sp=(32, 32), c0=128, c1=0),
    dict(n=2, sp=(16, 16), c0=256, c1=128),   # concatenation whose groups (12 channels) straddle the two tensors
    dict(n=5, sp=(8, 8), c0=256, c1=256),     # two images per conv tile
    dict(n=2, sp=(28, 28), c0=128, c1=0),     # ragged tile boxes (28 -> 32)
    dict(n=3, sp=(7, 7), c0=256, c1=0),
    dict(n=2, sp=(64, 64), c0=256, c1=128),
    dict(n=40, sp=(32, 32), c0=128, c1=0),    # enough 128-wide tiles for the paired-M-tile conv kernel variant
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"n{c['n']}_{c['sp'][0]}px_{c['c0']}+{c['c1']}")
def test_groupnorm_kernels_match_torch(case):
    from ddpm_ood_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(0)
    n, sp, c0, c1 = case["n"], case["sp"], case["c0"], case["c1"]
    dev = "cuda"
    outs, stats = [], []
    for c in (c0, c1):
        if not c:
            continue
        # produce the tensor with the conv kernel (1x1 conv) so that its epilogue emits the statistics
        x = (torch.randn((n,) + sp + (128,), generator=g, device=dev) * 1.5 + 0.3).half()
        w = torch.randn((c, 128, 1, 1), generator=g, device=dev) * 0.1
        wp = torch.zeros(c, 128, dtype=torch.float16, device=dev)
        ops.pack_conv_weight(w.contiguous(), wp)
        bias = torch.randn(c, generator=g, device=dev)
        parts = ops.conv_stats_parts(2, 1, sp[0], sp[1])
        assert parts > 0
        st = torch.full((n, parts, c // 4, 2), float("nan"), dtype=torch.float32, device=dev)
        out = ops.conv_forward([x], [1], wp, c, bias=bias, stats_out=st)
        torch.cuda.synchronize()
        assert torch.isfinite(st).all(), "every statistics part must be written"
        want_s, want_q = _stats_from(out, parts)
        got = st.sum(dim=1)
        assert torch.allclose(got[..., 0], want_s, rtol=1e-5, atol=1e-2), (got[..., 0] - want_s).abs().max()
        assert torch.allclose(got[..., 1], want_q, rtol=1e-5, atol=1e-2), (got[..., 1] - want_q).abs().max()
        outs.append(out)
        stats.append(st)
    C = c0 + c1
    gamma = 1 + 0.1 * torch.randn(C, generator=g, device=dev)
    beta = 0.1 * torch.randn(C, generator=g, device=dev)
    for silu in (True, False):
        want = _ref_gn(outs, gamma, beta, 32, 1e-6, silu)
        two = ops.gn_silu(outs[0], outs[1] if c1 else None, gamma, beta, 32, 1e-6, silu).float()
        one = ops.gn_apply(outs[0], stats[0], outs[1] if c1 else None, stats[1] if c1 else None, gamma, beta, 32, 1e-6,
                           silu).float()
        # fp32 statistics, fp16 output rounding (2^-11 relative)
        for name, got in (("two-pass", two), ("one-pass", one)):
            err = (got - want).abs().max().item()
            assert err < 4e-3 * max(1.0, want.abs().max().item()), (name, silu, err)
        assert ((one - two).abs() > 0).float().mean().item() < 2e-3, "only rare 1-ulp rounding flips may differ"


@pytest.mark.parametrize("n,hw,c,cout,parts", [(3, (32, 32), 128, 1, 4), (2, (32, 32), 128, 3, 1), (2, (64, 64), 128, 3, 2),
                                              (5, (28, 28), 128, 1, 1), (2, (16, 16), 256, 1, 3), (300, (32, 32), 128, 1, 2)])
def test_out_norm_conv_matches_torch(n, hw, c, cout, parts):
    """UNet tail (GroupNorm -> SiLU -> 3x3 conv to the image channels; generative DiffusionModelUNet.out) through
    ddpm_out_norm_conv: warp-MMA tap reduction + gather vs torch fp32 on the same fp16 activation. The normalised value
    is rounded to fp16 (as everywhere in the UNet), the weights stay fp32-accurate (hi + lo halves): 2e-3 of the output
    scale."""
    from ddpm_ood_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(n + c + cout)
    h, w_ = hw
    x = (torch.randn((n, h, w_, c), generator=g, device="cuda") * 1.3 + 0.2).half()
    # statistics partials as a producer would emit them: split the pixels into `parts` slices
    xs = x.float().reshape(n, h * w_, c // 4, 4)
    bounds = [round(i * h * w_ / parts) for i in range(parts + 1)]
    st = torch.stack([torch.stack([xs[:, a:b].sum(dim=(1, 3)), (xs[:, a:b] ** 2).sum(dim=(1, 3))], dim=-1)
                      for a, b in zip(bounds[:-1], bounds[1:])], dim=1).contiguous()
    gamma = 1 + 0.1 * torch.randn(c, generator=g, device="cuda")
    beta = 0.1 * torch.randn(c, generator=g, device="cuda")
    w = (torch.randn((cout, c, 3, 3), generator=g, device="cuda") * 0.05).contiguous()
    b = torch.randn(cout, generator=g, device="cuda")
    got = ops.out_norm_conv(x, st, gamma, beta, w, b, 32, 1e-6)
    z = F.silu(F.group_norm(x.float().permute(0, 3, 1, 2), 32, gamma, beta, 1e-6))
    want = F.conv2d(z.double(), w.double(), b.double(), padding=1).float()  # fp64: cuDNN fp32 may use TF32
    assert got.shape == want.shape
    err = (got - want).abs().max().item()
    assert err < 2e-3 * max(1.0, want.abs().max().item()), err


@pytest.mark.parametrize("n,cin,hw,cout", [(3, 1, (32, 32), 128), (2, 3, (32, 32), 128), (2, 3, (64, 64), 128),
                                           (5, 1, (28, 28), 128), (2, 1, (8, 8), 128), (2, 1, (32, 32), 256),
                                           (2, 1, (20, 12), 128), (300, 1, (32, 32), 128)])
def test_conv_in_matches_torch(n, cin, hw, cout):
    """UNet head (DiffusionModelUNet.conv_in) through ddpm_conv_in: fp32 image -> fp16 channels-last activation + the
    GroupNorm statistics of the next norm. The warp-MMA kernel splits both operands into fp16 hi + lo halves, so the only
    rounding is the fp16 output (2^-11 relative)."""
    from ddpm_ood_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(n + cin + cout)
    x = torch.randn((n, cin) + hw, generator=g, device="cuda") * 1.7
    w = (torch.randn((cout, cin, 3, 3), generator=g, device="cuda") * 0.3).contiguous()
    b = torch.randn(cout, generator=g, device="cuda")
    out, st = ops.conv_in(x, w, b)
    want = F.conv2d(x.double(), w.double(), b.double(), padding=1).permute(0, 2, 3, 1).float()  # fp64: cuDNN fp32 may use TF32
    err = ((out.float() - want).abs() / want.abs().clamp_min(1.0)).max().item()
    assert err < 6e-4, err
    assert st is not None and torch.isfinite(st).all(), "every statistics part must be written"
    want_s, want_q = _stats_from(out, st.shape[1])
    got = st.sum(dim=1)
    assert torch.allclose(got[..., 0], want_s, rtol=1e-5, atol=1e-2), (got[..., 0] - want_s).abs().max()
    assert torch.allclose(got[..., 1], want_q, rtol=1e-5, atol=1e-2), (got[..., 1] - want_q).abs().max()


@pytest.mark.parametrize("n,sp,c0,c1", [(37, (2, 2, 2), 256, 0), (5, (2, 2, 2), 256, 256), (9, (2, 2), 256, 0),
                                        (3, (1, 1), 512, 0), (600, (2, 2, 2), 256, 256)])
def test_groupnorm_small_maps(n, sp, c0, c1):
    """Maps of up to 8 pixels (the 2 x 2 x 2 level of the 3-D UNet): the one-thread-per-(image, group) kernel."""
    from ddpm_ood_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(n)
    xs = [(torch.randn((n,) + sp + (c,), generator=g, device="cuda") * 1.5 + 0.3).half() for c in (c0, c1) if c]
    C = c0 + c1
    gamma = 1 + 0.1 * torch.randn(C, generator=g, device="cuda")
    beta = 0.1 * torch.randn(C, generator=g, device="cuda")
    for silu in (True, False):
        want = _ref_gn(xs, gamma, beta, 32, 1e-6, silu)
        got = ops.gn_silu(xs[0], xs[1] if c1 else None, gamma, beta, 32, 1e-6, silu).float()
        err = (got - want).abs().max().item()
        assert err < 4e-3 * max(1.0, want.abs().max().item()), (silu, err)
