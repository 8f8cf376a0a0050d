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
