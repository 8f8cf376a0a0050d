"""GPU parity against the committed round-2 golden vectors (tests/golden/make_golden_r2.py, generated on CPU from the
fp32 oracle): the exact grids of the BASELINE.json configurations, which the oracle is too slow to re-run on the GPU box.

Tolerance (north_star): 1e-3 relative on the scores, t grid bit-exact. Per-forward eps tolerance: the CUDA path computes
with fp16 operands / fp32 accumulation through ~40 layers, the oracle in fp32 throughout: an eps element differs by up
to ~1e-3 of the eps range (fp16 has 11 bits; 40 layers of rounding): measured relative L2 error 0.86e-3 .. 1.04e-3 and
max|diff| / max|eps| 0.83e-3 .. 1.28e-3 over the five BASELINE shapes, asserted at 1.5e-3 / 3e-3; what the metric
reads - the scores after a whole chain - is held to 1e-3 (measured 2e-5 .. 1.9e-4 on the committed grids).
"""
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).parent / "golden"


def _models(sd, channels, seed, with_pl=True):
    from ddpm_ood_b200.losses import PerceptualLoss as OursPL
    from ddpm_ood_b200.networks import DiffusionModelUNet
    from oracle import unet as ou
    from oracle.lpips import PerceptualLoss as RefPL

    ref = ou.randomize_(ou.make_small(sd, channels), seed=seed)
    ours = DiffusionModelUNet(spatial_dims=sd, in_channels=channels, out_channels=channels,
                              num_channels=(128, 256, 256), attention_levels=(False, False, True), num_res_blocks=1,
                              num_head_channels=256, with_conditioning=False)
    ours.load_state_dict(ref.state_dict(), strict=True)
    ours = ours.to("cuda").eval()
    pl = None
    if with_pl:
        ref_pl = RefPL(dimensions=sd, include_pixel_loss=False, is_fake_3d=(sd == 3), lpips_normalize=True, spatial=False)
        pl = OursPL(dimensions=sd, include_pixel_loss=False, is_fake_3d=(sd == 3), lpips_normalize=True, spatial=False,
                    allow_synthetic_weights=True)
        pl.perceptual_function.load_lpips_state_dict(ref_pl.perceptual_function.state_dict())
        pl = pl.to("cuda")
    return ours, pl


def _check_scores(gold, got, keys=("mse", "perceptual_difference")):
    assert torch.equal(got["t"], gold["t"])  # bit-exact integer grid
    worst = {}
    for key in keys:
        w = gold[key]
        g_ = got[key].cpu()
        rel = ((g_ - w).abs() / w.abs().clamp_min(1e-12)).max().item()
        worst[key] = rel
        assert rel < 1e-3, (key, rel, w.flatten()[:8], g_.flatten()[:8])
    return worst


def _run(gold, sd, channels, with_pl=True, **cfg_kw):
    from ddpm_ood_b200.reconstruction import BatchReconstructor, ReconConfig

    ours, pl = _models(sd, channels, gold["weight_seed"], with_pl)
    cfg = ReconConfig(beta_schedule="scaled_linear_beta", beta_start=0.0015, beta_end=0.0195,
                      plms_state=gold["plms_state"], num_inference_steps=gold["num_inference_steps"],
                      spatial_dimension=sd, **cfg_kw)
    eng = BatchReconstructor(ours, pl, cfg, "cuda")
    return eng.score_batch(gold["x0"], gold["skip"], noise_fn=lambda i, t: gold["noise"][i].cuda(),
                           t_starts=gold["t_starts"])


def test_config2_benched_grid_carry():
    """BASELINE config 2 exactly as bench.py runs it: 1x32x32, skip 4, 25 t-starts, carry mode (the PLMS history leaks
    from chain to chain through all 1250 UNet evaluations)."""
    gold = torch.load(GOLDEN / "recon_cfg2_skip4_carry.pt")
    assert len(gold["t"]) == 25
    print(_check_scores(gold, _run(gold, 2, 1)))


def test_config4_celeba_skip1_first12():
    """BASELINE config 4: 3x64x64, skip 1 - the first 12 t-starts {10..120} of the 100, carry mode."""
    gold = torch.load(GOLDEN / "recon_cfg4_skip1_first12.pt")
    assert gold["t"].tolist() == [10 * (i + 1) for i in range(12)]
    print(_check_scores(gold, _run(gold, 2, 3)))


def test_config3_cifar_1000_steps_honoured():
    """BASELINE config 3: 3x32x32 with num_inference_steps=1000 honoured (the reference hard-codes 100,
    src/trainers/reconstruct.py:118), skip 100 -> t-starts {1, 101, ..., 901}, chains of up to 902 evaluations."""
    gold = torch.load(GOLDEN / "recon_cfg3_1000steps.pt")
    assert gold["t"].tolist() == [1 + 100 * i for i in range(10)]
    print(_check_scores(gold, _run(gold, 2, 3)))


def test_config5_latent_3d_chain():
    """BASELINE config 5's latent shape: [2,128,8,8,8] through the 3-D UNet, skip 32 -> 4 t-starts; MSE of the latent
    reconstruction (a 128-channel latent is not an LPIPS input; the image-space score needs the VQ-VAE, see
    test_vqvae_gpu.py)."""
    gold = torch.load(GOLDEN / "recon_cfg5_latent.pt")
    got = _run(gold, 3, 128, with_pl=False)
    print(_check_scores(gold, got, keys=("mse",)))


def test_lpips_3d_per_item():
    """The reference's 3-D scoring loop (src/trainers/reconstruct.py:181-187): PerceptualLoss(dimensions=3,
    is_fake_3d=True) per batch item, 2.5-D slices (src/losses/perceptual_loss.py:110-122)."""
    gold = torch.load(GOLDEN / "recon_cfg5_latent.pt")["lpips3d"]
    _, pl = _models(3, 128, 0)
    a, b = gold["a"].float().cuda(), gold["b"].float().cuda()
    got = torch.stack([pl(a[i, None], b[i, None]).reshape(()) for i in range(a.shape[0])]).cpu()
    assert torch.allclose(got, gold["pd"], rtol=2e-4, atol=1e-8), (got, gold["pd"])
    # the batched form the engine uses: all items' slices in one LPIPS call, mean per item
    batched = pl.per_item(a, b).cpu()
    assert torch.allclose(batched, got, rtol=1e-5, atol=1e-9), (batched, got)
    assert torch.allclose(pl.per_item(a, b, max_slices=1).cpu(), got, rtol=1e-5, atol=1e-9)  # one item per call


@pytest.mark.parametrize("key", ["1x32x32", "3x32x32", "3x64x64", "1x28x28", "128x8x8x8"])
def test_unet_eps_per_forward(key):
    """One UNet forward per BASELINE shape at t = {990, 10}: eps-level error, stated explicitly."""
    gold = torch.load(GOLDEN / "unet_eps_all_shapes.pt")
    case = gold["cases"][key]
    ours, _ = _models(case["sd"], case["channels"], gold["weight_seed"], with_pl=False)
    y = ours(case["x"].cuda(), case["t"].cuda()).cpu()
    want = case["y"]
    rel_l2 = ((y - want).norm() / want.norm()).item()
    max_rel = ((y - want).abs().max() / want.abs().max()).item()
    print(key, "rel-L2", rel_l2, "max|diff|/max|eps|", max_rel)
    # measured on B200 (round 2): rel-L2 0.86e-3 .. 1.04e-3, max|diff| / max|eps| 0.83e-3 .. 1.28e-3 over the five shapes
    assert rel_l2 < 1.5e-3, (key, rel_l2)
    assert max_rel < 3e-3, (key, max_rel)


def test_t_start_shards_union_equals_full_grid():
    """t-start sharding: in reset mode the union of the ranks' shares is the single-rank result - same t
    values bit for bit, scores to 1e-6 (the same kernels on the same inputs; only the order of chains differs)."""
    from ddpm_ood_b200.reconstruction import BatchReconstructor, ReconConfig, partition_t_starts
    from ddpm_ood_b200.synthetic import chain_lengths

    ours, pl = _models(2, 1, 0)
    cfg = ReconConfig(beta_schedule="scaled_linear_beta", beta_start=0.0015, beta_end=0.0195, plms_state="reset")
    eng = BatchReconstructor(ours, pl, cfg, "cuda")
    skip = 16
    x0 = torch.rand((3, 1, 32, 32), generator=torch.Generator().manual_seed(9))
    noise = [torch.randn((3, 1, 32, 32), generator=torch.Generator().manual_seed(50 + i)).cuda() for i in range(7)]
    full = eng.score_batch(x0, skip, noise_fn=lambda i, t: noise[i])
    parts = partition_t_starts(chain_lengths(100, skip), 2)
    merged_t = torch.empty_like(full["t"])
    merged = {k: torch.empty_like(full[k]) for k in ("mse", "perceptual_difference")}
    for r in range(2):
        got = eng.score_batch(x0, skip, noise_fn=lambda i, t: noise[i], t_indices=parts[r])
        assert got["t_index"] == parts[r]
        for j, i in enumerate(parts[r]):
            merged_t[i] = got["t"][j]
            for k in merged:
                merged[k][i] = got[k][j]
    assert torch.equal(merged_t, full["t"])
    for k in merged:
        assert torch.allclose(merged[k], full[k], rtol=1e-6, atol=0), k
    carry = BatchReconstructor(ours, pl, ReconConfig(beta_schedule="scaled_linear_beta", beta_start=0.0015,
                                                      beta_end=0.0195, plms_state="carry"), "cuda")
    with pytest.raises(ValueError):
        carry.score_batch(x0, skip, t_indices=[0])


FULL_SIZE = {
    # name: golden file, spatial dims, channels, batch bench.py runs the workload at, score keys
    "config2_fmnist_b1184": ("recon_cfg2_skip4_carry.pt", 2, 1, 1184, ("mse", "perceptual_difference")),
    "config4_celeba64_b296": ("recon_cfg4_skip1_first12.pt", 2, 3, 296, ("mse", "perceptual_difference")),
    "config5_latent_b592": ("recon_cfg5_latent.pt", 3, 128, 592, ("mse",)),
}


@pytest.mark.parametrize("name", list(FULL_SIZE))
def test_full_bench_batch_by_replication(name):
    """The BASELINE configurations at the batch bench.py measures them at (every UNet level a whole number of waves of
    the default tilings - kernels the small parity batches never launch). A chain only couples an image to itself, so
    the golden's images tiled to the bench batch (noise tiled the same way) must give, for EVERY (image, t-start) pair of
    the big batch, the oracle's score of the image it replicates: 1e-3 relative (north_star), the t grid bit-exact, and
    the replicas of one image agree with each other to fp32 reduction noise. config 2 runs its whole benched grid
    (25 t-starts, carry: 1250 UNet evaluations at batch 1184 - exactly one bench step)."""
    from ddpm_ood_b200.reconstruction import BatchReconstructor, ReconConfig

    fname, sd, ch, batch, keys = FULL_SIZE[name]
    gold = torch.load(GOLDEN / fname)
    base = gold["x0"].shape[0]
    reps = batch // base
    assert base * reps == batch
    ours, pl = _models(sd, ch, gold["weight_seed"], with_pl=len(keys) > 1)
    cfg = ReconConfig(beta_schedule="scaled_linear_beta", beta_start=0.0015, beta_end=0.0195,
                      plms_state=gold["plms_state"], num_inference_steps=gold["num_inference_steps"],
                      spatial_dimension=sd)
    eng = BatchReconstructor(ours, pl, cfg, "cuda")
    tile = (reps,) + (1,) * (gold["x0"].dim() - 1)
    x0 = gold["x0"].repeat(tile)  # image b of the big batch replicates golden image b % base
    got = eng.score_batch(x0, gold["skip"], noise_fn=lambda i, t: gold["noise"][i].repeat(tile).cuda(),
                          t_starts=gold["t_starts"])
    assert torch.equal(got["t"], gold["t"])
    n_t = len(gold["t"])
    for key in keys:
        g_ = got[key].cpu().reshape(n_t, reps, base)   # [t, replica, golden image]
        w = gold[key].reshape(n_t, 1, base)
        rel = ((g_ - w).abs() / w.abs().clamp_min(1e-12)).max().item()
        assert rel < 1e-3, (name, key, rel)
        spread = ((g_ - g_[:, :1]).abs() / g_[:, :1].abs().clamp_min(1e-12)).max().item()
        assert spread < 1e-5, (name, key, spread)
