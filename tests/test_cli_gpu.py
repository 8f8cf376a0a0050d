"""End to end through the reference-facing CLI surface (root reconstruct.py -> trainers.Reconstruct, the mirror of
reconstruct.py:91-96 + src/trainers/reconstruct.py:96-335): .npy images + split files + checkpoint.pth in, results_*.csv
out, and the CSV contract ood_detection.py:150-206 consumes (columns, row order t-start outer / image inner, every
out-of-distribution set typed "out", flip datasets by file-name suffix)."""
import sys
from pathlib import Path

import numpy as np
import pandas as pd
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def _dataset(tmp_path, name, n, seed):
    rng = np.random.default_rng(seed)
    d = tmp_path / name
    d.mkdir()
    paths = []
    for i in range(n):
        np.save(d / f"{name}_{i}.npy", rng.integers(0, 256, (32, 32), dtype=np.uint8))
        paths.append(str(d / f"{name}_{i}.npy"))
    ids = tmp_path / f"{name}_ids.csv"
    ids.write_text(",".join(paths) + "\n")  # one csv row of paths (get_train_and_val_dataloader.py:10-17)
    return ids, paths


def test_cli_writes_the_reference_csv_contract(tmp_path, monkeypatch):
    monkeypatch.setenv("DDPM_LPIPS_ALLOW_SYNTHETIC", "1")  # no lpips package / weights on the box: explicit opt-in
    sys.path.insert(0, str(ROOT))
    import reconstruct as cli
    from ddpm_ood_b200.data import SimpleLoader
    from ddpm_ood_b200.trainers import Reconstruct
    from oracle import unet as ou

    val_ids, val_paths = _dataset(tmp_path, "val", 5, 0)
    in_ids, _ = _dataset(tmp_path, "in", 2, 1)
    out_ids, out_paths = _dataset(tmp_path, "other", 3, 2)
    run = tmp_path / "runs" / "m"
    run.mkdir(parents=True)
    weights = ou.randomize_(ou.make_small(2, 1), seed=5).state_dict()
    torch.save({"epoch": 7, "global_step": 1, "best_loss": 0.5, "model_state_dict": weights}, run / "checkpoint.pth")

    out_flag = str(out_ids).replace(".csv", "_vflip.csv")  # "<set>_vflip" names the same ids file, flipped
    args = cli.parse_args([
        "--output_dir", str(tmp_path / "runs"), "--model_name", "m", "--validation_ids", str(val_ids),
        "--in_ids", str(in_ids), "--out_ids", f"{out_ids},{out_flag}", "--is_grayscale", "1", "--batch_size", "3",
        "--beta_schedule", "scaled_linear_beta", "--beta_start", "0.0015", "--beta_end", "0.0195",
        "--inference_skip_factor", "64"])
    recon = Reconstruct(args)
    torch.manual_seed(11)
    recon.reconstruct(args)

    ood = run / "ood"
    assert sorted(p.name for p in ood.glob("*.csv")) == ["results_in.csv", "results_other.csv",
                                                        "results_other_vflip.csv", "results_val.csv"]
    val = pd.read_csv(ood / "results_val.csv", index_col=0)
    assert list(val.columns) == ["filename", "type", "t", "perceptual_difference", "mse"]
    # 5 images in batches of 3 and 2; per batch: t-start outer (10, 650 at skip 64), image inner
    stems = [Path(p).stem for p in val_paths]
    assert list(val["filename"]) == stems[:3] * 2 + stems[3:] * 2
    assert list(val["t"]) == [10] * 3 + [650] * 3 + [10] * 2 + [650] * 2
    assert set(val["type"]) == {"val"}
    assert np.isfinite(val[["perceptual_difference", "mse"]].to_numpy()).all()
    # more noise, worse reconstruction
    assert (val[val.t == 650].mse.mean() > val[val.t == 10].mse.mean())
    for name in ("results_other.csv", "results_other_vflip.csv"):
        df = pd.read_csv(ood / name, index_col=0)
        assert set(df["type"]) == {"out"} and list(df["filename"]) == [Path(p).stem for p in out_paths] * 2

    # the values are the engine's: same seed, the HOST loader (the CLI ingested on the device), direct score_batch calls
    torch.manual_seed(11)
    engine = recon._engine()
    rows = []
    tf = dict(is_grayscale=True, spatial_dimension=2, image_size=None, image_roi=None, add_vflip=False, add_hflip=False)
    for batch in SimpleLoader([{"image": p} for p in val_paths], 3, False, **tf):
        res = engine.score_batch(batch["image"], 64)
        for i in range(len(res["t"])):
            rows += [(float(a), float(b)) for a, b in zip(res["perceptual_difference"][i].cpu(), res["mse"][i].cpu())]
    want = np.array(rows)
    got = val[["perceptual_difference", "mse"]].to_numpy()
    assert np.allclose(got, want, rtol=1e-5, atol=0)


def test_cli_refuses_made_up_lpips_weights(tmp_path, monkeypatch):
    """Without real LPIPS weights (and without the explicit opt-in) the production path raises instead of writing
    meaningless perceptual_difference columns into a reference-shaped CSV."""
    try:
        import lpips  # noqa: F401
        pytest.skip("lpips is installed: real weights load")
    except ImportError:
        pass
    monkeypatch.delenv("DDPM_LPIPS_ALLOW_SYNTHETIC", raising=False)
    monkeypatch.delenv("DDPM_LPIPS_STATE_DICT", raising=False)
    sys.path.insert(0, str(ROOT))
    import reconstruct as cli
    from ddpm_ood_b200._lib import DdpmError
    from ddpm_ood_b200.trainers import Reconstruct
    from oracle import unet as ou

    val_ids, _ = _dataset(tmp_path, "val", 2, 0)
    run = tmp_path / "runs" / "m"
    run.mkdir(parents=True)
    torch.save({"epoch": 1, "global_step": 1, "best_loss": 0.5,
                "model_state_dict": ou.randomize_(ou.make_small(2, 1), seed=5).state_dict()}, run / "checkpoint.pth")
    args = cli.parse_args(["--output_dir", str(tmp_path / "runs"), "--model_name", "m", "--validation_ids", str(val_ids),
                           "--in_ids", str(val_ids), "--out_ids", str(val_ids), "--run_in", "0", "--run_out", "0",
                           "--is_grayscale", "1", "--batch_size", "2", "--inference_skip_factor", "64"])
    recon = Reconstruct(args)
    with pytest.raises(DdpmError, match="LPIPS weights not found"):
        recon.reconstruct(args)
    assert not (run / "ood" / "results_val.csv").exists()


def test_cli_simplex_noise_mode(tmp_path, monkeypatch):
    monkeypatch.setenv("DDPM_LPIPS_ALLOW_SYNTHETIC", "1")
    """--simplex_noise=1 (reference reconstruct.py:83-88, trainers/reconstruct.py:133-139): the CLI's scores are the
    engine's with generate_simplex_noise as the noise source under the same numpy seed."""
    sys.path.insert(0, str(ROOT))
    import reconstruct as cli
    from ddpm_ood_b200.data import SimpleLoader
    from ddpm_ood_b200.simplex_noise import Simplex_CLASS, generate_simplex_noise
    from ddpm_ood_b200.trainers import Reconstruct
    from oracle import unet as ou

    val_ids, val_paths = _dataset(tmp_path, "val", 3, 0)
    run = tmp_path / "runs" / "m"
    run.mkdir(parents=True)
    weights = ou.randomize_(ou.make_small(2, 1), seed=5).state_dict()
    torch.save({"epoch": 7, "global_step": 1, "best_loss": 0.5, "model_state_dict": weights}, run / "checkpoint.pth")
    args = cli.parse_args([
        "--output_dir", str(tmp_path / "runs"), "--model_name", "m", "--validation_ids", str(val_ids),
        "--in_ids", str(val_ids), "--out_ids", str(val_ids), "--run_in", "0", "--run_out", "0", "--is_grayscale", "1",
        "--batch_size", "3", "--beta_schedule", "scaled_linear_beta", "--beta_start", "0.0015", "--beta_end", "0.0195",
        "--inference_skip_factor", "64", "--simplex_noise", "1"])
    np.random.seed(3)
    recon = Reconstruct(args)  # Simplex_CLASS() draws its first seed here, like the reference's BaseTrainer
    recon.reconstruct(args)
    val = pd.read_csv(run / "ood" / "results_val.csv", index_col=0)
    assert list(val["t"]) == [10] * 3 + [650] * 3 and np.isfinite(val[["perceptual_difference", "mse"]].to_numpy()).all()

    np.random.seed(3)
    simplex = Simplex_CLASS()
    engine = recon._engine()
    tf = dict(is_grayscale=True, spatial_dimension=2, image_size=None, image_roi=None, add_vflip=False, add_hflip=False)
    batch = next(iter(SimpleLoader([{"image": p} for p in val_paths], 3, False, **tf)))["image"]
    probe = torch.empty(tuple(batch.shape), device="cuda")
    res = engine.score_batch(batch, 64, noise_fn=lambda i, t: generate_simplex_noise(
        simplex, x=probe, t=torch.full((3,), int(t), dtype=torch.long), in_channels=1))
    want = np.stack([res["perceptual_difference"].cpu().numpy().reshape(-1), res["mse"].cpu().numpy().reshape(-1)], axis=1)
    assert np.allclose(val[["perceptual_difference", "mse"]].to_numpy(), want, rtol=1e-5, atol=0)


def test_cli_latent_diffusion_with_vqvae_checkpoint(tmp_path, monkeypatch):
    """--vqvae_checkpoint (reference src/trainers/base.py:44-61): the stage-1 VQ-VAE is built from vqvae_config.json next
    to its checkpoint, the DDPM gets embedding_dim channels, 3-D volumes are encoded once per batch and every
    reconstructed latent is decoded before scoring (src/trainers/reconstruct.py:124,166); per-item 2.5-D LPIPS."""
    import json

    monkeypatch.setenv("DDPM_LPIPS_ALLOW_SYNTHETIC", "1")
    sys.path.insert(0, str(ROOT))
    import reconstruct as cli
    from ddpm_ood_b200.data import SimpleLoader
    from ddpm_ood_b200.trainers import Reconstruct
    from ddpm_ood_b200.vqvae import VQVAE
    from oracle import unet as ou
    from oracle import vqvae as ov

    rng = np.random.default_rng(0)
    d = tmp_path / "vol"
    d.mkdir()
    paths = []
    for i in range(3):
        np.save(d / f"vol_{i}.npy", rng.integers(0, 256, (32, 32, 32), dtype=np.uint8))
        paths.append(str(d / f"vol_{i}.npy"))
    ids = tmp_path / "vol_ids.csv"
    ids.write_text(",".join(paths) + "\n")

    vq_cfg = dict(spatial_dims=3, in_channels=1, out_channels=1, num_channels=[128, 256], num_res_layers=1,
                  num_res_channels=[128, 256], downsample_parameters=[[2, 4, 1, 1]] * 2,
                  upsample_parameters=[[2, 4, 1, 1, 0]] * 2, num_embeddings=256, embedding_dim=128)
    vq_dir = tmp_path / "runs" / "vqvae"
    vq_dir.mkdir(parents=True)
    (vq_dir / "vqvae_config.json").write_text(json.dumps(vq_cfg))
    torch.save({"model_state_dict": ov.randomize_(ov.VQVAE(**vq_cfg), seed=0).state_dict()}, vq_dir / "checkpoint.pth")
    run = tmp_path / "runs" / "ldm"
    run.mkdir(parents=True)
    torch.save({"epoch": 1, "global_step": 1, "best_loss": 0.5,
                "model_state_dict": ou.randomize_(ou.make_small(3, 128), seed=5).state_dict()}, run / "checkpoint.pth")
    args = cli.parse_args([
        "--output_dir", str(tmp_path / "runs"), "--model_name", "ldm", "--validation_ids", str(ids), "--in_ids", str(ids),
        "--out_ids", str(ids), "--run_in", "0", "--run_out", "0", "--is_grayscale", "1", "--spatial_dimension", "3",
        "--batch_size", "2", "--vqvae_checkpoint", str(vq_dir / "checkpoint.pth"), "--beta_schedule", "scaled_linear_beta",
        "--beta_start", "0.0015", "--beta_end", "0.0195", "--inference_skip_factor", "64"])
    recon = Reconstruct(args)
    assert isinstance(recon.vqvae_model, VQVAE) and recon.model.in_channels == 128
    torch.manual_seed(11)
    recon.reconstruct(args)
    val = pd.read_csv(run / "ood" / "results_val.csv", index_col=0)
    stems = [Path(p).stem for p in paths]
    assert list(val["filename"]) == stems[:2] * 2 + stems[2:] * 2
    assert list(val["t"]) == [10] * 2 + [650] * 2 + [10] + [650]
    assert np.isfinite(val[["perceptual_difference", "mse"]].to_numpy()).all()

    torch.manual_seed(11)
    engine = recon._engine()
    assert engine.vqvae_model is recon.vqvae_model
    rows = []
    tf = dict(is_grayscale=True, spatial_dimension=3, image_size=None, image_roi=None, add_vflip=False, add_hflip=False)
    for batch in SimpleLoader([{"image": p} for p in paths], 2, False, **tf):
        res = engine.score_batch(batch["image"], 64)
        for i in range(len(res["t"])):
            rows += [(float(a), float(b)) for a, b in zip(res["perceptual_difference"][i].cpu(), res["mse"][i].cpu())]
    assert np.allclose(val[["perceptual_difference", "mse"]].to_numpy(), np.array(rows), rtol=1e-5, atol=0)
