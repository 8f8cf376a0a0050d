"""z-score / per-file mean / ROC-AUC post-processing.
CPU: oracle/ood_scores.py against the reference's own arithmetic (ood_detection.py:150-206 restated with the same
pandas / sklearn calls on the same rows, including the duplicate removal of :53-54,143-145).
GPU: the device kernels (ddpm_val_stats / ddpm_mean_z / ddpm_auc_counts) against the oracle."""
import numpy as np
import pytest
import torch


def _fake_scores(seed, n_t=7, n_val=40, n_in=33, n_out=29, shift=0.6):
    rng = np.random.default_rng(seed)
    val = rng.gamma(2.0, 0.01, size=(n_t, n_val)) * (1 + np.arange(n_t))[:, None]
    ins = rng.gamma(2.0, 0.01, size=(n_t, n_in)) * (1 + np.arange(n_t))[:, None]
    outs = rng.gamma(2.0, 0.01 * (1 + shift), size=(n_t, n_out)) * (1 + np.arange(n_t))[:, None]
    # fp32 like the device tensors, and a few exact ties between an in and an out image
    val, ins, outs = (a.astype(np.float32).astype(np.float64) for a in (val, ins, outs))
    outs[:, 0] = ins[:, 0]
    outs[:, 1] = ins[:, 5]
    return val, ins, outs


def _reference_route(val, ins, outs, target="mse"):
    """ood_detection.py:150-206 with pandas / sklearn on CSV-shaped rows (t-start outer, image inner; one padded
    duplicate per set as even_divisible sharding produces)."""
    import pandas as pd
    from sklearn.metrics import roc_auc_score

    ts = [10 + 160 * i for i in range(val.shape[0])]

    def frame(a, typ, prefix):
        rows = [{"filename": f"{prefix}{i:04d}", "type": typ, "t": ts[k], target: a[k, i]}
                for k in range(a.shape[0]) for i in range(a.shape[1])]
        rows += [dict(r) for r in rows[:3]]  # duplicates from the padded partition
        return pd.DataFrame(rows)

    df_val, df_in, df_out = frame(val, "val", "v"), frame(ins, "in", "i"), frame(outs, "out", "o")
    for df in (df_val, df_in, df_out):
        df.drop_duplicates(subset=["filename", "t"], keep="first", inplace=True)
    results = pd.concat((df_in, df_out))
    agg = (df_val.groupby(["t"]).agg({target: ["mean", "std"]})[target].reset_index()
           .rename({"mean": "val_mean", "std": "val_std"}, axis=1))
    results = results.merge(agg, on=["t"], how="left")
    results["z"] = (results[target] - results["val_mean"]) / results["val_std"]
    mean = results.groupby(["filename", "type"])[["z"]].mean().reset_index()
    scores = mean.loc[mean["type"] == "in"]["z"].tolist() + mean.loc[mean["type"] == "out"]["z"].tolist()
    classes = [0] * int((mean["type"] == "in").sum()) + [1] * int((mean["type"] == "out").sum())
    return roc_auc_score(classes, scores)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_oracle_matches_reference_pandas_sklearn_route(seed):
    from oracle import ood_scores

    val, ins, outs = _fake_scores(seed)
    want = _reference_route(val, ins, outs)
    got = ood_scores.ood_auc(val, ins, outs)
    assert abs(got - want) < 1e-12, (got, want)
    assert 0.5 < got < 1.0


@pytest.mark.gpu
@pytest.mark.parametrize("seed,n_in,n_out", [(0, 33, 29), (3, 2500, 1700), (4, 1, 1)])
def test_device_post_processing_matches_oracle(seed, n_in, n_out):
    from ddpm_ood_b200 import ood
    from oracle import ood_scores

    val, ins, outs = _fake_scores(seed, n_in=max(n_in, 6), n_out=max(n_out, 2))
    ins, outs = ins[:, :n_in], outs[:, :n_out]
    dv, di, do = (torch.from_numpy(a).float().cuda() for a in (val, ins, outs))
    m, s = ood.val_stats(dv)
    wm, ws = ood_scores.val_stats(val)
    assert np.allclose(m.cpu().numpy(), wm, rtol=1e-6)
    assert np.allclose(s.cpu().numpy(), ws, rtol=1e-6)
    zi = ood.mean_z(di, m, s).cpu().numpy()
    assert np.allclose(zi, ood_scores.mean_z(ins, wm, ws), rtol=1e-4, atol=1e-5)
    # the pair counts are exact integers: AUC of the DEVICE z-scores must equal the oracle's AUC of the same numbers
    zo = ood.mean_z(do, m, s)
    got = ood.roc_auc(torch.from_numpy(zi).cuda(), zo)
    assert got == ood_scores.roc_auc(zi, zo.cpu().numpy())
    # and end to end it agrees with the float64 route up to near-tie flips
    assert abs(ood.ood_auc(dv, di, do) - ood_scores.ood_auc(val, ins, outs)) < 2e-3


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.uint8, torch.float32])
def test_device_intensity_scaling_matches_host_transform(dtype):
    """per-image min-max scaling on the device equals the host loader's transform
    (ddpm_ood_b200/data.py:_transform, ScaleIntensityd of get_train_and_val_dataloader.py:76), incl. a constant image."""
    from ddpm_ood_b200.data import _transform, scale_intensity_on_device

    g = torch.Generator().manual_seed(3)
    raw = torch.randint(0, 256, (5, 1, 28, 28), generator=g).to(dtype)
    if dtype == torch.float32:
        raw = raw * 0.37 - 11.0
    raw[3] = raw[3, 0, 0, 0]  # constant image
    want = torch.stack([_transform(r.numpy(), is_grayscale=True, spatial_dimension=2, image_size=None, image_roi=None,
                                   add_vflip=False, add_hflip=False) for r in raw])
    got = scale_intensity_on_device(raw.cuda()).cpu()
    assert torch.allclose(got, want, rtol=0, atol=1e-7)
    assert float(got[3].abs().max()) == 0.0


@pytest.mark.gpu
def test_loader_device_ingest_matches_host_loader(tmp_path):
    """The loader's GPU-resident ingest (uint8 over PCIe, scaling + flips on the device) yields the host loader's batches."""
    import numpy as np

    from ddpm_ood_b200.data import SimpleLoader

    rng = np.random.default_rng(0)
    dicts = []
    for i in range(5):
        np.save(tmp_path / f"{i}.npy", rng.integers(0, 256, (28, 28), dtype=np.uint8))
        dicts.append({"image": str(tmp_path / f"{i}.npy")})
    tf = dict(is_grayscale=True, spatial_dimension=2, image_size=None, image_roi=[24, 24], add_vflip=True, add_hflip=True)
    host = list(SimpleLoader(dicts, 2, False, **tf))
    dev = list(SimpleLoader(dicts, 2, False, device="cuda", **tf))
    assert len(host) == len(dev) == 3
    for h, d in zip(host, dev):
        assert d["image"].is_cuda and d["image"].shape == h["image"].shape
        assert torch.allclose(d["image"].cpu(), h["image"], rtol=0, atol=1e-7)
        assert d["image_meta_dict"] == h["image_meta_dict"]
