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


# ------------------------------------------------------------------------------------------------ the reference itself
T_GRID = [10, 170, 330, 490, 650, 810, 970, 980, 980, 990]  # a skip-1-style tail: t = 980 twice


def _write_run(tmp_path, seed, model="fashionmnist_synth"):
    """results_*.csv files shaped like a reconstruction run's (t-start outer, image inner, pandas index column, a few
    duplicated rows as the padded multi-GPU partition produces), plus the dense arrays they were made from."""
    import pandas as pd

    rng = np.random.default_rng(seed)
    ood_dir = tmp_path / model / "ood"
    ood_dir.mkdir(parents=True)
    dense = {}
    sets = [("val", "val", 40, 0.0), ("in", "in", 33, 0.0), ("MNIST", "out", 29, 0.6),
            ("FashionMNIST_vflip", "out", 21, 0.3), ("FashionMNIST_hflip", "out", 25, 0.1)]
    for name, typ, n, shift in sets:
        a = {k: (rng.gamma(2.0, 0.01 * (1 + shift), size=(len(T_GRID), n)) * (1 + np.arange(len(T_GRID)))[:, None])
             .astype(np.float32).astype(np.float64) for k in ("mse", "perceptual_difference")}
        rows = [{"filename": f"{name}_{i:04d}", "type": typ, "t": T_GRID[k], "perceptual_difference":
                 a["perceptual_difference"][k, i], "mse": a["mse"][k, i]} for k in range(len(T_GRID)) for i in range(n)]
        rows += [dict(r) for r in rows[:3]]
        pd.DataFrame(rows).to_csv(ood_dir / f"results_{name}.csv")
        dense[name] = a
    return model, dense


def _oracle_aucs(dense, min_t, max_t, target="mse"):
    from oracle import ood_scores

    rows = ood_scores.select_t(T_GRID, min_t, max_t)
    pick = lambda name: dense[name][target][rows]  # noqa: E731
    return [ood_scores.ood_auc(pick("val"), pick("in"), pick(o)) for o in ("MNIST", "FashionMNIST_vflip", "FashionMNIST_hflip")]


@pytest.mark.parametrize("min_t,max_t", [(0, 1000), (100, 985)])
def test_oracle_against_the_reference_script_itself(tmp_path, monkeypatch, min_t, max_t):
    """Executes /root/reference/ood_detection.py's own main() on synthetic CSVs (its third-party imports that are absent
    here - generative's PNDMScheduler, monai's print_config / set_determinism, matplotlib - are stubbed; none of them
    touches the score arithmetic) and compares every roc_auc_score it computes with oracle/ood_scores.py."""
    import importlib.util
    import sys
    import types
    from pathlib import Path

    ref_path = Path("/root/reference/ood_detection.py")
    if not ref_path.exists():
        pytest.skip("reference tree not present (GPU box)")
    from oracle.pndm import PNDMScheduler

    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        monkeypatch.setitem(sys.modules, name, m)
        return m

    mod("generative"); mod("generative.networks")
    mod("generative.networks.schedulers", PNDMScheduler=PNDMScheduler)
    mod("monai"); mod("monai.config", print_config=lambda: None); mod("monai.utils", set_determinism=lambda seed=None: None)
    plt = mod("matplotlib.pyplot", figure=lambda *a, **k: None, plot=lambda *a, **k: None, show=lambda *a, **k: None)
    mod("matplotlib", pyplot=plt)
    spec = importlib.util.spec_from_file_location("ref_ood_detection", ref_path)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    seen = []
    real = ref.roc_auc_score
    monkeypatch.setattr(ref, "roc_auc_score", lambda y, s: seen.append(real(y, s)) or seen[-1])

    model, dense = _write_run(tmp_path, seed=11)
    ref.main(types.SimpleNamespace(seed=2, output_dir=str(tmp_path), model_name=model, max_t=max_t, min_t=min_t, t_skip=1))
    want = _oracle_aucs(dense, min_t, max_t)
    assert len(seen) == 3
    for got, w in zip(seen, want):
        assert abs(got - w) < 1e-12, (seen, want)


@pytest.mark.gpu
@pytest.mark.parametrize("min_t,max_t", [(0, 1000), (100, 985)])
def test_root_ood_detection_cli_on_device(tmp_path, min_t, max_t):
    """The repo's ood_detection.py (reference flags, device kernels) against the oracle on the same CSV files, including
    the duplicated t = 980 column and the t filter."""
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    import ood_detection as cli

    model, dense = _write_run(tmp_path, seed=12)
    res = cli.main(cli.parse_args(["--output_dir", str(tmp_path), "--model_name", model, "--min_t", str(min_t),
                                   "--max_t", str(max_t)]))
    want = _oracle_aucs(dense, min_t, max_t)
    assert res["ood_data"] == ["MNIST", "FashionMNIST_vflip", "FashionMNIST_hflip"]
    for got, w in zip(res["Zscore_mse"], want):
        assert abs(got - w) < 2e-3, (res, want)  # fp32 z-scores: only near-tie pairs can flip


@pytest.mark.gpu
def test_device_ood_auc_dedupes_and_filters_t():
    from ddpm_ood_b200 import ood
    from oracle import ood_scores

    rng = np.random.default_rng(5)
    val, ins, outs = (rng.gamma(2.0, 0.01, size=(len(T_GRID), n)).astype(np.float32) for n in (40, 30, 20))
    outs *= 1.5
    rows = ood_scores.select_t(T_GRID, 100, 985)
    assert [T_GRID[i] for i in rows] == [170, 330, 490, 650, 810, 970, 980]
    want = ood_scores.ood_auc(val[rows].astype(np.float64), ins[rows].astype(np.float64), outs[rows].astype(np.float64))
    got = ood.ood_auc(*(torch.from_numpy(a).cuda() for a in (val, ins, outs)), t=torch.tensor(T_GRID), min_t=100, max_t=985)
    assert abs(got - want) < 2e-3
