"""Pins of the oracle's LOOP and WRAPPER restatements against the reference's own files EXECUTED in this container
(SURVEY.md 8c): `/root/reference/src/trainers/reconstruct.py::Reconstruct.get_scores` and
`/root/reference/src/losses/perceptual_loss.py::PerceptualLoss` are loaded from where they lie, with only their absent
third-party imports replaced - `generative.networks.schedulers.PNDMScheduler` and `lpips.LPIPS` by the oracle's
restatements (the arithmetic that stays unpinned), matplotlib and the data loader by no-ops - and must produce, row for
row, what `oracle/recon_loop.py` / `oracle/lpips.py::PerceptualLoss` produce from the same model, images and noise.
What this pins: the t-start grid as the reference slices it, ONE scheduler shared by all t-starts of a batch (the PLMS
history carries over), the SNR shift, `add_noise` on `images * b_scale`, the chain `timesteps[timesteps <= t_start]`,
`/ b_scale`, `clamp_(0, 1)`, the 28 -> 32 zero pad in front of LPIPS, the MSE reduction, the row order and the
`filename` stem of the CSV rows; and the wrapper's 2.5-D slicing where the LAST view overwrites the others.

Skipped where /root/reference does not exist (the GPU box)."""
import importlib.util
import sys
import types
from pathlib import Path
from unittest import mock

import pytest
import torch

REF = Path("/root/reference")
pytestmark = pytest.mark.skipif(not REF.exists(), reason="reference tree not present (GPU box)")


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


def _restore(saved, names):
    """Undo what a test put into sys.modules under `names` - and nothing else (modules that real packages imported
    lazily in the meantime must stay, or their classes exist twice)."""
    for k in names:
        if k in saved:
            sys.modules[k] = saved[k]
        else:
            sys.modules.pop(k, None)


def _load(name, path, package=None):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    if package:
        mod.__package__ = package
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture()
def reference_modules(monkeypatch):
    """The reference's perceptual_loss.py and trainers/reconstruct.py, imported under their own names with stubs for
    what is not installed. sys.modules is restored afterwards."""
    from oracle import lpips as olp
    from oracle import pndm as opn

    saved = dict(sys.modules)
    plt = mock.MagicMock()
    plt.subplots.return_value = (mock.MagicMock(), mock.MagicMock())
    stubs = {
        "lpips": _module("lpips", LPIPS=lambda **kw: olp.LPIPS()),  # same seeded weights as the oracle wrapper's
        "generative": _module("generative"),
        "generative.networks": _module("generative.networks"),
        "generative.networks.schedulers": _module("generative.networks.schedulers", PNDMScheduler=opn.PNDMScheduler),
        "matplotlib": _module("matplotlib", pyplot=plt),
        "matplotlib.pyplot": plt,
        "src": _module("src", __path__=[str(REF / "src")]),
        "src.data": _module("src.data", __path__=[]),
        "src.data.get_train_and_val_dataloader": _module("src.data.get_train_and_val_dataloader",
                                                         get_training_data_loader=lambda **kw: None),
        "src.utils": _module("src.utils", __path__=[]),
        "src.utils.simplex_noise": _module("src.utils.simplex_noise", generate_simplex_noise=None),
        "src.trainers": _module("src.trainers", __path__=[str(REF / "src" / "trainers")]),
        "src.trainers.base": _module("src.trainers.base", BaseTrainer=type("BaseTrainer", (), {})),
    }
    sys.modules.update(stubs)
    try:
        pl_mod = _load("src.losses.perceptual_loss", REF / "src" / "losses" / "perceptual_loss.py")
        sys.modules["src.losses"] = _module("src.losses", PerceptualLoss=pl_mod.PerceptualLoss, __path__=[])
        rec_mod = _load("src.trainers.reconstruct", REF / "src" / "trainers" / "reconstruct.py", package="src.trainers")
        yield pl_mod, rec_mod
    finally:
        _restore(saved, list(stubs) + ["src.losses", "src.losses.perceptual_loss", "src.trainers.reconstruct"])


class _Passthrough:
    def encode_stage_2_inputs(self, x):
        return x

    def decode_stage_2_outputs(self, x):
        return x


def _reference_rows(rec_mod, model, x0, names, skip, snr_shift=1.0, b_scale=1.0, seed=7):
    """Reconstruct.get_scores of the reference on one batch, the trainer built without its __init__ (checkpoints,
    loaders): only the attributes the method reads."""
    tr = object.__new__(rec_mod.Reconstruct)
    tr.model = model
    tr.device = torch.device("cpu")
    tr.prediction_type = "epsilon"
    tr.beta_schedule = "scaled_linear_beta"
    tr.beta_start, tr.beta_end = 0.0015, 0.0195
    tr.snr_shift = snr_shift
    tr.b_scale = b_scale
    tr.vqvae_model = _Passthrough()
    tr.do_latent_pad = False
    tr.simplex_noise = False
    tr.spatial_dimension = 2
    loader = [{"image": x0, "image_meta_dict": {"filename_or_obj": names}}]
    torch.manual_seed(seed)  # the reference draws torch.randn_like(images) per t-start from the global generator
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # torch.cuda.amp.autocast on a CPU-only build
        return tr.get_scores(loader, "val", skip)


@pytest.mark.parametrize("case", ["fmnist_32_skip32", "native_28_snr_shift_bscale"])
def test_reference_get_scores_equals_oracle_loop(reference_modules, case):
    from oracle import unet as ou
    from oracle.lpips import PerceptualLoss as OraclePL
    from oracle.recon_loop import LoopConfig, reconstruct_batch

    _, rec_mod = reference_modules
    if case == "fmnist_32_skip32":
        shape, skip, snr, bs = (2, 1, 32, 32), 32, 1.0, 1.0     # t-starts {10, 330, 650, 970}: 200 chained evaluations
    else:
        shape, skip, snr, bs = (2, 1, 28, 28), 64, 0.5, 0.7     # 28 -> 32 pad, SNR shift, b_scale; t-starts {10, 650}
    model = ou.randomize_(ou.make_small(2, 1), seed=0).eval()
    x0 = torch.rand(shape, generator=torch.Generator().manual_seed(3))
    names = ["/data/fmnist/img_0001.npy", "/data/brats/sub-02_t1.nii.gz"]
    rows = _reference_rows(rec_mod, model, x0, names, skip, snr_shift=snr, b_scale=bs)

    torch.manual_seed(7)  # same global draws, in the same order
    cfg = LoopConfig(inference_skip_factor=skip, snr_shift=snr, b_scale=bs, plms_state="carry")
    pl = OraclePL(dimensions=2, include_pixel_loss=False, is_fake_3d=False, lpips_normalize=True, spatial=False)
    want = reconstruct_batch(model, pl, x0, lambda i, t: torch.randn_like(x0), cfg)

    n_t, B = want["mse"].shape
    assert len(rows) == n_t * B
    assert [r["t"] for r in rows] == [int(t) for t in want["t"] for _ in range(B)]      # t-start outer, image inner
    assert [r["filename"] for r in rows[:B]] == ["img_0001", "sub-02_t1"]               # stem, ".nii" / ".gz" removed
    assert all(r["type"] == "val" for r in rows)
    got_m = torch.tensor([r["mse"] for r in rows], dtype=torch.float64).reshape(n_t, B)
    got_p = torch.tensor([r["perceptual_difference"] for r in rows], dtype=torch.float64).reshape(n_t, B)
    # same torch ops in the same order on the same inputs: identical up to nothing at all
    assert torch.equal(got_m, want["mse"].double())
    assert torch.equal(got_p, want["perceptual_difference"].double())
    assert got_m.min() > 0 and torch.isfinite(got_p).all()


class _Meta(torch.Tensor):
    """Stand-in for monai's MetaTensor: the reference's 2.5-D branch calls `.as_tensor()` on its slices."""

    def as_tensor(self):
        return self.as_subclass(torch.Tensor)


@pytest.mark.parametrize("shape", [(1, 1, 32, 32, 32), (2, 1, 33, 40, 36)])
def test_reference_perceptual_loss_wrapper_equals_oracle(reference_modules, shape):
    """The reference's PerceptualLoss, 2-D and 2.5-D (`is_fake_3d=True`: the loop over the three views assigns, it does
    not accumulate, so the result is the LAST view's mean LPIPS - perceptual_loss.py:110-122)."""
    from oracle.lpips import PerceptualLoss as OraclePL

    pl_mod, _ = reference_modules
    g = torch.Generator().manual_seed(11)
    y, p = torch.rand(shape, generator=g), torch.rand(shape, generator=g)
    ref3 = pl_mod.PerceptualLoss(dimensions=3, include_pixel_loss=False, is_fake_3d=True, lpips_normalize=True, spatial=False)
    ora3 = OraclePL(dimensions=3, include_pixel_loss=False, is_fake_3d=True, lpips_normalize=True, spatial=False)
    with torch.no_grad():
        got = ref3(y.as_subclass(_Meta), p.as_subclass(_Meta)).as_subclass(torch.Tensor)
        want = ora3(y, p)
    # the reference shuffles the slices before the mean (randperm with keep_ratio 1): same set, another summation order
    assert torch.allclose(got, want, rtol=1e-5, atol=0)
    # and the 2-D form on one slice stack
    ref2 = pl_mod.PerceptualLoss(dimensions=2, include_pixel_loss=False, is_fake_3d=False, lpips_normalize=True, spatial=False)
    ora2 = OraclePL(dimensions=2, include_pixel_loss=False, is_fake_3d=False, lpips_normalize=True, spatial=False)
    with torch.no_grad():
        assert torch.equal(ref2(y[:, :, 0], p[:, :, 0]), ora2(y[:, :, 0], p[:, :, 0]))


# ------------------------------------------------------------------------------------------------ BaseTrainer
class _Recorder(torch.nn.Module):
    """Stands in for a network class on both sides: keeps the constructor kwargs, has one parameter for the optimizer."""
    calls = []

    def __init__(self, **kw):
        super().__init__()
        type(self).calls.append(kw)
        self.w = torch.nn.Parameter(torch.zeros(1))


def _args(**over):
    import argparse

    base = dict(output_dir="/tmp/none", model_name="m", vqvae_checkpoint=None, is_grayscale=1, model_type="small",
                spatial_dimension=2, prediction_type="epsilon", beta_schedule="scaled_linear_beta", beta_start=0.0015,
                beta_end=0.0195, b_scale=1.0, snr_shift=1.0, simplex_noise=0, image_size=None, latent_pad=None,
                ddpm_checkpoint_epoch=None)
    base.update(over)
    return argparse.Namespace(**base)


@pytest.mark.parametrize("over", [
    dict(),
    dict(is_grayscale=0, model_type="big", image_size="64"),
    dict(spatial_dimension=3, snr_shift=0.3, b_scale=0.8, latent_pad=[0, 0, 2, 2, 1, 1], prediction_type="v_prediction"),
    dict(beta_schedule="linear_beta", beta_start=1e-4, beta_end=2e-2, snr_shift=4.0),
], ids=["defaults_small_gray", "big_colour_64", "3d_snr_shift_latent_pad_vpred", "linear_beta_snr_up"])
def test_reference_base_trainer_construction_equals_ours(monkeypatch, over):
    """`src/trainers/base.py:18-164` executed (its third-party classes replaced by recorders, its scheduler by the
    oracle's) next to `ddpm_ood_b200.trainers.base.BaseTrainer`: the UNet constructor kwargs, the trainer attributes the
    reconstruction loop reads, and the scheduler's betas / alphas / alphas_cumprod after the SNR shift must coincide."""
    from oracle import pndm as opn

    saved = dict(sys.modules)
    ref_unet = type("RefUNet", (_Recorder,), {"calls": []})
    stubs = {
        "generative": _module("generative"),
        "generative.inferers": _module("generative.inferers", DiffusionInferer=lambda s: ("inferer", s)),
        "generative.networks": _module("generative.networks"),
        "generative.networks.nets": _module("generative.networks.nets", VQVAE=_Recorder, DiffusionModelUNet=ref_unet),
        "generative.networks.schedulers": _module("generative.networks.schedulers", DDPMScheduler=opn.DDPMScheduler),
        "src": _module("src", __path__=[str(REF / "src")]),
        "src.utils": _module("src.utils", __path__=[]),
        "src.utils.simplex_noise": _module("src.utils.simplex_noise", Simplex_CLASS=lambda: "simplex"),
    }
    sys.modules.update(stubs)
    monkeypatch.delenv("LOCAL_RANK", raising=False)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    try:
        pv = _load("src.networks.passthrough_vqvae", REF / "src" / "networks" / "passthrough_vqvae.py")
        sys.modules["src.networks"] = _module("src.networks", PassthroughVQVAE=pv.PassthroughVQVAE, __path__=[])
        ref_base = _load("ref_base_trainer", REF / "src" / "trainers" / "base.py")
        ref = ref_base.BaseTrainer(_args(**over))
    finally:
        _restore(saved, list(stubs) + ["src.networks", "src.networks.passthrough_vqvae", "ref_base_trainer"])

    import ddpm_ood_b200.trainers.base as ours_base

    our_unet = type("OurUNet", (_Recorder,), {"calls": []})
    monkeypatch.setattr(ours_base, "DiffusionModelUNet", our_unet)
    monkeypatch.setattr(ours_base.BaseTrainer, "_init_device", lambda self: setattr(self, "device", torch.device("cpu"))
                        or setattr(self, "ddp", False))
    ours = ours_base.BaseTrainer(_args(**over))

    assert len(ref_unet.calls) == 1 and len(our_unet.calls) == 1
    assert ref_unet.calls[0] == our_unet.calls[0]
    for name in ("prediction_type", "beta_schedule", "beta_start", "beta_end", "b_scale", "snr_shift", "spatial_dimension",
                 "image_size", "do_latent_pad", "simplex_noise", "run_dir", "found_checkpoint", "start_epoch",
                 "global_step", "best_loss", "ddp"):
        assert getattr(ref, name) == getattr(ours, name), name
    if over.get("latent_pad"):
        assert ref.latent_pad == ours.latent_pad and ref.inverse_latent_pad == ours.inverse_latent_pad
    assert type(ours.vqvae_model).__name__ == type(ref.vqvae_model).__name__ == "PassthroughVQVAE"
    x = torch.rand(2, 1, 4, 4)
    assert torch.equal(ref.vqvae_model.encode_stage_2_inputs(x), ours.vqvae_model.encode_stage_2_inputs(x))
    assert torch.equal(ref.vqvae_model.decode_stage_2_outputs(x), ours.vqvae_model.decode_stage_2_outputs(x))
    for name in ("betas", "alphas", "alphas_cumprod"):
        a, b = getattr(ref.scheduler, name), getattr(ours.scheduler, name)
        assert torch.equal(a, b), name  # the reference's own SNR-shift loop ran on the left-hand side
    assert ref.scheduler.prediction_type == ours.scheduler.prediction_type


# ------------------------------------------------------------------------------------------------ Reconstruct.__init__ / reconstruct()
def _recon_args(tmp, **over):
    import argparse

    base = dict(batch_size=7, validation_ids="/d/fmnist_val.csv", in_ids="/d/fmnist_test.csv",
                out_ids="/d/mnist_test.csv,/d/fmnist_test_vflip.csv,/d/kmnist_test_hflip.csv", augmentation=1,
                num_workers=3, cache_data=0, drop_last=0, first_n=None, first_n_val="5", is_grayscale=1,
                spatial_dimension=2, image_roi=None, run_val=1, run_in=1, run_out=1, inference_skip_factor=8)
    base.update(over)
    return argparse.Namespace(**base)


def _fake_rows(loader, dataset_name, skip):
    return [{"filename": f"{dataset_name}{i}", "type": dataset_name, "t": 10 + 10 * skip * i,
             "perceptual_difference": 0.125 * i + len(loader[1]["validation_ids"]), "mse": 0.5 ** i} for i in range(4)]


def test_reference_reconstruct_orchestration_equals_ours(reference_modules, monkeypatch, tmp_path):
    """`Reconstruct.__init__` and `Reconstruct.reconstruct` of the reference executed next to ours, both with a recording
    loader factory and a canned `get_scores`: the same loader arguments for the val / in / out sets (flip variants
    included), the same (dataset_name, skip) sequence, and byte-identical `ood/results_*.csv` files under the same names."""
    _, rec_mod = reference_modules
    import ddpm_ood_b200.trainers.reconstruct as ours_mod

    outs = {}
    for side, mod in (("ref", rec_mod), ("ours", ours_mod)):
        calls, score_calls = [], []
        run_dir = tmp_path / side
        run_dir.mkdir()

        def base_init(self, args, run_dir=run_dir):
            self.found_checkpoint, self.run_dir, self.image_size = True, run_dir, None
            self.device, self.do_latent_pad = "cpu", False

        def factory(calls=calls, **kw):
            calls.append(kw)
            return ("loader", kw)

        def get_scores(self, loader, dataset_name, skip, score_calls=score_calls):
            score_calls.append((dataset_name, skip, loader[1]["validation_ids"]))
            return _fake_rows(loader, dataset_name, skip)

        monkeypatch.setattr(mod.BaseTrainer, "__init__", base_init, raising=False)
        monkeypatch.setattr(mod, "get_training_data_loader", factory)
        monkeypatch.setattr(mod.Reconstruct, "get_scores", get_scores)
        args = _recon_args(tmp_path)
        tr = mod.Reconstruct(args)
        tr.reconstruct(args)
        files = {p.name: p.read_bytes() for p in sorted((run_dir / "ood").iterdir())}
        outs[side] = (calls, score_calls, files)

    ref_calls, ref_scores, ref_files = outs["ref"]
    our_calls, our_scores, our_files = outs["ours"]
    assert len(ref_calls) == len(our_calls) == 5  # val, in, three out sets
    for r, o in zip(ref_calls, our_calls):
        assert {k: o[k] for k in r} == r                                  # every argument the reference passes, same value
        assert set(o) - set(r) <= {"rank", "world_size", "device"}        # ours adds only the sharding / ingest knobs
    assert ref_scores == our_scores
    assert [s[0] for s in ref_scores] == ["val", "in", "out", "out", "out"]
    assert sorted(ref_files) == ["results_fmnist_vflip.csv", "results_in.csv", "results_kmnist_hflip.csv",
                                 "results_mnist.csv", "results_val.csv"]
    assert ref_files == our_files


def test_reference_get_data_dicts_equals_ours(tmp_path, capsys):
    """`get_data_dicts` of the reference (src/data/get_train_and_val_dataloader.py:8-34; its monai imports stubbed, they
    are not touched without a process group) against ours: the one-row CSV of paths, `first_n`, the printed count."""
    saved = dict(sys.modules)
    monai = _module("monai", transforms=_module("monai.transforms"))
    data = _module("monai.data", CacheDataset=None, Dataset=None, ThreadDataLoader=None, partition_dataset=None)
    sys.modules.update({"monai": monai, "monai.transforms": monai.transforms, "monai.data": data})
    try:
        ref = _load("ref_loader", REF / "src" / "data" / "get_train_and_val_dataloader.py")
    finally:
        _restore(saved, ["monai", "monai.transforms", "monai.data", "ref_loader"])
    from ddpm_ood_b200 import data as ours

    paths = [f"/data/set/img_{i:03d}.npy" for i in range(11)] + ["/data/set/with space.nii.gz"]
    ids = tmp_path / "ids.csv"
    ids.write_text(",".join(paths) + "\n")
    for first_n in (False, 5, 12, 40):
        want = ref.get_data_dicts(str(ids), shuffle=False, first_n=first_n)
        out_ref = capsys.readouterr().out
        got = ours.get_data_dicts(str(ids), first_n=first_n)
        out_ours = capsys.readouterr().out
        assert got == want
        assert out_ours == out_ref  # "Found N subjects."


def test_oracle_alexnet_trunk_equals_torchvision():
    """lpips' `alexnet` feature extractor IS torchvision's `alexnet().features`, tapped after ReLU 1..5 (slices [0:2],
    [2:5], [5:8], [8:10], [10:12]) [3P-RECALL for the tap points]; the trunk itself is pinned here against the installed
    torchvision: same parameter shapes under the same indices, identical feature maps from shared weights."""
    tv = pytest.importorskip("torchvision")
    from oracle.lpips import AlexSlices

    net = tv.models.alexnet(weights=None).features.eval()
    ours = AlexSlices().eval()
    tv_params = {k: v for k, v in net.state_dict().items()}
    sd = {}
    for name, p in ours.state_dict().items():  # "slice2.3.weight" -> features "3.weight"
        key = name.split(".", 1)[1]
        assert tv_params[key].shape == p.shape, name
        sd[name] = tv_params[key]
    assert len(sd) == len(tv_params) == 10
    ours.load_state_dict(sd)
    x = torch.randn((2, 3, 64, 64), generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        got = ours(x)
        for k, end in enumerate((2, 5, 8, 10, 12)):
            assert torch.equal(got[k], net[:end](x)), k
