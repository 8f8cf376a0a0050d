"""Simplex noise: the oracle restatement against golden vectors made by running the reference's own
code (tests/golden/make_simplex_golden.py), and the CUDA kernel against both - bit for bit: the path is fp64 with the
reference's operand order, cast to fp32 at the end like the reference's tensor assignment."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import simplex as osx

GOLD = np.load(Path(__file__).parent / "golden" / "simplex_golden.npz")


def test_oracle_tables_match_reference_init():
    for seed, perm, gi3 in zip(GOLD["table_seeds"], GOLD["table_perm"], GOLD["table_grad_index3"]):
        p, g = osx.tables(int(seed))
        assert np.array_equal(p, perm)
        assert np.array_equal(g * 3, gi3)  # the reference stores the row offset into its flat gradient table


def test_oracle_noise3_is_bit_identical_to_reference():
    p, g = osx.tables(int(GOLD["noise3_seed"]))
    got = np.array([osx.noise3(x, y, z, p, g) for x, y, z in GOLD["noise3_points"]])
    assert np.array_equal(got, GOLD["noise3_values"])


def _seeds_like_reference(c, b, np_seed):
    """generate_simplex_noise: Simplex_CLASS() draws once, then one draw per (channel, image)"""
    np.random.seed(np_seed)
    np.random.randint(-10000000000, 10000000000)
    return np.array([[np.random.randint(-10000000000, 10000000000) for _ in range(b)] for _ in range(c)], dtype=np.int64)


@pytest.mark.parametrize("name", ["a", "b"])
def test_oracle_generate_matches_reference(name):
    b, c, h, w = (int(v) for v in GOLD[f"gen_{name}_shape"])
    got = osx.simplex_noise(_seeds_like_reference(c, b, 7), GOLD[f"gen_{name}_t"], (h, w))
    assert np.array_equal(got, GOLD[f"gen_{name}_noise"])


def test_oracle_against_live_reference_when_present():
    if not Path("/root/reference/src/utils/simplex_noise.py").exists():
        pytest.skip("reference tree not present (GPU box)")
    import sys

    sys.path.insert(0, str(Path(__file__).parent / "golden"))
    from make_simplex_golden import import_reference

    ref = import_reference(stub_numba=False)  # the reference's kernels under the real numba JIT
    perm, pgi = ref._init(-31337)
    p, g = osx.tables(-31337)
    rng = np.random.default_rng(5)
    for x, y, z in rng.uniform(-20, 20, (1500, 3)):
        assert ref._noise3(x, y, z, perm, pgi) == osx.noise3(x, y, z, p, g)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["a", "b"])
def test_device_generate_matches_reference_golden(name):
    """The drop-in call (`generate_simplex_noise(Simplex_CLASS(), x, t, in_channels=C)`, src/trainers/reconstruct.py:
    133-139) after np.random.seed(7): identical fp32 noise to the reference's."""
    from ddpm_ood_b200.simplex_noise import Simplex_CLASS, generate_simplex_noise

    shape = tuple(int(v) for v in GOLD[f"gen_{name}_shape"])
    np.random.seed(7)
    x = torch.zeros(shape, device="cuda")
    t = torch.tensor(GOLD[f"gen_{name}_t"]).long()
    got = generate_simplex_noise(Simplex_CLASS(), x, t, in_channels=shape[1])
    assert got.dtype == torch.float32 and got.shape == x.shape
    assert np.array_equal(got.cpu().numpy(), GOLD[f"gen_{name}_noise"])


@pytest.mark.gpu
def test_device_matches_oracle_at_workload_shape():
    """BASELINE config 1 geometry (1x32x32) and a 3-channel 28x28 case, t over the whole grid, against the oracle."""
    from ddpm_ood_b200.simplex_noise import Simplex_CLASS, generate_simplex_noise

    for (b, c, h, w), ts in (((3, 1, 32, 32), [10, 490, 970]), ((2, 3, 28, 28), [0, 990])):
        np.random.seed(11)
        x = torch.zeros((b, c, h, w), device="cuda")
        got = generate_simplex_noise(Simplex_CLASS(), x, torch.tensor(ts), in_channels=c).cpu().numpy()
        want = osx.simplex_noise(_seeds_like_reference(c, b, 11), ts, (h, w))
        assert np.array_equal(got, want)
    # statistics the trainer relies on: zero-centred, O(1) amplitude
    assert abs(float(got.mean())) < 0.5 and 0.05 < float(got.std()) < 1.0
