"""CPU tests of the host-side scheduler logic (timestep grid, PLMS bookkeeping, per-step coefficients).

The device kernel `plms_apply` (csrc/kernels.cu) is emulated here in a few lines of torch so the coefficients the host
hands to it can be checked against the oracle without a GPU; the real kernel is checked in test_unet_gpu.py.
"""
import pytest
import torch

from ddpm_ood_b200.schedulers import PNDMScheduler, DDPMScheduler, make_betas
from oracle import pndm as op

README_TABLE = {1: 100, 2: 50, 3: 34, 4: 25, 5: 20, 8: 13, 16: 7, 32: 4, 64: 2}  # reference README.md:118-120


def _emulate(st, eps, ring, stash, sample):
    eb = st.c[0] * eps
    for k in range(3):
        if st.c[k + 1] != 0.0:
            eb = eb + st.c[k + 1] * ring[st.slot[k]]
    s = stash.clone() if st.use_stash else sample
    if st.write_stash:
        stash.copy_(sample)
    mo = st.vA * eb + st.vB * s
    out = st.A * s - st.Bc * mo
    if st.push:
        ring[st.slot_new] = eps
    return out


@pytest.mark.parametrize("k", sorted(README_TABLE))
def test_t_start_grid_matches_readme_table(k):
    s = PNDMScheduler(num_train_timesteps=1000, skip_prk_steps=True)
    s.set_timesteps(100)
    assert len(s.timesteps) == 101
    starts = reversed(s.timesteps)[1::k]
    assert len(starts) == README_TABLE[k]
    o = op.PNDMScheduler(num_train_timesteps=1000, skip_prk_steps=True)
    o.set_timesteps(100)
    assert torch.equal(s.timesteps, o.timesteps)
    assert torch.equal(starts, op.t_start_grid(o.timesteps, k))
    assert s.timesteps.dtype == torch.int64


def test_grid_values():
    s = PNDMScheduler(num_train_timesteps=1000, skip_prk_steps=True)
    s.set_timesteps(100)
    assert s.timesteps[:4].tolist() == [990, 980, 980, 970]
    assert s.timesteps[-3:].tolist() == [20, 10, 0]
    assert reversed(s.timesteps)[1::16].tolist() == [10, 170, 330, 490, 650, 810, 970]
    # boolean mask + iteration as in trainers/reconstruct.py:149
    assert [int(t) for t in s.timesteps[s.timesteps <= 30]] == [30, 20, 10, 0]


@pytest.mark.parametrize("schedule", ["linear_beta", "scaled_linear_beta", "linear", "scaled_linear"])
def test_betas_match_oracle(schedule):
    a = make_betas(schedule, 1000, 0.0015, 0.0195)
    b = op.make_betas(schedule, 1000, 0.0015, 0.0195)
    assert torch.equal(a, b)
    d = DDPMScheduler(num_train_timesteps=1000, schedule=schedule, beta_start=0.0015, beta_end=0.0195,
                      prediction_type="epsilon")
    assert torch.equal(d.alphas_cumprod, torch.cumprod(1 - b, 0))


@pytest.mark.parametrize("pred", ["epsilon", "v_prediction"])
@pytest.mark.parametrize("mode", ["carry", "reset"])
def test_plms_coefficients_reproduce_oracle(pred, mode):
    kw = dict(num_train_timesteps=1000, skip_prk_steps=True, schedule="scaled_linear_beta", beta_start=0.0015,
              beta_end=0.0195, prediction_type=pred)
    ref, ours = op.PNDMScheduler(**kw), PNDMScheduler(**kw)
    ref.set_timesteps(100)
    ours.set_timesteps(100)
    g = torch.Generator().manual_seed(0)
    shape = (2, 1, 4, 4)
    ring = torch.zeros((4,) + shape)
    stash = torch.zeros(shape)
    for t_start in (10, 20, 980, 990, 50):
        if mode == "reset":
            ref.reset_chain()
            ours.reset_chain()
        xa = torch.randn(shape, generator=g)
        xb = xa.clone()
        for step in ref.timesteps[ref.timesteps <= t_start]:
            eps = torch.randn(shape, generator=g)
            xa, _ = ref.step(eps, step, xa)
            st = ours._plan_step(int(step))
            xb = _emulate(st, eps, ring, stash, xb)
            assert torch.allclose(xa, xb, rtol=2e-5, atol=2e-5), (t_start, int(step))
            assert ref.counter == ours.counter and len(ref.ets) == ours.ets_len


def test_snr_shift_overwrite_is_honoured():
    """The reference overwrites betas/alphas/alphas_cumprod after construction (trainers/reconstruct.py:106-117)."""
    kw = dict(num_train_timesteps=1000, skip_prk_steps=True, schedule="linear_beta", beta_start=1e-4, beta_end=2e-2)
    ref, ours = op.PNDMScheduler(**kw), PNDMScheduler(**kw)
    op.snr_shift_(ref, 0.5)
    op.snr_shift_(ours, 0.5)
    ref.set_timesteps(100)
    ours.set_timesteps(100)
    assert torch.equal(ref.alphas_cumprod, ours.alphas_cumprod)
    g = torch.Generator().manual_seed(1)
    shape = (1, 1, 2, 2)
    ring, stash = torch.zeros((4,) + shape), torch.zeros(shape)
    xa = torch.randn(shape, generator=g)
    xb = xa.clone()
    for step in ref.timesteps[ref.timesteps <= 40]:
        eps = torch.randn(shape, generator=g)
        xa, _ = ref.step(eps, step, xa)
        xb = _emulate(ours._plan_step(int(step)), eps, ring, stash, xb)
    assert torch.allclose(xa, xb, rtol=2e-5, atol=2e-5)


def test_plms_float64_rederivation():
    """Independent check of the oracle itself: the PLMS transfer formula equals the DDIM-style update
    x_prev = sqrt(a_prev) * x0_hat + sqrt(1 - a_prev) * eps with x0_hat = (x - sqrt(1-a_t) eps)/sqrt(a_t), in float64."""
    s = op.PNDMScheduler(num_train_timesteps=1000, skip_prk_steps=True, schedule="scaled_linear_beta",
                         beta_start=0.0015, beta_end=0.0195)
    s.set_timesteps(100)
    ac = s.alphas_cumprod.double()
    g = torch.Generator().manual_seed(2)
    x = torch.randn(5, generator=g, dtype=torch.float64)
    e = torch.randn(5, generator=g, dtype=torch.float64)
    for t, tp in ((990, 980), (500, 490), (10, 0)):
        a_t, a_p = ac[t], ac[tp]
        x0 = (x - (1 - a_t).sqrt() * e) / a_t.sqrt()
        want = a_p.sqrt() * x0 + (1 - a_p).sqrt() * e
        got = s._get_prev_sample(x.float(), t, tp, e.float()).double()
        assert torch.allclose(got, want, rtol=1e-4, atol=1e-5)


def test_final_step_is_identity():
    """At t=0 prev_t<0 -> alpha_prev = final_alpha_cumprod = alphas_cumprod[0]: x is unchanged."""
    s = PNDMScheduler(num_train_timesteps=1000, skip_prk_steps=True)
    s.set_timesteps(100)
    s.counter = 5
    s._hist = [0, 1, 2]
    st = s._plan_step(0)
    assert st.A == pytest.approx(1.0) and st.Bc == pytest.approx(0.0)
