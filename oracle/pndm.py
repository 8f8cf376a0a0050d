"""ORACLE (test infrastructure, not product code): restatement of monai-generative's `PNDMScheduler` and the base
`Scheduler.add_noise`, as used by the reference at src/trainers/reconstruct.py:98-120,143-147,155-157 and of
`DDPMScheduler`'s schedule attributes at src/trainers/base.py:97-116.

PARITY UNPINNED for the arithmetic (the `generative` package is not available). One reference-side
pin exists and is tested: the skip-factor table README.md:118-120 fixes len(timesteps) == 101 for set_timesteps(100).

The scheduler keeps its PLMS state (`ets`, `counter`, `cur_sample`) between calls; the reference builds one scheduler
per batch and calls set_timesteps once (trainers/reconstruct.py:98-118), so state carries over from one t-start chain
to the next ("carry" mode). `reset_chain()` gives the per-chain-state variant ("reset" mode).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch


def make_betas(schedule: str, num_train_timesteps: int, beta_start: float, beta_end: float) -> torch.Tensor:
    """NoiseSchedules registry: 'linear_beta' and 'scaled_linear_beta' (the reference CLI also spells them 'linear',
    'scaled_linear': reconstruct.py:55-58, README_additional.md:13)."""
    if schedule in ("linear_beta", "linear"):
        return torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
    if schedule in ("scaled_linear_beta", "scaled_linear"):
        return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    raise ValueError(f"unknown beta schedule {schedule}")


def snr_shift_(sched, snr_shift: float) -> None:
    """src/trainers/base.py:104-116 / trainers/reconstruct.py:106-117, verbatim semantics."""
    if snr_shift == 1:
        return
    snr = sched.alphas_cumprod / (1 - sched.alphas_cumprod)
    target_snr = snr * snr_shift
    new_ac = 1 / (torch.pow(target_snr, -1) + 1)
    new_alphas = torch.zeros_like(new_ac)
    new_alphas[0] = new_ac[0]
    for i in range(1, len(new_alphas)):
        new_alphas[i] = new_ac[i] / new_ac[i - 1]
    sched.betas = 1 - new_alphas
    sched.alphas = new_alphas
    sched.alphas_cumprod = new_ac


class Scheduler:
    def __init__(self, num_train_timesteps=1000, schedule="linear_beta", beta_start=1e-4, beta_end=2e-2):
        self.num_train_timesteps = num_train_timesteps
        self.betas = make_betas(schedule, num_train_timesteps, beta_start, beta_end)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.one = torch.tensor(1.0)
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].astype(np.int64))

    def add_noise(self, original_samples: torch.Tensor, noise: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
        ac = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        timesteps = timesteps.to(original_samples.device)
        sa = ac[timesteps] ** 0.5
        sb = (1 - ac[timesteps]) ** 0.5
        while sa.dim() < original_samples.dim():
            sa = sa.unsqueeze(-1)
            sb = sb.unsqueeze(-1)
        return sa * original_samples + sb * noise


class DDPMScheduler(Scheduler):
    """Only the schedule attributes are on the hot path (base.py:97-116)."""

    def __init__(self, num_train_timesteps=1000, schedule="linear_beta", prediction_type="epsilon", **kw):
        super().__init__(num_train_timesteps, schedule, **kw)
        self.prediction_type = prediction_type


class PNDMScheduler(Scheduler):
    def __init__(self, num_train_timesteps=1000, schedule="linear_beta", skip_prk_steps=False, set_alpha_to_one=False,
                 prediction_type="epsilon", steps_offset=0, **schedule_args):
        super().__init__(num_train_timesteps, schedule, **schedule_args)
        if not skip_prk_steps:
            raise NotImplementedError("the reference only uses skip_prk_steps=True (PLMS)")
        self._final_from_schedule = not set_alpha_to_one
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.prediction_type = prediction_type
        self.skip_prk_steps = skip_prk_steps
        self.steps_offset = steps_offset
        self.pndm_order = 4
        self.cur_model_output = 0
        self.counter = 0
        self.cur_sample: Optional[torch.Tensor] = None
        self.ets: List[torch.Tensor] = []
        self.set_timesteps(num_train_timesteps)

    def set_timesteps(self, num_inference_steps: int, device=None) -> None:
        self.num_inference_steps = num_inference_steps
        step_ratio = self.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * step_ratio).round().astype(np.int64) + self.steps_offset
        plms = np.concatenate([ts[:-1], ts[-2:-1], ts[-1:]])[::-1].copy()
        self.timesteps = torch.from_numpy(plms.astype(np.int64))
        self.ets = []
        self.counter = 0
        self.cur_sample = None

    def reset_chain(self) -> None:
        """Per-chain PLMS state ("reset" mode); not something the reference does."""
        self.ets = []
        self.counter = 0
        self.cur_sample = None

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor) -> Tuple[torch.Tensor, None]:
        return self.step_plms(model_output, int(timestep), sample)

    def step_plms(self, model_output, timestep: int, sample):
        ratio = self.num_train_timesteps // self.num_inference_steps
        prev_timestep = timestep - ratio
        if self.counter != 1:
            self.ets = self.ets[-3:]
            self.ets.append(model_output)
        else:
            prev_timestep = timestep
            timestep = timestep + ratio
        if len(self.ets) == 1 and self.counter == 0:
            self.cur_sample = sample
        elif len(self.ets) == 1 and self.counter == 1:
            model_output = (model_output + self.ets[-1]) / 2
            sample = self.cur_sample
            self.cur_sample = None
        elif len(self.ets) == 2:
            model_output = (3 * self.ets[-1] - self.ets[-2]) / 2
        elif len(self.ets) == 3:
            model_output = (23 * self.ets[-1] - 16 * self.ets[-2] + 5 * self.ets[-3]) / 12
        else:
            model_output = (1 / 24) * (55 * self.ets[-1] - 59 * self.ets[-2] + 37 * self.ets[-3] - 9 * self.ets[-4])
        prev_sample = self._get_prev_sample(sample, timestep, prev_timestep, model_output)
        self.counter += 1
        return prev_sample, None

    def _get_prev_sample(self, sample, timestep: int, prev_timestep: int, model_output):
        # final_alpha_cumprod follows alphas_cumprod[0] only as of construction time in the third-party code; the
        # reference overwrites alphas_cumprod afterwards when snr_shift != 1 (trainers/reconstruct.py:117) and the
        # stale value then stays — restated here by NOT refreshing it.
        a_t = self.alphas_cumprod[timestep]
        a_prev = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        b_prev = 1 - a_prev
        if self.prediction_type == "v_prediction":
            model_output = (a_t ** 0.5) * model_output + (b_t ** 0.5) * sample
        sample_coeff = (a_prev / a_t) ** 0.5
        denom = a_t * b_prev ** 0.5 + (a_t * b_t * a_prev) ** 0.5
        return sample_coeff * sample - (a_prev - a_t) * model_output / denom


def t_start_grid(timesteps: torch.Tensor, inference_skip_factor: int) -> torch.Tensor:
    """src/trainers/reconstruct.py:118-120."""
    return reversed(timesteps)[1::inference_skip_factor]
