"""Drop-in `PNDMScheduler` / `DDPMScheduler` for the reconstruction hot path.

Mirrors the monai-generative operator surface the reference touches (src/trainers/reconstruct.py:98-120,143-157;
src/trainers/base.py:97-116): constructor kwargs, mutable `.betas/.alphas/.alphas_cumprod`, `.set_timesteps`,
`.timesteps`, `.add_noise`, `.step -> (prev_sample, None)`.

Host side (this file): timestep grid, PLMS bookkeeping (`counter`, which ring slot holds which past epsilon) and the
per-step scalar coefficients. Device side (libddpm_ood_b200.so): the elementwise tensor arithmetic, either stand-alone
(`step`, `add_noise`) or fused behind the UNet's output conv (`run_chain`). No CPU fallback for tensor math.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib


def make_betas(schedule: str, num_train_timesteps: int, beta_start: float = 1e-4, beta_end: float = 2e-2) -> torch.Tensor:
    # the reference spells the names four ways (reconstruct.py:55-58, train_ddpm.py:49-52, README.md:69,
    # README_additional.md:13,26): accept all.
    if schedule in ("linear_beta", "linear"):
        return torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
    if schedule in ("scaled_linear_beta", "scaled_linear"):
        return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    raise ValueError(f"Beta schedule {schedule} is not implemented")


class Scheduler:
    def __init__(self, num_train_timesteps: int = 1000, schedule: str = "linear_beta", **schedule_args) -> None:
        self.num_train_timesteps = num_train_timesteps
        self.betas = make_betas(schedule, num_train_timesteps, **schedule_args)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.one = torch.tensor(1.0)
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].astype(np.int64))
        self._ac_dev: Optional[torch.Tensor] = None
        self._ac_key = None

    def _alphas_cumprod_on(self, device) -> torch.Tensor:
        ac = self.alphas_cumprod
        key = (ac.data_ptr(), ac._version, str(device))
        if self._ac_dev is None or self._ac_key != key:
            self._ac_dev = ac.detach().to(device=device, dtype=torch.float32).contiguous()
            self._ac_key = key
        return self._ac_dev

    def add_noise(self, original_samples: torch.Tensor, noise: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
        if not original_samples.is_cuda:
            raise _lib.DdpmError("add_noise needs CUDA tensors; there is no CPU fallback")
        dev = original_samples.device
        x0 = original_samples.detach().float().contiguous()
        nz = noise.detach().to(dev).float().contiguous()
        ts = timesteps.to(device=dev, dtype=torch.int64).contiguous()
        n = x0.shape[0]
        if ts.numel() != n:
            raise ValueError("timesteps must have one entry per batch item")
        out = torch.empty_like(x0)
        with torch.cuda.device(dev):
            _lib.check(
                _lib.lib().ddpm_add_noise(x0.data_ptr(), nz.data_ptr(), self._alphas_cumprod_on(dev).data_ptr(),
                                          ts.data_ptr(), 0, 1.0, out.data_ptr(), n, x0.numel() // max(n, 1),
                                          torch.cuda.current_stream().cuda_stream),
                "ddpm_add_noise")
        return out


class DDPMScheduler(Scheduler):
    """Only the schedule attributes are on the reconstruction path (src/trainers/base.py:97-116)."""

    def __init__(self, num_train_timesteps: int = 1000, schedule: str = "linear_beta", variance_type: str = "fixed_small",
                 clip_sample: bool = True, prediction_type: str = "epsilon", **schedule_args) -> None:
        super().__init__(num_train_timesteps, schedule, **schedule_args)
        self.prediction_type = prediction_type
        self.variance_type = variance_type
        self.clip_sample = clip_sample


class PNDMScheduler(Scheduler):
    def __init__(self, num_train_timesteps: int = 1000, schedule: str = "linear_beta", skip_prk_steps: bool = False,
                 set_alpha_to_one: bool = False, prediction_type: str = "epsilon", steps_offset: int = 0,
                 **schedule_args) -> None:
        super().__init__(num_train_timesteps, schedule, **schedule_args)
        if not skip_prk_steps:
            raise NotImplementedError("Runge-Kutta warm-up steps are not on the reference path (skip_prk_steps=True, "
                                      "src/trainers/reconstruct.py:100)")
        if prediction_type not in ("epsilon", "v_prediction"):
            raise ValueError("Argument `prediction_type` must be a member of PNDMPredictionType")
        self.prediction_type = prediction_type
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.pndm_order = 4
        self.skip_prk_steps = skip_prk_steps
        self.steps_offset = steps_offset
        self.cur_model_output = 0
        # PLMS state: `counter`, the ring slots holding past eps (oldest..newest), device buffers
        self.counter = 0
        self._hist: List[int] = []
        self._ring: Optional[torch.Tensor] = None
        self._stash: Optional[torch.Tensor] = None
        self._coeff_key = None
        self._chain_cache = {}
        self.set_timesteps(num_train_timesteps)

    # ------------------------------------------------------------------ reference surface
    def set_timesteps(self, num_inference_steps: int, device=None) -> None:
        self.num_inference_steps = num_inference_steps
        step_ratio = self.num_train_timesteps // self.num_inference_steps
        ts = (np.arange(0, num_inference_steps) * step_ratio).round().astype(np.int64) + self.steps_offset
        self._timesteps = ts
        self.prk_timesteps = np.array([])
        self.plms_timesteps = np.concatenate([ts[:-1], ts[-2:-1], ts[-1:]])[::-1].copy()
        timesteps = self.plms_timesteps.astype(np.int64)
        self.timesteps = torch.from_numpy(timesteps).to(device) if device is not None else torch.from_numpy(timesteps)
        self.ets: List[int] = []  # kept for API familiarity; the tensors live in the device ring
        self.reset_chain()

    def reset_chain(self) -> None:
        """Forget the PLMS history (what set_timesteps does). Calling this per t-start gives the "reset" mode."""
        self.counter = 0
        self._hist = []

    @property
    def ets_len(self) -> int:
        return len(self._hist)

    def _buffers(self, like: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        numel = like.numel()
        if self._ring is None or self._ring.device != like.device or self._ring.shape[1] != numel:
            if self._hist or self.counter:
                if self._ring is not None:
                    raise _lib.DdpmError("sample shape changed while the PLMS history is non-empty")
            self._ring = torch.zeros((4, numel), dtype=torch.float32, device=like.device)
            self._stash = torch.zeros((numel,), dtype=torch.float32, device=like.device)
        return self._ring, self._stash

    def _coeff_table(self):
        """Host copy of alphas_cumprod (fp32, as the reference's torch scalars) and final_alpha_cumprod, refreshed when
        the caller overwrites the schedule attributes (SNR shift, src/trainers/reconstruct.py:106-117)."""
        ac = self.alphas_cumprod
        fa = self.final_alpha_cumprod
        key = (ac.data_ptr(), ac._version, fa.data_ptr(), fa._version)
        if self._coeff_key != key:
            self._ac_host = ac.detach().float().cpu().numpy().astype(np.float32)
            self._fa_host = np.float32(float(fa.detach().float().cpu()))
            self._coeff_key = key
            self._chain_cache.clear()
        return self._ac_host, self._fa_host

    def _plan_step(self, timestep: int) -> _lib.PlmsStep:
        """Advance the host-side PLMS state by one step and return the device coefficients (step_plms + _get_prev_sample)."""
        ratio = self.num_train_timesteps // self.num_inference_steps
        prev_timestep = timestep - ratio
        st = _lib.PlmsStep()
        if self.counter != 1:
            before = self._hist[-3:]
            free = [s for s in range(4) if s not in before]
            st.push = 1
            st.slot_new = free[0]
            n_after = len(before) + 1
        else:
            before = list(self._hist)
            st.push = 0
            st.slot_new = 0
            n_after = len(before)
            prev_timestep = timestep
            timestep = timestep + ratio
        if n_after == 1 and self.counter == 0:
            c = (1.0, 0.0, 0.0, 0.0)
            st.write_stash = 1
        elif n_after == 1 and self.counter == 1:
            c = (0.5, 0.5, 0.0, 0.0)
            st.use_stash = 1
        elif n_after == 2:
            c = (3.0 / 2.0, -1.0 / 2.0, 0.0, 0.0)
        elif n_after == 3:
            c = (23.0 / 12.0, -16.0 / 12.0, 5.0 / 12.0, 0.0)
        else:
            c = (55.0 / 24.0, -59.0 / 24.0, 37.0 / 24.0, -9.0 / 24.0)
        for i in range(4):
            st.c[i] = c[i]
        newest_first = list(reversed(before))
        for i in range(3):
            st.slot[i] = newest_first[i] if i < len(newest_first) else 0
        # _get_prev_sample in fp32 scalars, like the reference's 0-d torch tensors (numpy float32 arithmetic, no
        # per-step tensor traffic)
        ac, fa = self._coeff_table()
        a_t = ac[timestep]
        a_prev = ac[prev_timestep] if prev_timestep >= 0 else fa
        one = np.float32(1.0)
        b_t = one - a_t
        b_prev = one - a_prev
        if self.prediction_type == "v_prediction":
            st.vA = float(np.sqrt(a_t))
            st.vB = float(np.sqrt(b_t))
        else:
            st.vA, st.vB = 1.0, 0.0
        st.A = float(np.sqrt(a_prev / a_t))
        denom = a_t * np.sqrt(b_prev) + np.sqrt(a_t * b_t * a_prev)
        st.Bc = float((a_prev - a_t) / denom)
        # state advance
        if st.push:
            self._hist = before + [st.slot_new]
        self.counter += 1
        return st

    def plan_chain(self, timesteps: Sequence[int]):
        """Coefficients of a whole chain (ctypes array of PlmsStep), advancing the host state. The result depends only
        on (timesteps, counter in {0, 1, >=2}, ring occupancy, schedule): cached, so the 25-100 chains of every batch
        after the first cost one dictionary lookup instead of per-step Python."""
        self._coeff_table()
        ts = tuple(int(t) for t in timesteps)
        key = (ts, min(self.counter, 2), tuple(self._hist), self.num_inference_steps, self.prediction_type)
        hit = self._chain_cache.get(key)
        if hit is not None:
            steps, hist_after = hit
            self.counter += len(ts)
            self._hist = list(hist_after)
            return steps
        steps = (_lib.PlmsStep * max(len(ts), 1))()
        for i, t in enumerate(ts):
            steps[i] = self._plan_step(t)
        if len(self._chain_cache) < 4096:
            self._chain_cache[key] = (steps, tuple(self._hist))
        return steps

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor) -> Tuple[torch.Tensor, None]:
        if not sample.is_cuda:
            raise _lib.DdpmError("PNDMScheduler.step needs CUDA tensors; there is no CPU fallback")
        mo = model_output.detach().float().contiguous()
        sm = sample.detach().float().contiguous()
        ring, stash = self._buffers(sm)
        st = self._plan_step(int(timestep))
        out = torch.empty_like(sm)
        with torch.cuda.device(sm.device):
            _lib.check(
                _lib.lib().ddpm_plms_update(mo.data_ptr(), C.byref(st), ring.data_ptr(), stash.data_ptr(), sm.data_ptr(),
                                            out.data_ptr(), sm.numel(), torch.cuda.current_stream().cuda_stream),
                "ddpm_plms_update")
        return out, None

    # ------------------------------------------------------------------ fused path
    def run_chain(self, model, sample: torch.Tensor, timesteps: Sequence[int]) -> torch.Tensor:
        """`for step in timesteps: sample, _ = self.step(model(sample, step), step, sample)` as one engine call
        (src/trainers/reconstruct.py:149-157). `sample` (fp32, contiguous) is updated in place and returned."""
        if sample.dtype != torch.float32 or not sample.is_contiguous():
            raise ValueError("run_chain needs a contiguous fp32 sample")
        ring, stash = self._buffers(sample)
        ts = [int(t) for t in timesteps]
        steps = self.plan_chain(ts)
        model.run_chain(sample, ts, steps, ring, stash)
        return sample
