"""The reconstruction hot loop (src/trainers/reconstruct.py:96-204) on the B200 engine.

`BatchReconstructor.score_batch` is what one iteration of the reference's `for batch in loader:` body computes:
for every t-start of the grid, forward-noise the batch, run the PLMS reverse chain of UNet evaluations (ONE engine call
per t-start: UNet forward + fused scheduler update per step), un-scale/clamp, LPIPS and MSE. All per-(image, t-start)
scores stay on the device until the batch is done (the reference syncs with `.item()` 2·B times per t-start).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

from . import _lib
from .schedulers import PNDMScheduler


def snr_shift_(scheduler, snr_shift: float) -> None:
    """src/trainers/base.py:104-116 and src/trainers/reconstruct.py:106-117 (identical blocks)."""
    if snr_shift == 1:
        return
    snr = scheduler.alphas_cumprod / (1 - scheduler.alphas_cumprod)
    target_snr = snr * snr_shift
    new_alphas_cumprod = 1 / (torch.pow(target_snr, -1) + 1)
    new_alphas = torch.zeros_like(new_alphas_cumprod)
    new_alphas[0] = new_alphas_cumprod[0]
    for i in range(1, len(new_alphas)):
        new_alphas[i] = new_alphas_cumprod[i] / new_alphas_cumprod[i - 1]
    scheduler.betas = 1 - new_alphas
    scheduler.alphas = new_alphas
    scheduler.alphas_cumprod = new_alphas_cumprod


def clamp_mse(x: torch.Tensor, x0: torch.Tensor, b_scale: float):
    """recon = clamp(x / b_scale, 0, 1); mse[b] = mean((x0 - recon)^2) — trainers/reconstruct.py:167-168,188-191."""
    n = x.shape[0]
    recon = torch.empty_like(x)
    mse = torch.empty((n,), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().ddpm_clamp_mse(x.data_ptr(), x0.data_ptr(), float(b_scale), recon.data_ptr(),
                                             mse.data_ptr(), n, x.numel() // max(n, 1),
                                             torch.cuda.current_stream().cuda_stream), "ddpm_clamp_mse")
    return recon, mse


def partition_t_starts(chain_lens: Sequence[int], world: int) -> List[List[int]]:
    """Split the t-start grid over `world` ranks, balancing UNet evaluations (a chain from t costs ~t/10 forwards, so a
    round-robin split leaves the last rank with up to twice the work of the first): longest chain first onto the least
    loaded rank, ties to the lowest rank - deterministic, every rank computes the same table. Returns the grid
    positions of each rank, ascending."""
    order = sorted(range(len(chain_lens)), key=lambda i: (-int(chain_lens[i]), i))
    load = [0] * world
    parts: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        parts[r].append(i)
        load[r] += int(chain_lens[i])
    # local search: move or swap chains between the most loaded rank and any other while that lowers the larger of the two
    for _ in range(64 * world):
        hi = max(range(world), key=lambda k: (load[k], -k))
        best, best_max = None, load[hi]
        for lo in range(world):
            if lo == hi:
                continue
            for a in parts[hi]:
                la = int(chain_lens[a])
                for b in [None] + parts[lo]:
                    lb = int(chain_lens[b]) if b is not None else 0
                    m = max(load[hi] - la + lb, load[lo] + la - lb)
                    if m < best_max:
                        best, best_max = (lo, a, b), m
        if best is None:
            break
        lo, a, b = best
        parts[hi].remove(a)
        parts[lo].append(a)
        load[hi] -= int(chain_lens[a])
        load[lo] += int(chain_lens[a])
        if b is not None:
            parts[lo].remove(b)
            parts[hi].append(b)
            load[lo] -= int(chain_lens[b])
            load[hi] += int(chain_lens[b])
    return [sorted(p) for p in parts]


@dataclass
class ReconConfig:
    prediction_type: str = "epsilon"
    beta_schedule: str = "linear"
    beta_start: float = 1e-4
    beta_end: float = 2e-2
    b_scale: float = 1.0
    snr_shift: float = 1.0
    spatial_dimension: int = 2
    num_inference_steps: int = 100  # the reference hard-codes 100 (trainers/reconstruct.py:118)
    plms_state: str = "carry"       # "carry": scheduler state survives across t-starts (reference-faithful); "reset"


class BatchReconstructor:
    def __init__(self, model, perceptual, cfg: ReconConfig, device, vqvae_model=None, latent_pad=None):
        self.model = model
        self.pl = perceptual
        self.cfg = cfg
        self.device = torch.device(device)
        self.vqvae_model = vqvae_model
        self.latent_pad = latent_pad
        self._sched = None
        if cfg.plms_state not in ("carry", "reset"):
            raise ValueError("plms_state must be 'carry' or 'reset'")

    def make_scheduler(self) -> PNDMScheduler:
        """The reference builds a fresh PNDMScheduler per batch (trainers/reconstruct.py:98-118); here ONE object is
        kept and `set_timesteps` puts it back into the freshly-constructed state (empty PLMS history, counter 0), so
        its cached per-chain coefficient tables survive from batch to batch."""
        c = self.cfg
        if self._sched is None:
            s = PNDMScheduler(num_train_timesteps=1000, skip_prk_steps=True, prediction_type=c.prediction_type,
                              schedule=c.beta_schedule, beta_start=c.beta_start, beta_end=c.beta_end)
            snr_shift_(s, c.snr_shift)
            self._sched = s
        self._sched.set_timesteps(c.num_inference_steps)
        return self._sched

    @torch.no_grad()
    def score_batch(self, images_original: torch.Tensor, inference_skip_factor: int,
                    noise_fn: Optional[Callable[[int, int], torch.Tensor]] = None,
                    keep_recons: bool = False, t_starts: Optional[Sequence[int]] = None,
                    t_indices: Optional[Sequence[int]] = None) -> Dict[str, object]:
        """images_original: [B, C, ...] in [0, 1], host or device. Returns device tensors:
        {"t": int64 [n_t] (host), "perceptual_difference": fp32 [n_t, B], "mse": fp32 [n_t, B]}.

        t_starts: run these grid values instead of the whole grid (bounded parity samples).
        t_indices: run only these positions of the grid - this rank's share when the grid is sharded over ranks
        (`partition_t_starts`; exact only with plms_state="reset"). `noise_fn(i, t)` always receives the
        position i in the FULL grid, so every rank draws the noise the single-rank run would."""
        c = self.cfg
        sched = self.make_scheduler()
        timesteps = sched.timesteps
        starts = reversed(timesteps)[1::inference_skip_factor]  # the t-start grid, trainers/reconstruct.py:119-120
        grid_pos = list(range(len(starts)))
        if t_starts is not None:
            starts = torch.tensor([int(t) for t in t_starts], dtype=torch.long)
            grid_pos = list(range(len(starts)))
        if t_indices is not None:
            if c.plms_state != "reset":
                raise ValueError("sharding the t-start grid needs plms_state='reset': in 'carry' mode the PLMS history "
                                 "couples every chain to its predecessor")
            grid_pos = [int(i) for i in t_indices]
            starts = starts[torch.tensor(grid_pos, dtype=torch.long)] if grid_pos else starts[:0]
        images_original = images_original.to(self.device, non_blocking=True).float().contiguous()
        if images_original.shape[0] == 0:  # an empty batch scores nothing (the reference's loop would emit no rows)
            empty = torch.empty((len(starts), 0), dtype=torch.float32, device=self.device)
            return {"t": starts.clone(), "t_index": grid_pos, "perceptual_difference": empty, "mse": empty.clone()}
        images = images_original if self.vqvae_model is None else self.vqvae_model.encode_stage_2_inputs(images_original)
        if self.latent_pad:
            images = F.pad(input=images, pad=self.latent_pad, mode="constant", value=0)
        images = images.contiguous()
        B = images.shape[0]
        n_t = len(starts)
        pd_all = torch.empty((n_t, B), dtype=torch.float32, device=self.device)
        mse_all = torch.empty((n_t, B), dtype=torch.float32, device=self.device)
        recons = []
        scaled = images * c.b_scale if c.b_scale != 1 else images
        for i, t_start in enumerate(starts):
            if c.plms_state == "reset":
                sched.reset_chain()
            start_timesteps = torch.Tensor([t_start] * B).long()
            if noise_fn is None:
                noise = torch.randn_like(images)
            elif getattr(noise_fn, "wants_like", False):  # noise shaped like what is noised (the latent for an LDM)
                noise = noise_fn(grid_pos[i], int(t_start), images)
            else:
                noise = noise_fn(grid_pos[i], int(t_start))
            x = sched.add_noise(original_samples=scaled, noise=noise, timesteps=start_timesteps)
            chain = [int(s) for s in timesteps[timesteps <= t_start]]
            sched.run_chain(self.model, x, chain)
            if self.latent_pad:
                x = F.pad(input=x, pad=[-p for p in self.latent_pad], mode="constant", value=0).contiguous()
            if self.vqvae_model is not None:
                x = self.vqvae_model.decode_stage_2_outputs(x).float().contiguous()
            recon, mse = clamp_mse(x, images_original, c.b_scale)
            if self.pl is None:  # latent-only runs (a 128-channel latent is not an LPIPS input): MSE only
                pd_all[i] = float("nan")
            elif c.spatial_dimension == 2:
                if images_original.shape[3] == 28:
                    pd = self.pl(F.pad(images_original, (2, 2, 2, 2)), F.pad(recon, (2, 2, 2, 2)))
                else:
                    pd = self.pl(images_original, recon)
                pd_all[i] = pd.reshape(B)
            elif hasattr(self.pl, "per_item"):
                # one score per item, as the reference's 3-D loop produces (trainers/reconstruct.py:181-187), with the
                # slices of several items per LPIPS call
                pd_all[i] = self.pl.per_item(images_original, recon).reshape(B)
            else:
                for b in range(B):  # per item, as the reference does in 3-D (trainers/reconstruct.py:181-187)
                    pd_all[i, b] = self.pl(images_original[b, None, ...], recon[b, None, ...])
            mse_all[i] = mse
            if keep_recons:
                recons.append(recon)
        out: Dict[str, object] = {"t": starts.clone(), "t_index": grid_pos, "perceptual_difference": pd_all,
                                  "mse": mse_all}
        if keep_recons:
            out["recons"] = recons
        return out
