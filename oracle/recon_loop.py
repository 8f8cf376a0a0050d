"""ORACLE (test infrastructure, not product code): the reconstruction loop of the reference,
src/trainers/reconstruct.py:96-204, restated with injected noise (the reference draws `torch.randn_like` per t-start and
never seeds it) and without the data loader / plotting.

THE LOOP IS PINNED: tests/test_reference_loop_pin_cpu.py loads the reference's own `Reconstruct.get_scores` from
/root/reference (absent third-party imports replaced by the oracle's classes) and requires its CSV rows to equal this
function's output bit for bit on the same model, images and noise (32 x 32 and 28 x 28, SNR shift, b_scale).
PARITY UNPINNED for what the loop calls: the UNet, scheduler and LPIPS arithmetic are restatements (oracle/unet.py,
oracle/pndm.py, oracle/lpips.py) of packages that are not available here.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional

import torch
import torch.nn.functional as F

from .pndm import PNDMScheduler, snr_shift_, t_start_grid


@dataclass
class LoopConfig:
    prediction_type: str = "epsilon"
    beta_schedule: str = "scaled_linear_beta"
    beta_start: float = 0.0015
    beta_end: float = 0.0195
    b_scale: float = 1.0
    snr_shift: float = 1.0
    inference_skip_factor: int = 1
    num_inference_steps: int = 100  # hard-coded 100 in the reference (trainers/reconstruct.py:118)
    spatial_dimension: int = 2
    plms_state: str = "carry"  # "carry" = reference-faithful (state survives across t-starts), "reset" = per chain


def make_scheduler(cfg: LoopConfig) -> PNDMScheduler:
    s = PNDMScheduler(num_train_timesteps=1000, skip_prk_steps=True, prediction_type=cfg.prediction_type,
                      schedule=cfg.beta_schedule, beta_start=cfg.beta_start, beta_end=cfg.beta_end)
    snr_shift_(s, cfg.snr_shift)
    s.set_timesteps(cfg.num_inference_steps)
    return s


def reconstruct_batch(
    model: Callable[[torch.Tensor, torch.Tensor], torch.Tensor],
    perceptual: Callable[[torch.Tensor, torch.Tensor], torch.Tensor],
    images_original: torch.Tensor,
    noise_fn: Callable[[int, int], torch.Tensor],
    cfg: LoopConfig,
    keep_recons: bool = False,
    t_starts: Optional[List[int]] = None,
    vqvae=None,
    latent_pad: Optional[List[int]] = None,
    keep_indices: bool = False,
) -> Dict[str, object]:
    """One batch of trainers/reconstruct.py:97-204. vqvae=None is the PassthroughVQVAE of pixel-space models; otherwise
    `vqvae.encode_stage_2_inputs` before the t-start loop (:124), optional constant latent padding (:125-126, undone at
    :159-165) and `vqvae.decode_stage_2_outputs` per t-start (:166). keep_indices (vqvae only): also return the codebook
    rows of the encoding and of every reconstructed latent ("enc_indices", "dec_indices").

    noise_fn(i, t_start) -> noise tensor like images for the i-th t-start.
    Returns {"t": LongTensor[n_t], "perceptual_difference": [n_t, B], "mse": [n_t, B], ("recons": list)}.
    """
    sched = make_scheduler(cfg)
    timesteps = sched.timesteps
    starts = t_start_grid(timesteps, cfg.inference_skip_factor)
    if t_starts is not None:  # bounded samples of the grid (bench.py's CPU arm); every entry must be a grid value
        starts = torch.tensor([int(t) for t in t_starts], dtype=torch.long)
    with torch.no_grad():
        images = images_original if vqvae is None else vqvae.encode_stage_2_inputs(images_original)
        enc_indices = vqvae.index_quantize(images_original) if (vqvae is not None and keep_indices) else None
    if latent_pad:
        images = F.pad(input=images, pad=latent_pad, mode="constant", value=0)
    dec_indices: List[torch.Tensor] = []
    B = images.shape[0]
    out_t: List[int] = []
    out_p: List[torch.Tensor] = []
    out_m: List[torch.Tensor] = []
    recons: List[torch.Tensor] = []
    with torch.no_grad():
        for i, t_start in enumerate(starts):
            if cfg.plms_state == "reset":
                sched.reset_chain()
            start_timesteps = torch.Tensor([t_start] * B).long()
            noise = noise_fn(i, int(t_start))
            x = sched.add_noise(original_samples=images * cfg.b_scale, noise=noise, timesteps=start_timesteps)
            for step in timesteps[timesteps <= t_start]:
                ts = torch.Tensor([step] * B).long()
                eps = model(x, ts)
                x, _ = sched.step(eps, step, x)
            if latent_pad:
                x = F.pad(input=x, pad=[-p for p in latent_pad], mode="constant", value=0)
            if vqvae is not None:
                if keep_indices:
                    dec_indices.append(vqvae.quantizer.quantize(x))
                x = vqvae.decode_stage_2_outputs(x)
            x = x / cfg.b_scale
            x = x.clamp(0, 1)
            if perceptual is None:  # latent-only checks (a 128-channel latent is not an LPIPS input): MSE only
                pd = torch.full((B,), float("nan"))
            elif cfg.spatial_dimension == 2:
                if images_original.shape[3] == 28:
                    pd = perceptual(F.pad(images_original, (2, 2, 2, 2)), F.pad(x, (2, 2, 2, 2)))
                else:
                    pd = perceptual(images_original, x)
                pd = pd.reshape(B)
            else:
                pd = torch.empty(B)
                for b in range(B):
                    pd[b] = perceptual(images_original[b, None, ...], x[b, None, ...])
            non_batch = tuple(range(images_original.dim()))[1:]
            mse = torch.square(images_original - x).mean(dim=non_batch)
            out_t.append(int(t_start))
            out_p.append(pd.float())
            out_m.append(mse.float())
            if keep_recons:
                recons.append(x.clone())
    res: Dict[str, object] = {
        "t": torch.tensor(out_t, dtype=torch.long),
        "perceptual_difference": torch.stack(out_p),
        "mse": torch.stack(out_m),
    }
    if keep_recons:
        res["recons"] = recons
    if enc_indices is not None:
        res["enc_indices"] = enc_indices
        res["dec_indices"] = torch.stack(dec_indices)
    return res


def unet_evals_per_batch(skip_factor: int, num_inference_steps: int = 100) -> int:
    s = PNDMScheduler(num_train_timesteps=1000, skip_prk_steps=True)
    s.set_timesteps(num_inference_steps)
    ts = s.timesteps
    return int(sum(int((ts <= t).sum()) for t in t_start_grid(ts, skip_factor)))
