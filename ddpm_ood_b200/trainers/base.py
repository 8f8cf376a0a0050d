"""`BaseTrainer`: what the reconstruction path needs from the reference's trainer base class (src/trainers/base.py:18-164) -
one process per GPU under torchrun, the stage-1 model (pass-through or VQ-VAE), the diffusion UNet, the scheduler with the
optional SNR shift, and the checkpoint. Same attribute names as the reference so `Reconstruct` reads like its counterpart;
training-only state (optimizer, GradScaler, DDP wrapper, inferer) is out of scope."""
from __future__ import annotations

import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

from ..networks import DiffusionModelUNet, PassthroughVQVAE
from ..reconstruction import snr_shift_
from ..schedulers import DDPMScheduler
from ..simplex_noise import Simplex_CLASS
from ..vqvae import VQVAE

# --model_type -> DiffusionModelUNet geometry (src/trainers/base.py:66-86)
UNET_TYPES = {
    "small": dict(num_channels=(128, 256, 256), attention_levels=(False, False, True), num_res_blocks=1),
    "big": dict(num_channels=(256, 512, 768), attention_levels=(True, True, True), num_res_blocks=2),
}


class BaseTrainer:
    def __init__(self, args):
        self._init_device()
        print(f"Arguments: {args}")
        for name, value in vars(args).items():
            print(f"  {name}: {value}")

        channels = self._init_stage1(args)
        if args.model_type not in UNET_TYPES:
            raise ValueError(f"Do not recognise model type {args.model_type}")
        self.model = DiffusionModelUNet(spatial_dims=args.spatial_dimension, in_channels=channels, out_channels=channels,
                                        num_head_channels=256, with_conditioning=False,
                                        **UNET_TYPES[args.model_type]).to(self.device)
        print(f"{sum(p.numel() for p in self.model.parameters()):,} model parameters")

        for name in ("prediction_type", "beta_schedule", "beta_start", "beta_end", "b_scale", "snr_shift",
                     "spatial_dimension"):
            setattr(self, name, getattr(args, name))
        self.scheduler = DDPMScheduler(num_train_timesteps=1000, prediction_type=self.prediction_type,
                                       schedule=self.beta_schedule, beta_start=self.beta_start, beta_end=self.beta_end)
        if self.snr_shift != 1:  # base.py:106-116: rescale the cumulative alphas, rebuild alphas / betas from them
            print("Changing scheduler parameters to shift SNR")
            snr_shift_(self.scheduler, self.snr_shift)

        self.simplex_noise = bool(args.simplex_noise)
        if self.simplex_noise:
            if args.spatial_dimension != 2:
                raise NotImplementedError("simplex noise is defined for 2-D images (src/utils/simplex_noise.py:15-79)")
            self.simplex = Simplex_CLASS()
        self.image_size = int(args.image_size) if args.image_size else args.image_size
        self.do_latent_pad = bool(args.latent_pad)
        self.latent_pad = args.latent_pad if self.do_latent_pad else None
        if self.do_latent_pad:
            self.inverse_latent_pad = [-p for p in self.latent_pad]

        self.run_dir = Path(args.output_dir) / args.model_name
        self._load_checkpoint(args)

    # ------------------------------------------------------------------------------------------------ pieces
    def _init_device(self) -> None:
        """torchrun launches one process per GPU (LOCAL_RANK set): NCCL process group, output of the other ranks muted
        (base.py:21-37). Without torchrun: the first GPU. There is no CPU path."""
        self.ddp = "LOCAL_RANK" in os.environ
        if self.ddp:
            print("Setting up DDP.")
            local_rank = int(os.environ["LOCAL_RANK"])
            if local_rank != 0:
                sys.stdout = sys.stderr = open(os.devnull, "w")
            if not dist.is_initialized():
                dist.init_process_group(backend="nccl", init_method="env://")
            self.device = torch.device(f"cuda:{local_rank}")
        else:
            if not torch.cuda.is_available():
                raise RuntimeError("ddpm_ood_b200 needs a CUDA device (sm_100a); there is no CPU path")
            self.device = torch.device("cuda:0")
        torch.cuda.set_device(self.device)

    def _init_stage1(self, args) -> int:
        """Pixel-space models score images directly (PassthroughVQVAE); latent models load the VQ-VAE whose constructor kwargs
        sit in `vqvae_config.json` next to `--vqvae_checkpoint` (base.py:44-61). Returns the UNet's channel count."""
        if not args.vqvae_checkpoint:
            self.vqvae_model = PassthroughVQVAE()
            return 1 if args.is_grayscale else 3
        ckpt = Path(args.vqvae_checkpoint)
        cfg_file = ckpt.parent / "vqvae_config.json"
        for what, path in (("checkpoint", ckpt), ("config", cfg_file)):
            if not path.exists():
                raise FileNotFoundError(f"Cannot find VQ-VAE {what} {path}")
        self.vqvae_config = json.loads(cfg_file.read_text())
        self.vqvae_model = VQVAE(**self.vqvae_config)
        state = torch.load(ckpt, map_location="cpu", weights_only=False)
        self.vqvae_model.load_state_dict(state["model_state_dict"])
        self.vqvae_model.to(self.device).eval()
        print("Loaded vqvae model with config:")
        for name, value in self.vqvae_config.items():
            print(f"  {name}: {value}")
        return self.vqvae_config["embedding_dim"]

    def _load_checkpoint(self, args) -> None:
        """`checkpoint.pth` (or `checkpoint_<epoch>.pth`) of the run directory, if present (base.py:137-164)."""
        epoch = args.ddpm_checkpoint_epoch
        path = self.run_dir / (f"checkpoint_{int(epoch)}.pth" if epoch else "checkpoint.pth")
        self.found_checkpoint = path.exists()
        self.start_epoch, self.global_step, self.best_loss = 0, 0, 1000
        if self.found_checkpoint:
            state = torch.load(path, map_location=self.device, weights_only=False)
            self.model.load_state_dict(state["model_state_dict"])
            self.start_epoch = state["epoch"] + 1
            self.global_step = state["global_step"]
            self.best_loss = state["best_loss"]
            print(f"Resuming training using checkpoint {path} at epoch {self.start_epoch}")
