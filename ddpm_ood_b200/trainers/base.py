"""`BaseTrainer`, reconstruction-path subset of src/trainers/base.py:18-164: process-group init, model and scheduler
construction, SNR shift, checkpoint loading. Training-only state (optimizer, GradScaler, DDP wrapper, inferer) is out of
scope."""
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


class BaseTrainer:
    def __init__(self, args):
        # initialise the process group if launched with torchrun (base.py:21-37): one process per GPU
        if "LOCAL_RANK" in os.environ:
            print("Setting up DDP.")
            self.ddp = True
            local_rank = int(os.environ["LOCAL_RANK"])
            if local_rank != 0:
                f = open(os.devnull, "w")
                sys.stdout = sys.stderr = f
            if not dist.is_initialized():
                dist.init_process_group(backend="nccl", init_method="env://")
            self.device = torch.device(f"cuda:{local_rank}")
        else:
            self.ddp = False
            if not torch.cuda.is_available():
                raise RuntimeError("ddpm_ood_b200 needs a CUDA device (sm_100a); there is no CPU path")
            self.device = torch.device("cuda:0")
        torch.cuda.set_device(self.device)

        print(f"Arguments: {str(args)}")
        for k, v in vars(args).items():
            print(f"  {k}: {v}")

        if args.vqvae_checkpoint:  # latent diffusion: stage-1 VQ-VAE built from the json next to its checkpoint (base.py:44-61)
            vqvae_checkpoint_path = Path(args.vqvae_checkpoint)
            vqvae_config_path = vqvae_checkpoint_path.parent / "vqvae_config.json"
            if not vqvae_checkpoint_path.exists():
                raise FileNotFoundError(f"Cannot find VQ-VAE checkpoint {vqvae_checkpoint_path}")
            if not vqvae_config_path.exists():
                raise FileNotFoundError(f"Cannot find VQ-VAE config {vqvae_config_path}")
            with open(vqvae_config_path, "r") as f:
                self.vqvae_config = json.load(f)
            self.vqvae_model = VQVAE(**self.vqvae_config)
            vqvae_checkpoint = torch.load(vqvae_checkpoint_path, map_location="cpu", weights_only=False)
            self.vqvae_model.load_state_dict(vqvae_checkpoint["model_state_dict"])
            self.vqvae_model.to(self.device)
            self.vqvae_model.eval()
            print("Loaded vqvae model with config:")
            for k, v in self.vqvae_config.items():
                print(f"  {k}: {v}")
            ddpm_channels = self.vqvae_config["embedding_dim"]
        else:
            self.vqvae_model = PassthroughVQVAE()
            ddpm_channels = 1 if args.is_grayscale else 3
        if args.model_type == "small":
            self.model = DiffusionModelUNet(
                spatial_dims=args.spatial_dimension, in_channels=ddpm_channels, out_channels=ddpm_channels,
                num_channels=(128, 256, 256), attention_levels=(False, False, True), num_res_blocks=1,
                num_head_channels=256, with_conditioning=False).to(self.device)
        elif args.model_type == "big":
            self.model = DiffusionModelUNet(
                spatial_dims=args.spatial_dimension, in_channels=ddpm_channels, out_channels=ddpm_channels,
                num_channels=(256, 512, 768), attention_levels=(True, True, True), num_res_blocks=2,
                num_head_channels=256, with_conditioning=False).to(self.device)
        else:
            raise ValueError(f"Do not recognise model type {args.model_type}")
        print(f"{sum(p.numel() for p in self.model.parameters()):,} model parameters")
        self.prediction_type = args.prediction_type
        self.beta_schedule = args.beta_schedule
        self.beta_start = args.beta_start
        self.beta_end = args.beta_end
        self.b_scale = args.b_scale
        self.snr_shift = args.snr_shift
        self.scheduler = DDPMScheduler(num_train_timesteps=1000, prediction_type=self.prediction_type,
                                       schedule=self.beta_schedule, beta_start=self.beta_start, beta_end=self.beta_end)
        if self.snr_shift != 1:
            print("Changing scheduler parameters to shift SNR")
            snr_shift_(self.scheduler, self.snr_shift)
        self.simplex_noise = bool(args.simplex_noise)
        if self.simplex_noise:
            if args.spatial_dimension != 2:
                raise NotImplementedError("simplex noise is defined for 2-D images (src/utils/simplex_noise.py:15-79)")
            self.simplex = Simplex_CLASS()
        self.spatial_dimension = args.spatial_dimension
        self.image_size = int(args.image_size) if args.image_size else args.image_size
        if args.latent_pad:
            self.do_latent_pad = True
            self.latent_pad = args.latent_pad
            self.inverse_latent_pad = [-x for x in self.latent_pad]
        else:
            self.do_latent_pad = False
            self.latent_pad = None

        self.run_dir = Path(args.output_dir) / args.model_name
        if args.ddpm_checkpoint_epoch:
            checkpoint_path = self.run_dir / f"checkpoint_{int(args.ddpm_checkpoint_epoch)}.pth"
        else:
            checkpoint_path = self.run_dir / "checkpoint.pth"
        if checkpoint_path.exists():
            checkpoint = torch.load(checkpoint_path, map_location=self.device, weights_only=False)
            self.found_checkpoint = True
            self.start_epoch = checkpoint["epoch"] + 1
            self.global_step = checkpoint["global_step"]
            self.model.load_state_dict(checkpoint["model_state_dict"])
            self.best_loss = checkpoint["best_loss"]
            print(f"Resuming training using checkpoint {checkpoint_path} at epoch {self.start_epoch}")
        else:
            self.start_epoch = 0
            self.best_loss = 1000
            self.global_step = 0
            self.found_checkpoint = False
