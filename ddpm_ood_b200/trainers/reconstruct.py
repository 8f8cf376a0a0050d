"""`Reconstruct` trainer: src/trainers/reconstruct.py on the B200 engine. Same constructor, `get_scores`, `reconstruct`
and CSV contract (`ood/results_{val,in,<name>}.csv`, columns filename,type,t,perceptual_difference,mse + pandas index).

Differences that do not change results: per-(image, t-start) scores stay on the device until a batch is finished (one
D2H copy per batch instead of 2·B `.item()` syncs per t-start, reference :192-204); the multi-GPU gather moves the
score tensor with one NCCL all-gather (+ a host all_gather_object for the file names) instead of pickling every row
(reference :238-242); no matplotlib figure.
"""
from __future__ import annotations

import os
import sys
import time
from pathlib import Path

import pandas as pd
import torch
import torch.distributed as dist

from ..data import get_training_data_loader
from ..losses import PerceptualLoss
from ..networks import PassthroughVQVAE
from ..reconstruction import BatchReconstructor, ReconConfig, partition_t_starts
from ..simplex_noise import generate_simplex_noise
from .base import BaseTrainer


def gather_scores(scores: torch.Tensor, names, device):
    """The one collective of the path (reference: `dist.all_gather_object(results)`, trainers/reconstruct.py:238-242):
    every rank contributes its [rows, 3] float64 (t, perceptual_difference, mse) tensor - equal row counts, because the
    image partition is even_divisible - with ONE all-gather (NCCL over NVLink on GPUs, gloo in the CPU tests), plus the
    file names as objects. Returns (all rows in rank order, all names in the same order) on every rank."""
    world = dist.get_world_size()
    local = scores.to(device).contiguous()
    gathered = torch.empty((world * local.shape[0], local.shape[1]), dtype=local.dtype, device=device)
    dist.all_gather_into_tensor(gathered, local)
    all_names = [None] * world
    dist.all_gather_object(all_names, list(names))
    return gathered.cpu(), [n for sub in all_names for n in sub]


def gather_t_sharded(full: torch.Tensor, owner: torch.Tensor, device):
    """t-start sharding (second bullet): every rank holds the same [n_t, n_images, 2] score tensor with
    only the rows of ITS t-starts filled. One all-gather of that tensor (NCCL / gloo), then row i is taken from rank
    owner[i]. Returns the complete tensor on every rank."""
    world = dist.get_world_size()
    local = full.to(device).contiguous()
    gathered = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=device)
    dist.all_gather_into_tensor(gathered, local)  # concatenated along dim 0 in rank order
    gathered = gathered.view((world,) + tuple(local.shape))
    idx = owner.to(device=device, dtype=torch.long).view(1, -1, 1, 1).expand(1, *local.shape)
    return torch.gather(gathered, 0, idx)[0]


class Reconstruct(BaseTrainer):
    def __init__(self, args):
        super().__init__(args)
        if not self.found_checkpoint:
            raise FileNotFoundError("Failed to find a saved model checkpoint.")
        self.out_dir = self.run_dir / "ood"
        self.out_dir.mkdir(exist_ok=True)
        self.args = args
        self.val_loader = self._loader(args, args.validation_ids, args.first_n_val)
        self.in_loader = self._loader(args, args.in_ids, args.first_n)
        self._pl = None

    def _loader(self, args, ids, first_n, **flip):
        rank = dist.get_rank() if dist.is_initialized() else None
        world = dist.get_world_size() if dist.is_initialized() else None
        if getattr(args, "shard", "images") == "t_starts":
            rank = world = None  # every rank sees every image and takes a share of the t-start grid instead
        return get_training_data_loader(
            batch_size=args.batch_size, training_ids=ids, validation_ids=ids, augmentation=bool(args.augmentation),
            only_val=True, num_workers=args.num_workers, num_val_workers=args.num_workers,
            cache_data=bool(args.cache_data), drop_last=bool(args.drop_last),
            first_n=int(first_n) if first_n else first_n, is_grayscale=bool(args.is_grayscale),
            spatial_dimension=args.spatial_dimension, image_size=self.image_size, image_roi=args.image_roi,
            rank=rank, world_size=world,
            # GPU-resident ingest whenever no area resize precedes the intensity scaling
            device=self.device if (not self.image_size and torch.device(self.device).type == "cuda") else None,
            **flip)

    def _engine(self) -> BatchReconstructor:
        if self._pl is None:
            self._pl = PerceptualLoss(dimensions=self.spatial_dimension, include_pixel_loss=False,
                                      is_fake_3d=True if self.spatial_dimension == 3 else False, lpips_normalize=True,
                                      spatial=False).to(self.device)
        steps = 100  # hard-coded in the reference (trainers/reconstruct.py:118); --num_inference_steps is parsed, unused
        if getattr(self.args, "honour_num_inference_steps", 0):
            steps = int(self.args.num_inference_steps)
        cfg = ReconConfig(prediction_type=self.prediction_type, beta_schedule=self.beta_schedule,
                          beta_start=self.beta_start, beta_end=self.beta_end, b_scale=self.b_scale,
                          snr_shift=self.snr_shift, spatial_dimension=self.spatial_dimension,
                          num_inference_steps=steps, plms_state=getattr(self.args, "plms_state", "carry"))
        vq = None if isinstance(self.vqvae_model, PassthroughVQVAE) else self.vqvae_model
        return BatchReconstructor(self.model, self._pl, cfg, self.device, vqvae_model=vq,
                                  latent_pad=self.latent_pad if self.do_latent_pad else None)

    def _simplex_fn(self, images):
        """--simplex_noise=1: generate_simplex_noise per t-start over whatever is noised - the image, or the (padded)
        latent of an LDM (reference trainers/reconstruct.py:133-139) - else None (Gaussian noise drawn by the engine)."""
        if not self.simplex_noise:
            return None

        def fn(i, t_start, like):
            t = torch.full((like.shape[0],), int(t_start), dtype=torch.long)
            return generate_simplex_noise(self.simplex, x=like, t=t, in_channels=like.shape[1])

        fn.wants_like = True
        return fn

    def get_scores(self, loader, dataset_name, inference_skip_factor):
        if dist.is_initialized():
            sys.stdout = sys.__stdout__
            sys.stderr = sys.__stderr__
            print(f"{dist.get_rank()}: {dataset_name}")
        else:
            print(f"{dataset_name}")
        engine = self._engine()
        self.model.eval()
        if dist.is_initialized() and getattr(self.args, "shard", "images") == "t_starts":
            return self._get_scores_t_sharded(engine, loader, dataset_name, inference_skip_factor)
        names, ts, pds, mses = [], [], [], []
        for batch in loader:
            t1 = time.time()
            res = engine.score_batch(batch["image"], inference_skip_factor, noise_fn=self._simplex_fn(batch["image"]))
            t_grid = res["t"]
            pd_host = res["perceptual_difference"].cpu()  # one D2H per batch
            mse_host = res["mse"].cpu()
            B = pd_host.shape[1]
            stems = [Path(f).stem.replace(".nii", "").replace(".gz", "")
                     for f in batch["image_meta_dict"]["filename_or_obj"]]
            for i in range(len(t_grid)):  # row order of the reference: t-start outer, batch item inner
                names.extend(stems)
                ts.append(torch.full((B,), int(t_grid[i]), dtype=torch.float64))
                pds.append(pd_host[i].double())
                mses.append(mse_host[i].double())
            t2 = time.time()
            if dist.is_initialized():
                print(f"{dist.get_rank()}: Took {t2-t1}s for a batch size of {B}")
            else:
                print(f"Took {t2-t1}s for a batch size of {B}")
        scores = torch.stack([torch.cat(ts), torch.cat(pds), torch.cat(mses)], dim=1) if ts else torch.zeros((0, 3), dtype=torch.float64)
        if dist.is_initialized():
            scores, names = gather_scores(scores, names, self.device)
            local_rank = int(os.environ["LOCAL_RANK"])
            if local_rank != 0:
                f = open(os.devnull, "w")
                sys.stdout = sys.stderr = f
        return [
            {"filename": names[r], "type": dataset_name, "t": int(scores[r, 0]),
             "perceptual_difference": float(scores[r, 1]), "mse": float(scores[r, 2])}
            for r in range(scores.shape[0])
        ]

    def _get_scores_t_sharded(self, engine, loader, dataset_name, inference_skip_factor):
        """`--shard t_starts`: rank r runs its share of the t-start grid (balanced by chain length) on EVERY image, so a
        dataset too small to fill 8 GPUs by images still does; scores meet in one all-gather of the [n_t, n, 2] tensor.
        Exact because plms_state='reset' gives every chain its own PLMS history (refused otherwise)."""
        rank, world = dist.get_rank(), dist.get_world_size()
        sched = engine.make_scheduler()
        timesteps = sched.timesteps
        grid = reversed(timesteps)[1::inference_skip_factor]
        lens = [int((timesteps <= t).sum()) for t in grid]
        parts = partition_t_starts(lens, world)
        owner = torch.empty(len(grid), dtype=torch.long)
        for r, idxs in enumerate(parts):
            owner[idxs] = r
        fulls, names, sizes = [], [], []
        for batch in loader:
            t1 = time.time()
            res = engine.score_batch(batch["image"], inference_skip_factor, noise_fn=self._simplex_fn(batch["image"]),
                                     t_indices=parts[rank])
            B = res["mse"].shape[1]
            full = torch.full((len(grid), B, 2), float("nan"), dtype=torch.float32, device=self.device)
            if parts[rank]:
                sel = torch.tensor(parts[rank], dtype=torch.long, device=self.device)
                full[sel] = torch.stack([res["perceptual_difference"], res["mse"]], dim=-1)
            fulls.append(full)
            sizes.append(B)
            names.append([Path(f).stem.replace(".nii", "").replace(".gz", "")
                          for f in batch["image_meta_dict"]["filename_or_obj"]])
            print(f"{rank}: Took {time.time()-t1}s for {len(parts[rank])} of {len(grid)} t-starts, batch size {B}")
        rows = []
        if fulls:
            scores = gather_t_sharded(torch.cat(fulls, dim=1), owner, self.device).cpu().double()
            off = 0
            for stems, B in zip(names, sizes):  # the reference's row order: per batch, t-start outer, item inner
                for i in range(len(grid)):
                    for b in range(B):
                        rows.append({"filename": stems[b], "type": dataset_name, "t": int(grid[i]),
                                     "perceptual_difference": float(scores[i, off + b, 0]),
                                     "mse": float(scores[i, off + b, 1])})
                off += B
        local_rank = int(os.environ["LOCAL_RANK"])
        if local_rank != 0:
            sys.stdout = sys.stderr = open(os.devnull, "w")
        return rows

    def reconstruct(self, args):
        if bool(args.run_val):
            pd.DataFrame(self.get_scores(self.val_loader, "val", args.inference_skip_factor)).to_csv(
                self.out_dir / "results_val.csv")
        if bool(args.run_in):
            pd.DataFrame(self.get_scores(self.in_loader, "in", args.inference_skip_factor)).to_csv(
                self.out_dir / "results_in.csv")
        if bool(args.run_out):
            for out in args.out_ids.split(","):
                print(out)
                if "vflip" in out:
                    out = out.replace("_vflip", "")
                    out_loader = self._loader(args, out, args.first_n, add_vflip=True)
                    dataset_name = Path(out).stem.split("_")[0] + "_vflip"
                elif "hflip" in out:
                    out = out.replace("_hflip", "")
                    out_loader = self._loader(args, out, args.first_n, add_hflip=True)
                    dataset_name = Path(out).stem.split("_")[0] + "_hflip"
                else:
                    out_loader = self._loader(args, out, args.first_n)
                    dataset_name = Path(out).stem.split("_")[0]
                # every out-of-distribution set gets type "out" (reference :328); the name only reaches the file name
                results_list = self.get_scores(out_loader, "out", args.inference_skip_factor)
                pd.DataFrame(results_list).to_csv(self.out_dir / f"results_{dataset_name}.csv")
