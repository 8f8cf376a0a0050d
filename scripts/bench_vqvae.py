"""Throughput of the VQ-VAE stage-1 engine on the reference's README configuration (README.md:153-159: four (2,4,1,1)
downsamplings, 256 channels, 3 residual layers per level, 2048 x 128 codebook) - encode runs once per batch, decode once
per t-start (src/trainers/reconstruct.py:124,166).
  python scripts/bench_vqvae.py [--batch 2] [--size 128]"""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ddpm_ood_b200.vqvae import VQVAE  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--iters", type=int, default=5)
    args = ap.parse_args()
    cfg = dict(spatial_dims=3, in_channels=1, out_channels=1, num_channels=[256] * 4, num_res_layers=3,
               num_res_channels=[256] * 4, downsample_parameters=[[2, 4, 1, 1]] * 4,
               upsample_parameters=[[2, 4, 1, 1, 0]] * 4, num_embeddings=2048, embedding_dim=128)
    torch.manual_seed(0)
    for precise in (True, False):
        m = VQVAE(**cfg, precise_encode=precise).cuda().eval()
        x = torch.rand((args.batch, 1) + (args.size,) * 3, device="cuda")
        lat = m.encode_stage_2_inputs(x)
        img = m.decode_stage_2_outputs(lat)
        torch.cuda.synchronize()

        def timed(fn):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / args.iters

        t_enc = timed(lambda: m.encode_stage_2_inputs(x))
        t_dec = timed(lambda: m.decode_stage_2_outputs(lat))
        print(f"precise_encode={precise}: image {tuple(x.shape)} -> latent {tuple(lat.shape)} -> {tuple(img.shape)}; "
              f"encode {t_enc:.1f} ms, decode {t_dec:.1f} ms per batch of {args.batch}; "
              f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
        del m


if __name__ == "__main__":
    main()
