"""Time the per-item 2.5-D LPIPS of a 3-D volume (src/trainers/reconstruct.py:181-187).  python scripts/bench_lpips3d.py [--size 128]"""
import argparse
import sys
import warnings
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ddpm_ood_b200.losses import PerceptualLoss  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pl = PerceptualLoss(dimensions=3, include_pixel_loss=False, is_fake_3d=True, lpips_normalize=True, spatial=False,
                            allow_synthetic_weights=True).cuda()
    a = torch.rand((1, 1) + (args.size,) * 3, device="cuda")
    b = torch.rand_like(a)
    for _ in range(3):
        pl(a, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        pl(a, b)
    e1.record()
    torch.cuda.synchronize()
    print(f"2.5-D LPIPS of one {args.size}^3 volume pair: {e0.elapsed_time(e1) / args.iters:.2f} ms")


if __name__ == "__main__":
    main()
