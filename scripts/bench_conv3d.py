"""Micro-benchmark of the im2col-tile conv kernel on the small 3-D levels of the latent UNet (BASELINE configs[4]).
  python scripts/bench_conv3d.py [--batch 592]"""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ddpm_ood_b200 import ops  # noqa: E402

SHAPES = [("level2 2^3 256->256", (2, 2, 2), 256, 256), ("level1 4^3 256->256", (4, 4, 4), 256, 256),
          ("level1 4^3 512->256", (4, 4, 4), 512, 256), ("level0 8^3 128->128", (8, 8, 8), 128, 128)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=592)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--impl", type=int, default=0)
    args = ap.parse_args()
    g = torch.Generator(device="cuda").manual_seed(0)
    for name, sp, cin, cout in SHAPES:
        k = 27 * cin
        wp = (torch.randn(cout, k, generator=g, device="cuda") / k ** 0.5).half()
        xs = [torch.randn((args.batch,) + sp + (cin,), generator=g, device="cuda").half() for _ in range(3)]
        outs = [torch.empty((args.batch,) + sp + (cout,), dtype=torch.float16, device="cuda") for _ in range(3)]
        flops = 2.0 * args.batch * sp[0] * sp[1] * sp[2] * cout * k
        for i in range(3):
            ops.conv_forward([xs[i]], [3], wp, cout, out=outs[i], impl=args.impl)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.iters):
            ops.conv_forward([xs[i % 3]], [3], wp, cout, out=outs[i % 3], impl=args.impl)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1000.0 / args.iters
        print(f"{name:24s} {flops / 1e9:7.1f} GF  {us:7.1f} us  {flops / us / 1e6:6.0f} TF/s", flush=True)


if __name__ == "__main__":
    main()
