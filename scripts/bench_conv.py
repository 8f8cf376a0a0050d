"""Micro-benchmark of the conv kernels on the UNet's layer shapes (batch 256, FashionMNIST config), isolated from the
engine: CUDA-event time per launch, cycling through enough distinct input/output buffers to defeat L2 residency.
  python scripts/bench_conv.py [--batch 256] [--impls 0,3] [--gn]
"""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ddpm_ood_b200 import ops  # noqa: E402

# (name, H=W, segs [(channels, ksize)], Cout)
SHAPES = [
    ("down0.res.conv1   32px 128->128", 32, [(128, 3)], 128),
    ("up2.res0.conv1    32px 384->128", 32, [(256, 3), (128, 3)], 128),
    ("up2.res0.conv2+sk 32px 128(+384)", 32, [(128, 3), (256, 1), (128, 1)], 128),
    ("up2.res1.conv1    32px 256->128", 32, [(128, 3), (128, 3)], 128),
    ("down1.res.conv1   16px 128->256", 16, [(128, 3)], 256),
    ("down1.res.conv2+sk16px 256(+128)", 16, [(256, 3), (128, 1)], 256),
    ("up1.res0.conv1    16px 512->256", 16, [(256, 3), (256, 3)], 256),
    ("up1.res1.conv2+sk 16px 256(+384)", 16, [(256, 3), (256, 1), (128, 1)], 256),
    ("mid.res.conv1      8px 256->256", 8, [(256, 3)], 256),
    ("up0.res0.conv1     8px 512->256", 8, [(256, 3), (256, 3)], 256),
    ("up0.res0.conv2+sk  8px 256(+512)", 8, [(256, 3), (256, 1), (256, 1)], 256),
    ("attn.qkv 1x1       8px 256->768", 8, [(256, 1)], 768),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--impls", default="0,3")
    ap.add_argument("--gn", action="store_true")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--shapes", type=int, default=len(SHAPES), help="only the first n shapes")
    ap.add_argument("--first", type=int, default=0, help="skip the first n shapes")
    args = ap.parse_args()
    dev = "cuda"
    n = args.batch
    g = torch.Generator(device=dev).manual_seed(0)
    for name, hw, segs, cout in SHAPES[args.first:args.shapes]:
        ktot = sum(c * k * k for c, k in segs)
        wp = (torch.randn(cout, ktot, generator=g, device=dev) / ktot ** 0.5).half()
        bias = torch.randn(cout, generator=g, device=dev)
        nbuf = 3
        xs = [[torch.randn((n, hw, hw, c), generator=g, device=dev).half() for c, _ in segs] for _ in range(nbuf)]
        outs = [torch.empty((n, hw, hw, cout), dtype=torch.float16, device=dev) for _ in range(nbuf)]
        c3 = sum(c for c, k in segs if k == 3)
        ab = torch.stack([0.5 + torch.rand((n, c3), generator=g, device=dev),
                          torch.randn((n, c3), generator=g, device=dev)], dim=-1).contiguous()
        flops = 2.0 * n * hw * hw * cout * ktot
        row = f"{name:36s} {flops / 1e9:7.1f} GF "
        for impl in [int(v) for v in args.impls.split(",")]:
            st = None
            if impl == 3:
                st = torch.empty((n, ops.conv_halo_stats_parts(hw, hw), cout // 4, 2), device=dev)
            else:
                parts = ops.conv_stats_parts(2, 1, hw, hw)
                st = torch.empty((n, parts, cout // 4, 2), device=dev) if parts else None
            kw = dict(bias=bias, impl=impl, stats_out=st)
            if impl == 3 and args.gn:
                kw["gn_scale_shift"] = ab if c3 else torch.cat([ab, torch.stack([torch.ones((n, segs[0][0]), device=dev), torch.zeros((n, segs[0][0]), device=dev)], dim=-1)], dim=1).contiguous()
            for i in range(3):
                ops.conv_forward(xs[i % nbuf], [k for _, k in segs], wp, cout, out=outs[i % nbuf], **kw)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(args.iters):
                ops.conv_forward(xs[i % nbuf], [k for _, k in segs], wp, cout, out=outs[i % nbuf], **kw)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1000.0 / args.iters
            row += f" | impl{impl} {us:7.1f} us {flops / us / 1e6:6.0f} TF/s"
        print(row, flush=True)


if __name__ == "__main__":
    main()
