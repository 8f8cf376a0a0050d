"""Times the UNet's head (conv_in + statistics) and tail (out norm + SiLU + conv_out taps + gather) kernels alone, CUDA
events, L2 flushed between launches. `--once` does a single launch of each (for ncu)."""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ddpm_ood_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--once", action="store_true")
    a = ap.parse_args()
    g = torch.Generator(device="cuda").manual_seed(0)
    n = a.batch
    x = torch.randn((n, 1, 32, 32), generator=g, device="cuda")
    w = torch.randn((128, 1, 3, 3), generator=g, device="cuda") * 0.3
    b = torch.randn(128, generator=g, device="cuda")
    act = torch.randn((n, 32, 32, 128), generator=g, device="cuda").half()
    xs = act.float().reshape(n, 1024, 32, 4)
    st = torch.stack([xs.sum(dim=(1, 3)), (xs * xs).sum(dim=(1, 3))], dim=-1)[:, None].contiguous()
    gamma = torch.ones(128, device="cuda")
    beta = torch.zeros(128, device="cuda")
    wo = torch.randn((1, 128, 3, 3), generator=g, device="cuda") * 0.05
    bo = torch.zeros(1, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    fns = {"conv_in": lambda: ops.conv_in(x, w, b), "out_norm_conv": lambda: ops.out_norm_conv(act, st, gamma, beta, wo, bo, 32, 1e-6)}
    for name, fn in fns.items():
        fn()
        torch.cuda.synchronize()
        if a.once:
            continue
        ts = []
        for _ in range(a.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        print(f"{name}: median {ts[len(ts) // 2]:.1f} us, min {ts[0]:.1f} us (batch {n}; includes torch.empty + launch overhead)")


if __name__ == "__main__":
    main()
