"""Times generate_simplex_noise (device) for a batch of 1x32x32 images; the host-side seed drawing is inside the timed
call, like in the trainer."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ddpm_ood_b200.simplex_noise import Simplex_CLASS, generate_simplex_noise  # noqa: E402

for b in (256, 1184):
    x = torch.zeros((b, 1, 32, 32), device="cuda")
    t = torch.full((b,), 490, dtype=torch.long)
    s = Simplex_CLASS()
    np.random.seed(0)
    generate_simplex_noise(s, x, t, in_channels=1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        out = generate_simplex_noise(s, x, t, in_channels=1)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 100
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    print(f"batch {b}: {ms:.2f} ms per call wall clock (seed drawing + H2D + 2 kernels), std {float(out.std()):.3f}")
