"""Simplex noise for `--simplex_noise=1`: the surface of the reference's src/utils/simplex_noise.py that the
reconstruction path touches (`Simplex_CLASS`, `generate_simplex_noise`, :15-97), evaluated on the GPU.

The reference draws a fresh seed per (channel, image) from numpy's global RNG, rebuilds the permutation tables on the
host and calls a numba kernel per (channel, image), each with a D2H / H2D round trip. Here the seeds are drawn the same
way and in the same order (so `np.random.seed(k)` reproduces the reference's noise bit for bit), and one
`ddpm_simplex_noise` call builds all tables and evaluates every pixel."""
from __future__ import annotations

import random

import numpy as np
import torch

from . import _lib

_RANDOM_PARAMS = 23  # entries of the reference's random_param list (:28-52); its draw is overwritten at :69, see below


class Simplex_CLASS:
    """Seed holder. `newSeed()` draws like the reference (:87-90); the tables themselves are built on the device."""

    def __init__(self):
        self.newSeed()

    def newSeed(self, seed=None):
        if not seed:
            seed = np.random.randint(-10000000000, 10000000000)
        self.seed = int(seed)


def generate_simplex_noise(Simplex_instance, x, t, random_param=False, octave=6, persistence=0.8, frequency=64,
                           in_channels=1):
    """fp32 noise of x's shape [B, C, H, W]: noise[j, i] = 6-octave 3-D OpenSimplex on the plane z = t[j] / frequency with
    its own freshly drawn seed. `random_param=True` consumes Python's RNG like the reference but, as there (:55-78: the
    result of the random-parameter branch is overwritten by the fixed-parameter call), does not change the output."""
    if not x.is_cuda:
        raise _lib.DdpmError("generate_simplex_noise needs a CUDA tensor; there is no CPU fallback")
    if x.dim() != 4:
        raise NotImplementedError("simplex noise is defined for 2-D images [B, C, H, W]")
    b, c, h, w = x.shape
    if c != in_channels or len(t) != b:
        raise ValueError("in_channels / t do not match x")
    seeds = np.empty((in_channels, b), dtype=np.int64)
    for i in range(in_channels):
        for j in range(b):
            Simplex_instance.newSeed()
            if random_param:
                random.randrange(_RANDOM_PARAMS)  # random.choice over 23 entries draws the same way
            seeds[i, j] = Simplex_instance.seed
    dev = x.device
    seeds_d = torch.from_numpy(seeds.reshape(-1)).to(dev)
    t_d = t.to(device=dev, dtype=torch.int64).contiguous()
    out = torch.empty((b, c, h, w), dtype=torch.float32, device=dev)
    ws = torch.empty(in_channels * b * 256, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().ddpm_simplex_noise(seeds_d.data_ptr(), t_d.data_ptr(), out.data_ptr(), ws.data_ptr(), b, c,
                                                 h, w, int(octave), float(persistence), float(frequency),
                                                 torch.cuda.current_stream().cuda_stream), "ddpm_simplex_noise")
    return out
