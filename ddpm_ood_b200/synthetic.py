"""Synthetic weights / work accounting for measurement (bench.py, __graft_entry__.smoke, tests).

There is no network for datasets or checkpoints, so benchmarks use random-init weights of the reference architecture
(src/trainers/base.py:66-86). The reference zero-initialises every ResnetBlock's conv2 and the output conv
(`zero_module`), which would make a fresh model output exactly 0 and any parity or timing claim
meaningless, so every parameter is randomised.
"""
from __future__ import annotations

import math
from typing import Sequence

import torch
import torch.nn as nn


def randomize_(model: nn.Module, seed: int = 0, std: float = 0.02) -> nn.Module:
    """Non-zero everywhere: matrices ~ N(0, 1/fan_in) (activations stay O(1) through ~40 layers), norm gains
    ~ 1 + N(0, .1), norm shifts ~ N(0, .1), other biases ~ N(0, std). Walks parameters in sorted-name order, so any two
    modules with the same state_dict keys (ours and the oracle's) get identical values for the same seed."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in sorted(model.named_parameters()):
            if p.dim() >= 2:
                fan_in = p[0].numel()
                v = torch.randn(p.shape, generator=g) * (1.0 / math.sqrt(fan_in))
            elif "norm" in name or name.startswith("out.0"):
                if name.endswith("weight"):
                    v = 1.0 + 0.1 * torch.randn(p.shape, generator=g)
                else:
                    v = 0.1 * torch.randn(p.shape, generator=g)
            else:
                v = std * torch.randn(p.shape, generator=g)
            p.copy_(v.to(p.device))
    return model


def unet_flops_per_image(num_channels: Sequence[int], attention_levels: Sequence[bool], num_res_blocks: Sequence[int],
                         in_channels: int, out_channels: int, spatial: Sequence[int]) -> float:
    """Algorithmic FLOPs (2 per MAC) of one DiffusionModelUNet forward on ONE image: convs, linears (timestep MLP,
    per-block projections, attention q/k/v/proj) and the two attention matmuls; norms/activations excluded.
    Topology per SURVEY.md A.1 (the same walk as ddpm_ood_b200/csrc/engine.cu:init)."""
    sd = len(spatial)
    taps = 3 ** sd
    L = len(num_channels)
    sizes = [tuple(spatial)]
    for _ in range(L - 1):
        sizes.append(tuple((s + 1) // 2 for s in sizes[-1]))

    def S(level: int) -> int:
        n = 1
        for s in sizes[level]:
            n *= s
        return n

    ted = 4 * num_channels[0]
    macs = 0.0
    macs += S(0) * num_channels[0] * taps * in_channels  # conv_in
    macs += num_channels[0] * ted + ted * ted  # time_embed

    def res(level: int, cin: int, cout: int) -> float:
        m = S(level) * cout * taps * cin + S(level) * cout * taps * cout + ted * cout
        if cin != cout:
            m += S(level) * cout * cin
        return m

    def attn(level: int, c: int) -> float:
        t = S(level)
        return 4.0 * t * c * c + 2.0 * t * t * c

    oc = num_channels[0]
    for i in range(L):
        ic, oc = oc, num_channels[i]
        for j in range(num_res_blocks[i]):
            macs += res(i, ic if j == 0 else oc, oc)
            if attention_levels[i]:
                macs += attn(i, oc)
        if i != L - 1:
            macs += S(i + 1) * oc * taps * oc
    cm = num_channels[-1]
    macs += 2 * res(L - 1, cm, cm) + attn(L - 1, cm)
    oc = cm
    for i in range(L):
        lvl = L - 1 - i
        prev, oc = oc, num_channels[lvl]
        ic = num_channels[max(lvl - 1, 0)]
        n = num_res_blocks[lvl] + 1
        for j in range(n):
            res_skip = ic if j == n - 1 else oc
            res_in = prev if j == 0 else oc
            macs += res(lvl, res_in + res_skip, oc)
            if attention_levels[lvl]:
                macs += attn(lvl, oc)
        if i != L - 1:
            macs += S(lvl - 1) * oc * taps * oc
    macs += S(0) * out_channels * taps * num_channels[0]
    return 2.0 * macs


def chain_lengths(num_inference_steps: int, inference_skip_factor: int, num_train_timesteps: int = 1000):
    """UNet evaluations of each t-start's chain: len(timesteps[timesteps <= t_start]) for the PLMS timestep vector and
    the grid `reversed(timesteps)[1::skip]` (src/trainers/reconstruct.py:118-120,149)."""
    ratio = num_train_timesteps // num_inference_steps
    ts = [i * ratio for i in range(num_inference_steps)]
    plms = (ts[:-1] + ts[-2:-1] + ts[-1:])[::-1]
    starts = plms[::-1][1::inference_skip_factor]
    return [sum(1 for t in plms if t <= s) for s in starts]
