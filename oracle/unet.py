"""ORACLE (test infrastructure, not product code): fp32 PyTorch restatement of monai-generative's
`DiffusionModelUNet` as configured by the reference at src/trainers/base.py:66-86 and called at
src/trainers/reconstruct.py:150-153.

PARITY UNPINNED: `generative` (monai-generative, unpinned in requirements.txt:5) is not in /root/reference nor
installable here, so this file restates its published architecture rather than importing it.
The only reference-side pins are the constructor kwargs (base.py:66-75), the call form (trainers/reconstruct.py:151-153)
and the checkpoint key contract (base.py:145). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.

state_dict key names follow MONAI's (`Convolution` wraps a `.conv`), so a reference-trained checkpoint would load:
conv_in.conv, time_embed.{0,2}, down_blocks.i.resnets.j.{norm1,conv1.conv,time_emb_proj,norm2,conv2.conv,
skip_connection.conv}, down_blocks.i.attentions.j.{norm,to_q,to_k,to_v,proj_attn}, down_blocks.i.downsampler.op.conv,
middle_block.{resnet_1,attention,resnet_2}, up_blocks.i.{resnets,attentions}.j, up_blocks.i.upsampler.conv.conv,
out.0, out.2.conv.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F


def _conv_cls(spatial_dims: int):
    return {2: nn.Conv2d, 3: nn.Conv3d}[spatial_dims]


class Convolution(nn.Sequential):
    """MONAI `Convolution(..., conv_only=True)`: a Sequential whose only child is named `conv`."""

    def __init__(self, spatial_dims, in_channels, out_channels, strides=1, kernel_size=3, padding=1):
        super().__init__()
        self.add_module(
            "conv",
            _conv_cls(spatial_dims)(in_channels, out_channels, kernel_size=kernel_size, stride=strides, padding=padding),
        )


def zero_module(m: nn.Module) -> nn.Module:
    for p in m.parameters():
        p.detach().zero_()
    return m


def get_timestep_embedding(timesteps: torch.Tensor, embedding_dim: int, max_period: int = 10000) -> torch.Tensor:
    """Sinusoidal embedding, cat([cos, sin])."""
    half = embedding_dim // 2
    exponent = -math.log(max_period) * torch.arange(0, half, dtype=torch.float32, device=timesteps.device)
    freqs = torch.exp(exponent / half)
    args = timesteps[:, None].float() * freqs[None, :]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if embedding_dim % 2 == 1:
        emb = F.pad(emb, (0, 1, 0, 0))
    return emb


class ResnetBlock(nn.Module):
    def __init__(self, spatial_dims, in_channels, temb_channels, out_channels=None, norm_num_groups=32, norm_eps=1e-6):
        super().__init__()
        self.spatial_dims = spatial_dims
        out_channels = out_channels or in_channels
        self.norm1 = nn.GroupNorm(norm_num_groups, in_channels, eps=norm_eps, affine=True)
        self.nonlinearity = nn.SiLU()
        self.conv1 = Convolution(spatial_dims, in_channels, out_channels)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = nn.GroupNorm(norm_num_groups, out_channels, eps=norm_eps, affine=True)
        self.conv2 = zero_module(Convolution(spatial_dims, out_channels, out_channels))
        if out_channels == in_channels:
            self.skip_connection = nn.Identity()
        else:
            self.skip_connection = Convolution(spatial_dims, in_channels, out_channels, kernel_size=1, padding=0)

    def forward(self, x, emb):
        h = self.conv1(self.nonlinearity(self.norm1(x)))
        temb = self.time_emb_proj(self.nonlinearity(emb))
        temb = temb[(...,) + (None,) * self.spatial_dims]
        h = h + temb
        h = self.conv2(self.nonlinearity(self.norm2(h)))
        return self.skip_connection(x) + h


class AttentionBlock(nn.Module):
    def __init__(self, spatial_dims, num_channels, num_head_channels=None, norm_num_groups=32, norm_eps=1e-6):
        super().__init__()
        self.spatial_dims = spatial_dims
        self.num_heads = num_channels // num_head_channels if num_head_channels is not None else 1
        self.scale = 1.0 / math.sqrt(num_channels / self.num_heads)
        self.norm = nn.GroupNorm(norm_num_groups, num_channels, eps=norm_eps, affine=True)
        self.to_q = nn.Linear(num_channels, num_channels)
        self.to_k = nn.Linear(num_channels, num_channels)
        self.to_v = nn.Linear(num_channels, num_channels)
        self.proj_attn = nn.Linear(num_channels, num_channels)

    def forward(self, x):
        residual = x
        b, c = x.shape[:2]
        spatial = x.shape[2:]
        x = self.norm(x)
        x = x.reshape(b, c, -1).transpose(1, 2)  # [B, T, C]
        q, k, v = self.to_q(x), self.to_k(x), self.to_v(x)
        hd = c // self.num_heads

        def split(t):  # [B, T, C] -> [B*heads, T, hd]
            return t.reshape(b, -1, self.num_heads, hd).permute(0, 2, 1, 3).reshape(b * self.num_heads, -1, hd)

        q, k, v = split(q), split(k), split(v)
        scores = torch.bmm(q, k.transpose(-1, -2)) * self.scale
        probs = scores.softmax(dim=-1)
        o = torch.bmm(probs, v)
        o = o.reshape(b, self.num_heads, -1, hd).permute(0, 2, 1, 3).reshape(b, -1, c)
        o = self.proj_attn(o)
        o = o.transpose(-1, -2).reshape(b, c, *spatial)
        return o + residual


class Downsample(nn.Module):
    def __init__(self, spatial_dims, num_channels):
        super().__init__()
        self.op = Convolution(spatial_dims, num_channels, num_channels, strides=2, kernel_size=3, padding=1)

    def forward(self, x, emb=None):
        return self.op(x)


class Upsample(nn.Module):
    def __init__(self, spatial_dims, num_channels):
        super().__init__()
        self.conv = Convolution(spatial_dims, num_channels, num_channels)

    def forward(self, x, emb=None):
        x = F.interpolate(x, scale_factor=2.0, mode="nearest")
        return self.conv(x)


class DownBlock(nn.Module):
    def __init__(self, spatial_dims, in_channels, out_channels, temb_channels, num_res_blocks, norm_num_groups,
                 norm_eps, add_downsample, with_attn, num_head_channels):
        super().__init__()
        resnets, attentions = [], []
        for i in range(num_res_blocks):
            resnets.append(ResnetBlock(spatial_dims, in_channels if i == 0 else out_channels, temb_channels,
                                       out_channels, norm_num_groups, norm_eps))
            if with_attn:
                attentions.append(AttentionBlock(spatial_dims, out_channels, num_head_channels, norm_num_groups,
                                                 norm_eps))
        self.resnets = nn.ModuleList(resnets)
        if with_attn:
            self.attentions = nn.ModuleList(attentions)
        self.with_attn = with_attn
        self.downsampler = Downsample(spatial_dims, out_channels) if add_downsample else None

    def forward(self, h, temb):
        outs = []
        for i, r in enumerate(self.resnets):
            h = r(h, temb)
            if self.with_attn:
                h = self.attentions[i](h)
            outs.append(h)
        if self.downsampler is not None:
            h = self.downsampler(h, temb)
            outs.append(h)
        return h, outs


class MidBlock(nn.Module):
    def __init__(self, spatial_dims, in_channels, temb_channels, norm_num_groups, norm_eps, num_head_channels):
        super().__init__()
        self.resnet_1 = ResnetBlock(spatial_dims, in_channels, temb_channels, in_channels, norm_num_groups, norm_eps)
        self.attention = AttentionBlock(spatial_dims, in_channels, num_head_channels, norm_num_groups, norm_eps)
        self.resnet_2 = ResnetBlock(spatial_dims, in_channels, temb_channels, in_channels, norm_num_groups, norm_eps)

    def forward(self, h, temb):
        return self.resnet_2(self.attention(self.resnet_1(h, temb)), temb)


class UpBlock(nn.Module):
    def __init__(self, spatial_dims, in_channels, prev_output_channel, out_channels, temb_channels, num_res_blocks,
                 norm_num_groups, norm_eps, add_upsample, with_attn, num_head_channels):
        super().__init__()
        resnets, attentions = [], []
        for i in range(num_res_blocks):
            res_skip_channels = in_channels if (i == num_res_blocks - 1) else out_channels
            resnet_in_channels = prev_output_channel if i == 0 else out_channels
            resnets.append(ResnetBlock(spatial_dims, resnet_in_channels + res_skip_channels, temb_channels,
                                       out_channels, norm_num_groups, norm_eps))
            if with_attn:
                attentions.append(AttentionBlock(spatial_dims, out_channels, num_head_channels, norm_num_groups,
                                                 norm_eps))
        self.resnets = nn.ModuleList(resnets)
        if with_attn:
            self.attentions = nn.ModuleList(attentions)
        self.with_attn = with_attn
        self.upsampler = Upsample(spatial_dims, out_channels) if add_upsample else None

    def forward(self, h, res_list: List[torch.Tensor], temb):
        for i, r in enumerate(self.resnets):
            skip = res_list[-1]
            res_list = res_list[:-1]
            h = torch.cat([h, skip], dim=1)
            h = r(h, temb)
            if self.with_attn:
                h = self.attentions[i](h)
        if self.upsampler is not None:
            h = self.upsampler(h, temb)
        return h


class DiffusionModelUNet(nn.Module):
    """Unconditioned (with_conditioning=False) DiffusionModelUNet."""

    def __init__(
        self,
        spatial_dims: int,
        in_channels: int,
        out_channels: int,
        num_res_blocks: Sequence[int] | int = (2, 2, 2, 2),
        num_channels: Sequence[int] = (32, 64, 64, 64),
        attention_levels: Sequence[bool] = (False, False, True, True),
        norm_num_groups: int = 32,
        norm_eps: float = 1e-6,
        resblock_updown: bool = False,
        num_head_channels: int | Sequence[int] = 8,
        with_conditioning: bool = False,
    ):
        super().__init__()
        if with_conditioning or resblock_updown:
            raise NotImplementedError("oracle restates only the configuration used by the reference hot path")
        if isinstance(num_res_blocks, int):
            num_res_blocks = (num_res_blocks,) * len(num_channels)
        if isinstance(num_head_channels, int):
            num_head_channels = (num_head_channels,) * len(attention_levels)
        self.spatial_dims = spatial_dims
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.block_out_channels = tuple(num_channels)
        self.num_res_blocks = tuple(num_res_blocks)
        self.attention_levels = tuple(attention_levels)
        self.num_head_channels = tuple(num_head_channels)

        self.conv_in = Convolution(spatial_dims, in_channels, num_channels[0])
        ted = num_channels[0] * 4
        self.time_embed = nn.Sequential(nn.Linear(num_channels[0], ted), nn.SiLU(), nn.Linear(ted, ted))

        self.down_blocks = nn.ModuleList()
        oc = num_channels[0]
        for i in range(len(num_channels)):
            ic, oc = oc, num_channels[i]
            final = i == len(num_channels) - 1
            self.down_blocks.append(DownBlock(spatial_dims, ic, oc, ted, num_res_blocks[i], norm_num_groups, norm_eps,
                                              not final, attention_levels[i], num_head_channels[i]))
        self.middle_block = MidBlock(spatial_dims, num_channels[-1], ted, norm_num_groups, norm_eps,
                                     num_head_channels[-1])
        self.up_blocks = nn.ModuleList()
        rc = list(reversed(num_channels))
        rr = list(reversed(num_res_blocks))
        ra = list(reversed(attention_levels))
        rh = list(reversed(num_head_channels))
        oc = rc[0]
        for i in range(len(rc)):
            prev, oc = oc, rc[i]
            ic = rc[min(i + 1, len(num_channels) - 1)]
            final = i == len(num_channels) - 1
            self.up_blocks.append(UpBlock(spatial_dims, ic, prev, oc, ted, rr[i] + 1, norm_num_groups, norm_eps,
                                          not final, ra[i], rh[i]))
        self.out = nn.Sequential(
            nn.GroupNorm(norm_num_groups, num_channels[0], eps=norm_eps, affine=True),
            nn.SiLU(),
            zero_module(Convolution(spatial_dims, num_channels[0], out_channels)),
        )

    def forward(self, x: torch.Tensor, timesteps: torch.Tensor, context: Optional[torch.Tensor] = None):
        t_emb = get_timestep_embedding(timesteps, self.block_out_channels[0]).to(dtype=x.dtype)
        emb = self.time_embed(t_emb)
        h = self.conv_in(x)
        skips = [h]
        for blk in self.down_blocks:
            h, outs = blk(h, emb)
            skips.extend(outs)
        h = self.middle_block(h, emb)
        for blk in self.up_blocks:
            n = len(blk.resnets)
            res = skips[-n:]
            skips = skips[:-n]
            h = blk(h, res, emb)
        return self.out(h)


def make_small(spatial_dims: int, channels: int) -> DiffusionModelUNet:
    """`--model_type small`, src/trainers/base.py:66-75."""
    return DiffusionModelUNet(spatial_dims=spatial_dims, in_channels=channels, out_channels=channels,
                              num_channels=(128, 256, 256), attention_levels=(False, False, True), num_res_blocks=1,
                              num_head_channels=256, with_conditioning=False)


def make_big(spatial_dims: int, channels: int) -> DiffusionModelUNet:
    """`--model_type big`, src/trainers/base.py:76-86."""
    return DiffusionModelUNet(spatial_dims=spatial_dims, in_channels=channels, out_channels=channels,
                              num_channels=(256, 512, 768), attention_levels=(True, True, True), num_res_blocks=2,
                              num_head_channels=256, with_conditioning=False)


def randomize_(model: nn.Module, seed: int = 0, std: float = 0.02) -> nn.Module:
    """Non-zero everywhere (the zero-initialised convs would make parity trivial): conv/linear weights
    ~ N(0, 1/fan_in)-scaled so activations stay O(1) through ~40 layers, biases ~ N(0, std), norm gains ~ 1 + N(0, .1).
    """
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in sorted(model.named_parameters()):
            if p.dim() >= 2:
                fan_in = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) * (1.0 / math.sqrt(fan_in)))
            elif "norm" in name or name.startswith("out.0"):
                if name.endswith("weight"):
                    p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
                else:
                    p.copy_(0.1 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(std * torch.randn(p.shape, generator=g))
    return model


def count_flops(model: DiffusionModelUNet, spatial: Sequence[int]) -> float:
    """FLOPs (2 per MAC) of one forward on ONE image: convs, linears and the two attention matmuls."""
    total = 0.0
    hooks = []

    def conv_hook(m, inp, out):
        nonlocal total
        k = 1
        for s in m.kernel_size:
            k *= s
        total += 2.0 * out.numel() * m.in_channels * k

    def lin_hook(m, inp, out):
        nonlocal total
        total += 2.0 * out.numel() * m.in_features

    def attn_hook(m, inp, out):
        nonlocal total
        x = inp[0]
        t = 1
        for s in x.shape[2:]:
            t *= s
        c = x.shape[1]
        total += 2.0 * 2.0 * t * t * c  # QK^T and PV

    for m in model.modules():
        if isinstance(m, (nn.Conv2d, nn.Conv3d)):
            hooks.append(m.register_forward_hook(conv_hook))
        elif isinstance(m, nn.Linear):
            hooks.append(m.register_forward_hook(lin_hook))
        elif isinstance(m, AttentionBlock):
            hooks.append(m.register_forward_hook(attn_hook))
    with torch.no_grad():
        x = torch.zeros((1, model.in_channels) + tuple(spatial))
        model(x, torch.zeros(1, dtype=torch.long))
    for h in hooks:
        h.remove()
    return total
