"""ORACLE (test infrastructure, not product code): fp32 PyTorch restatement of monai-generative's `VQVAE`
(`generative.networks.nets.VQVAE`) as the reference builds it from `vqvae_config.json` at src/trainers/base.py:44-61 and
calls it at src/trainers/reconstruct.py:124 (`encode_stage_2_inputs`) and :166 (`decode_stage_2_outputs`).

PARITY UNPINNED [3P-RECALL]: the `generative` package (unpinned, requirements.txt:5) is absent from /root/reference and not
installable here, and the reference holds no VQ-VAE test or golden vector; this file restates the published 0.2.x
architecture from memory. Reference-side pins: the constructor kwargs (src/trainers/vqvae_trainer.py:52-68: spatial_dims,
in/out_channels, num_res_layers, downsample_parameters, upsample_parameters, num_channels, num_res_channels,
num_embeddings, embedding_dim, decay, commitment_cost, epsilon, dropout, ddp_sync), the README configuration
(README.md:153-159: four (2,4,1,1) downsamplings, four (2,4,1,1,0) upsamplings, 256 channels, 2048 x 128 codebook) and the
two calls above. Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.

state_dict keys follow MONAI's modules so a reference-trained vqvae checkpoint would load: encoder.blocks.i.conv.{weight,
bias} (`Convolution` wraps `.conv`), encoder.blocks.i.{conv1,conv2}.conv.* (residual units), decoder.blocks.i...,
quantizer.quantizer.embedding.weight, buffers quantizer.quantizer.{ema_cluster_size, ema_w}.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


def _conv(spatial_dims: int, transposed: bool):
    return {(2, False): nn.Conv2d, (3, False): nn.Conv3d, (2, True): nn.ConvTranspose2d, (3, True): nn.ConvTranspose3d}[
        (spatial_dims, transposed)]


class Convolution(nn.Sequential):
    """MONAI `Convolution`: child `conv`, then (unless conv_only) ADN with ordering "DA" = dropout (p = 0 here) then ReLU."""

    def __init__(self, spatial_dims, in_channels, out_channels, strides=1, kernel_size=3, padding=1, dilation=1,
                 conv_only=False, is_transposed=False, output_padding=0):
        super().__init__()
        if is_transposed:
            conv = _conv(spatial_dims, True)(in_channels, out_channels, kernel_size=kernel_size, stride=strides,
                                             padding=padding, output_padding=output_padding, dilation=dilation)
        else:
            conv = _conv(spatial_dims, False)(in_channels, out_channels, kernel_size=kernel_size, stride=strides,
                                              padding=padding, dilation=dilation)
        self.add_module("conv", conv)
        if not conv_only:
            adn = nn.Sequential()
            adn.add_module("A", nn.ReLU())
            self.add_module("adn", adn)


class VQVAEResidualUnit(nn.Module):
    def __init__(self, spatial_dims, num_channels, num_res_channels):
        super().__init__()
        self.conv1 = Convolution(spatial_dims, num_channels, num_res_channels)
        self.conv2 = Convolution(spatial_dims, num_res_channels, num_channels, conv_only=True)

    def forward(self, x):
        return F.relu(x + self.conv2(self.conv1(x)), True)


class Encoder(nn.Module):
    def __init__(self, spatial_dims, in_channels, out_channels, num_channels, num_res_layers, num_res_channels,
                 downsample_parameters):
        super().__init__()
        blocks = []
        for i in range(len(num_channels)):
            s, k, d, p = downsample_parameters[i]
            blocks.append(Convolution(spatial_dims, in_channels if i == 0 else num_channels[i - 1], num_channels[i],
                                      strides=s, kernel_size=k, dilation=d, padding=p))
            for _ in range(num_res_layers):
                blocks.append(VQVAEResidualUnit(spatial_dims, num_channels[i], num_res_channels[i]))
        blocks.append(Convolution(spatial_dims, num_channels[-1], out_channels, strides=1, kernel_size=3, padding=1,
                                  conv_only=True))
        self.blocks = nn.ModuleList(blocks)

    def forward(self, x):
        for b in self.blocks:
            x = b(x)
        return x


class Decoder(nn.Module):
    def __init__(self, spatial_dims, in_channels, out_channels, num_channels, num_res_layers, num_res_channels,
                 upsample_parameters, output_act=None):
        super().__init__()
        rc = list(reversed(num_channels))
        rr = list(reversed(num_res_channels))
        blocks = [Convolution(spatial_dims, in_channels, rc[0], strides=1, kernel_size=3, padding=1, conv_only=True)]
        n = len(num_channels)
        for i in range(n):
            for _ in range(num_res_layers):
                blocks.append(VQVAEResidualUnit(spatial_dims, rc[i], rr[i]))
            s, k, d, p, op = upsample_parameters[i]
            blocks.append(Convolution(spatial_dims, rc[i], out_channels if i == n - 1 else rc[i + 1], strides=s,
                                      kernel_size=k, dilation=d, padding=p, output_padding=op, conv_only=(i == n - 1),
                                      is_transposed=True))
        if output_act:
            raise NotImplementedError("output_act is not used by the reference (vqvae_trainer.py:52-68)")
        self.blocks = nn.ModuleList(blocks)

    def forward(self, x):
        for b in self.blocks:
            x = b(x)
        return x


class EMAQuantizer(nn.Module):
    def __init__(self, spatial_dims, num_embeddings, embedding_dim, commitment_cost=0.25, decay=0.99, epsilon=1e-5):
        super().__init__()
        self.spatial_dims = spatial_dims
        self.embedding_dim = embedding_dim
        self.num_embeddings = num_embeddings
        self.embedding = nn.Embedding(num_embeddings, embedding_dim)
        self.embedding.weight.requires_grad = False
        self.commitment_cost = commitment_cost
        self.register_buffer("ema_cluster_size", torch.zeros(num_embeddings))
        self.register_buffer("ema_w", self.embedding.weight.data.clone())
        self.flatten_permutation = [0] + list(range(2, spatial_dims + 2)) + [1]
        self.quantization_permutation = [0, spatial_dims + 1] + list(range(1, spatial_dims + 1))

    def quantize(self, inputs):
        flat = inputs.permute(self.flatten_permutation).contiguous().view(-1, self.embedding_dim)
        distances = ((flat ** 2).sum(dim=1, keepdim=True) + (self.embedding.weight.t() ** 2).sum(dim=0, keepdim=True)
                     - 2 * torch.mm(flat, self.embedding.weight.t()))
        idx = torch.max(-distances, dim=1)[1]
        shape = list(inputs.shape)
        del shape[1]
        return flat, distances, idx.view(shape)

    def embed(self, embedding_indices):
        return self.embedding(embedding_indices).permute(self.quantization_permutation).contiguous()

    def forward(self, inputs):
        _, _, idx = self.quantize(inputs)
        quantized = self.embed(idx)
        loss = self.commitment_cost * F.mse_loss(quantized.detach(), inputs)
        quantized = inputs + (quantized - inputs).detach()  # straight-through estimator
        return quantized, loss, idx


class VectorQuantizer(nn.Module):
    def __init__(self, quantizer):
        super().__init__()
        self.quantizer = quantizer

    def forward(self, inputs):
        quantized, loss, _ = self.quantizer(inputs)
        return loss, quantized

    def embed(self, embedding_indices):
        return self.quantizer.embed(embedding_indices)

    def quantize(self, encodings):
        return self.quantizer(encodings)[2]


class VQVAE(nn.Module):
    def __init__(self, spatial_dims: int, in_channels: int, out_channels: int,
                 num_channels: Sequence[int] = (96, 96, 192), num_res_layers: int = 3,
                 num_res_channels: Sequence[int] = (96, 96, 192),
                 downsample_parameters: Sequence[Tuple[int, int, int, int]] = ((2, 4, 1, 1),) * 3,
                 upsample_parameters: Sequence[Tuple[int, int, int, int, int]] = ((2, 4, 1, 1, 0),) * 3,
                 num_embeddings: int = 32, embedding_dim: int = 64, embedding_init: str = "normal",
                 commitment_cost: float = 0.25, decay: float = 0.5, epsilon: float = 1e-5, dropout: float = 0.0,
                 ddp_sync: bool = True, use_checkpointing: bool = False):
        super().__init__()
        if isinstance(num_res_channels, int):
            num_res_channels = (num_res_channels,) * len(num_channels)
        if dropout:
            raise NotImplementedError("dropout is inactive at inference")
        self.in_channels, self.out_channels, self.spatial_dims = in_channels, out_channels, spatial_dims
        self.num_channels, self.num_embeddings, self.embedding_dim = tuple(num_channels), num_embeddings, embedding_dim
        self.encoder = Encoder(spatial_dims, in_channels, embedding_dim, num_channels, num_res_layers, num_res_channels,
                               downsample_parameters)
        self.decoder = Decoder(spatial_dims, embedding_dim, out_channels, num_channels, num_res_layers, num_res_channels,
                               upsample_parameters)
        self.quantizer = VectorQuantizer(EMAQuantizer(spatial_dims, num_embeddings, embedding_dim, commitment_cost, decay,
                                                      epsilon))

    def encode(self, images):
        return self.encoder(images)

    def quantize(self, encodings):
        loss, x = self.quantizer(encodings)
        return x, loss

    def decode(self, quantizations):
        return self.decoder(quantizations)

    def index_quantize(self, images):
        return self.quantizer.quantize(self.encode(images))

    def decode_samples(self, embedding_indices):
        return self.decode(self.quantizer.embed(embedding_indices))

    def forward(self, images):
        q, loss = self.quantize(self.encode(images))
        return self.decode(q), loss

    def encode_stage_2_inputs(self, x):
        e, _ = self.quantize(self.encode(x))
        return e

    def decode_stage_2_outputs(self, z):
        e, _ = self.quantize(z)
        return self.decode(e)


def randomize_(model: nn.Module, seed: int = 0) -> nn.Module:
    """Non-degenerate synthetic weights: conv matrices ~ N(0, 2/fan_in) (ReLU gain; transposed convs count the taps that
    reach one output), biases ~ N(0, .02), codebook ~ N(0, 1). Sorted-name order: same values for any module with the
    same state_dict keys (the oracle and the product class)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in sorted(model.named_parameters()):
            if name.endswith("embedding.weight"):
                v = torch.randn(p.shape, generator=g)
            elif p.dim() >= 2:
                transposed = "decoder" in name and p.shape[2] == 4  # ConvTranspose weight [Cin, Cout, k...]
                taps = p[0][0].numel()
                fan_in = (p.shape[0] * taps / (2 ** (p.dim() - 2))) if transposed else p[0].numel()
                v = torch.randn(p.shape, generator=g) * (2.0 / fan_in) ** 0.5
            else:
                v = 0.02 * torch.randn(p.shape, generator=g)
            p.copy_(v)
    return model
