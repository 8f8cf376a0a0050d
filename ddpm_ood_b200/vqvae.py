"""Drop-in `VQVAE` (monai-generative's `generative.networks.nets.VQVAE`) for the latent-diffusion side of the
reconstruction path: the reference builds it from `vqvae_config.json` next to the checkpoint
(src/trainers/base.py:44-61) and calls `encode_stage_2_inputs` once per batch (src/trainers/reconstruct.py:124) and
`decode_stage_2_outputs` once per t-start (:166).

Same constructor kwargs, `state_dict` keys (MONAI module names, including the quantizer's EMA buffers) and method names;
every tensor operation is one call into libddpm_ood_b200.so (csrc/vqvae.cu: tcgen05 implicit-GEMM convs, fp32 codebook
search). There is no PyTorch fallback: CPU tensors raise.

Encoder precision: the reference encodes in fp32 (src/trainers/reconstruct.py:124 sits outside its autocast block) and
the nearest-row search turns a 1e-3 relative error of the latent into a DIFFERENT codebook row wherever two rows are
nearly equidistant. `precise_encode=True` (default; an engine option, not a MONAI kwarg) therefore carries every encoder
activation and weight as fp16 hi + fp16 lo halves - three tcgen05 products per MAC with fp32 accumulation, ~22 mantissa
bits - so the rows agree with an fp32 encoder except at fp32-noise ties; `precise_encode=False` runs plain fp16 operands
(3x fewer encoder FLOPs, ~97-99 % identical rows). The decoder (:166, also outside the reference's autocast block) has no
such discontinuity: it runs fp16 operands / fp32 accumulation and is held to 3e-3 relative L2 of the fp32 oracle
(tests/test_vqvae_gpu.py).

Supported configuration: the reference's (README.md:153-159) - every level downsamples with (stride 2, kernel 4,
dilation 1, padding 1) and upsamples with (2, 4, 1, 1, 0), channel counts are multiples of 128, ReLU, no dropout, no
output activation. Anything else raises at construction.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _lib
from .networks import _conv_init


class VQVAE(nn.Module):
    def __init__(
        self,
        spatial_dims: int,
        in_channels: int,
        out_channels: int,
        num_channels: Sequence[int] = (96, 96, 192),
        num_res_layers: int = 3,
        num_res_channels: Sequence[int] | int = (96, 96, 192),
        downsample_parameters: Sequence[Tuple[int, int, int, int]] = ((2, 4, 1, 1), (2, 4, 1, 1), (2, 4, 1, 1)),
        upsample_parameters: Sequence[Tuple[int, int, int, int, int]] = ((2, 4, 1, 1, 0), (2, 4, 1, 1, 0), (2, 4, 1, 1, 0)),
        num_embeddings: int = 32,
        embedding_dim: int = 64,
        embedding_init: str = "normal",
        commitment_cost: float = 0.25,
        decay: float = 0.5,
        epsilon: float = 1e-5,
        dropout: float = 0.0,
        act="RELU",
        output_act=None,
        ddp_sync: bool = True,
        use_checkpointing: bool = False,
        precise_encode: bool = True,
    ) -> None:
        super().__init__()
        if isinstance(num_res_channels, int):
            num_res_channels = (num_res_channels,) * len(num_channels)
        num_channels = tuple(int(c) for c in num_channels)
        num_res_channels = tuple(int(c) for c in num_res_channels)
        if len(num_res_channels) != len(num_channels):
            raise ValueError("`num_res_channels` should be a single integer or a tuple of integers with the same length as "
                             "`num_channels`.")
        if len(downsample_parameters) != len(num_channels) or len(upsample_parameters) != len(num_channels):
            raise ValueError("`downsample_parameters` / `upsample_parameters` should have the same length as `num_channels`.")
        if any(tuple(p) != (2, 4, 1, 1) for p in downsample_parameters):
            raise NotImplementedError("the B200 engine implements downsample_parameters (2, 4, 1, 1) per level "
                                      "(the reference configuration, README.md:153)")
        if any(tuple(p) != (2, 4, 1, 1, 0) for p in upsample_parameters):
            raise NotImplementedError("the B200 engine implements upsample_parameters (2, 4, 1, 1, 0) per level "
                                      "(the reference configuration, README.md:154)")
        if dropout:
            raise NotImplementedError("dropout is inactive at inference; pass dropout=0.0")
        if output_act is not None:
            raise NotImplementedError("output_act is not used by the reference (src/trainers/vqvae_trainer.py:52-68)")
        if str(act).upper().split(".")[-1] != "RELU":
            raise NotImplementedError("only ReLU activations are implemented (the VQVAE default)")
        self.spatial_dims = spatial_dims
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.num_channels = num_channels
        self.num_res_layers = num_res_layers
        self.num_res_channels = num_res_channels
        self.num_embeddings = num_embeddings
        self.embedding_dim = embedding_dim
        self.latent_channels = embedding_dim
        # B200-engine option (not a MONAI kwarg): encoder on split-precision operands, see the module docstring
        self.precise_encode = bool(precise_encode)
        self._handle: Optional[C.c_void_p] = None
        self._handle_device = None
        self._synced = None
        self._ws: Dict[Tuple[int, ...], torch.Tensor] = {}
        self._build_parameters(embedding_init)

    # ------------------------------------------------------------------ parameter tree (MONAI key names)
    def _add(self, path: str, tensor: torch.Tensor, buffer: bool = False) -> None:
        parts = path.split(".")
        mod: nn.Module = self
        for p in parts[:-1]:
            if p not in mod._modules:
                mod.add_module(p, nn.Module())
            mod = mod._modules[p]
        if buffer:
            mod.register_buffer(parts[-1], tensor)
        else:
            mod.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=False))

    def _add_conv(self, path: str, cout: int, cin: int, k: int, transposed: bool = False) -> None:
        shape = ((cin, cout) if transposed else (cout, cin)) + (k,) * self.spatial_dims
        w, _ = _conv_init(shape)
        b = torch.zeros(cout).uniform_(-0.05, 0.05)
        self._add(path + ".weight", w)
        self._add(path + ".bias", b)

    def _add_res(self, path: str, c: int, r: int) -> None:
        self._add_conv(path + ".conv1.conv", r, c, 3)
        self._add_conv(path + ".conv2.conv", c, r, 3)

    def _build_parameters(self, embedding_init: str) -> None:
        nc, nr = self.num_channels, self.num_res_channels
        b = 0
        for i in range(len(nc)):
            self._add_conv(f"encoder.blocks.{b}.conv", nc[i], self.in_channels if i == 0 else nc[i - 1], 4)
            b += 1
            for _ in range(self.num_res_layers):
                self._add_res(f"encoder.blocks.{b}", nc[i], nr[i])
                b += 1
        self._add_conv(f"encoder.blocks.{b}.conv", self.embedding_dim, nc[-1], 3)
        rc, rr = list(reversed(nc)), list(reversed(nr))
        self._add_conv("decoder.blocks.0.conv", rc[0], self.embedding_dim, 3)
        b = 1
        for i in range(len(nc)):
            for _ in range(self.num_res_layers):
                self._add_res(f"decoder.blocks.{b}", rc[i], rr[i])
                b += 1
            last = i == len(nc) - 1
            self._add_conv(f"decoder.blocks.{b}.conv", self.out_channels if last else rc[i + 1], rc[i], 4, transposed=True)
            b += 1
        emb = torch.empty(self.num_embeddings, self.embedding_dim)
        if embedding_init == "kaiming_uniform":
            nn.init.kaiming_uniform_(emb, mode="fan_in", nonlinearity="linear")
        else:
            emb.normal_()
        self._add("quantizer.quantizer.embedding.weight", emb)
        self._add("quantizer.quantizer.ema_cluster_size", torch.zeros(self.num_embeddings), buffer=True)
        self._add("quantizer.quantizer.ema_w", emb.clone(), buffer=True)

    # ------------------------------------------------------------------ engine handle
    def _release(self) -> None:
        if self._handle is not None:
            _lib.lib().ddpm_vqvae_destroy(self._handle)
            self._handle = None
            self._ws.clear()
            self._synced = None

    def __del__(self):  # pragma: no cover
        try:
            self._release()
        except Exception:
            pass

    def _sync(self) -> None:
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise _lib.DdpmError("VQVAE (B200 engine) needs its parameters on a CUDA device; there is no CPU fallback")
        L = _lib.lib()
        if self._handle is not None and self._handle_device != dev:
            self._release()
        ver = tuple((p._version, p.data_ptr()) for p in self.parameters())
        if self._handle is not None and ver == self._synced:
            return
        with torch.cuda.device(dev):
            if self._handle is None:
                cfg = _lib.VqVaeConfig()
                cfg.spatial_dims = self.spatial_dims
                cfg.in_channels, cfg.out_channels = self.in_channels, self.out_channels
                cfg.num_levels = len(self.num_channels)
                cfg.num_res_layers = self.num_res_layers
                for i, (c, r) in enumerate(zip(self.num_channels, self.num_res_channels)):
                    cfg.num_channels[i] = c
                    cfg.num_res_channels[i] = r
                cfg.num_embeddings, cfg.embedding_dim = self.num_embeddings, self.embedding_dim
                cfg.precise_encode = 1 if self.precise_encode else 0
                h = C.c_void_p()
                _lib.check(L.ddpm_vqvae_create(C.byref(cfg), C.byref(h)), "ddpm_vqvae_create")
                self._handle = h
                self._handle_device = dev
            stream = torch.cuda.current_stream().cuda_stream
            for name, p in self.named_parameters():
                t = p.detach().float().contiguous()
                _lib.check(L.ddpm_vqvae_set_param(self._handle, name.encode(), t.data_ptr(), t.numel(), stream),
                           f"ddpm_vqvae_set_param({name})")
            _lib.check(L.ddpm_vqvae_finalize(self._handle, stream), "ddpm_vqvae_finalize")
            torch.cuda.current_stream().synchronize()
        self._synced = ver

    def _dims(self, spatial: Sequence[int]) -> Tuple[int, int, int]:
        if self.spatial_dims == 2:
            return 1, int(spatial[0]), int(spatial[1])
        return int(spatial[0]), int(spatial[1]), int(spatial[2])

    def _workspace(self, n: int, d: int, h: int, w: int, device) -> torch.Tensor:
        key = (n, d, h, w)
        ws = self._ws.get(key)
        if ws is None:
            need = _lib.lib().ddpm_vqvae_workspace_bytes(self._handle, n, d, h, w)
            if need <= 0:
                _lib.check(1, "ddpm_vqvae_workspace_bytes")
            ws = torch.empty(need, dtype=torch.uint8, device=device)
            self._ws[key] = ws
        return ws

    # ------------------------------------------------------------------ engine calls
    @torch.no_grad()
    def _encode(self, images: torch.Tensor, want_indices: bool):
        if not images.is_cuda:
            raise _lib.DdpmError("VQVAE needs CUDA tensors; there is no CPU fallback")
        if images.dim() != self.spatial_dims + 2 or images.shape[1] != self.in_channels:
            raise ValueError(f"expected [B, {self.in_channels}, " + "*, " * (self.spatial_dims - 1) + f"*], got {tuple(images.shape)}")
        self._sync()
        x = images.detach().float().contiguous()
        n = x.shape[0]
        d, h, w = self._dims(x.shape[2:])
        f = 1 << len(self.num_channels)
        lat_sp = tuple(s // f for s in x.shape[2:])
        latent = torch.empty((n, self.embedding_dim) + lat_sp, dtype=torch.float32, device=x.device)
        idx = torch.empty((n,) + lat_sp, dtype=torch.int32, device=x.device) if want_indices else None
        with torch.cuda.device(x.device):
            ws = self._workspace(n, d, h, w, x.device)
            _lib.check(_lib.lib().ddpm_vqvae_encode(self._handle, x.data_ptr(), latent.data_ptr(),
                                                    idx.data_ptr() if idx is not None else None, n, d, h, w,
                                                    ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream),
                       "ddpm_vqvae_encode")
        return latent, idx

    @torch.no_grad()
    def _decode(self, z: Optional[torch.Tensor], indices: Optional[torch.Tensor], want_indices: bool = False):
        src = z if z is not None else indices
        if not src.is_cuda:
            raise _lib.DdpmError("VQVAE needs CUDA tensors; there is no CPU fallback")
        self._sync()
        f = 1 << len(self.num_channels)
        if z is not None:
            if z.dim() != self.spatial_dims + 2 or z.shape[1] != self.embedding_dim:
                raise ValueError(f"expected a [B, {self.embedding_dim}, ...] latent, got {tuple(z.shape)}")
            zz = z.detach().float().contiguous()
            lat_sp = tuple(zz.shape[2:])
        else:
            zz = None
            indices = indices.detach().to(torch.int32).contiguous()
            lat_sp = tuple(indices.shape[1:])
        n = src.shape[0]
        img_sp = tuple(s * f for s in lat_sp)
        d, h, w = self._dims(img_sp)
        image = torch.empty((n, self.out_channels) + img_sp, dtype=torch.float32, device=src.device)
        idx_out = torch.empty((n,) + lat_sp, dtype=torch.int32, device=src.device) if want_indices else None
        with torch.cuda.device(src.device):
            ws = self._workspace(n, d, h, w, src.device)
            _lib.check(_lib.lib().ddpm_vqvae_decode(self._handle, zz.data_ptr() if zz is not None else None,
                                                    indices.data_ptr() if zz is None else None, image.data_ptr(),
                                                    idx_out.data_ptr() if idx_out is not None else None, n, d, h, w,
                                                    ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream),
                       "ddpm_vqvae_decode")
        return image, idx_out

    # ------------------------------------------------------------------ reference surface
    def index_quantize(self, images: torch.Tensor) -> torch.Tensor:
        return self._encode(images, True)[1].long()

    def decode_samples(self, embedding_indices: torch.Tensor) -> torch.Tensor:
        return self._decode(None, embedding_indices)[0]

    def encode_stage_2_inputs(self, x: torch.Tensor) -> torch.Tensor:
        """encode + quantize: the nearest codebook row per latent position, [B, embedding_dim, ...]."""
        return self._encode(x, False)[0]

    def decode_stage_2_outputs(self, z: torch.Tensor) -> torch.Tensor:
        """quantize + decode."""
        return self._decode(z, None)[0]

    def decode(self, quantizations: torch.Tensor) -> torch.Tensor:
        """Decoder on already-quantised latents (codebook rows are their own nearest rows, so this equals
        decode_stage_2_outputs)."""
        return self._decode(quantizations, None)[0]

    def quantize(self, encodings: torch.Tensor):
        """(quantised latent, loss placeholder) for a raw latent: the rows the decoder would use."""
        _, idx = self._decode(encodings, None, want_indices=True)
        emb = self.quantizer.quantizer.embedding.weight
        perm = [0, self.spatial_dims + 1] + list(range(1, self.spatial_dims + 1))
        return emb[idx.long()].permute(perm).contiguous(), torch.zeros((), device=encodings.device)

    def forward(self, images: torch.Tensor):
        latent, _ = self._encode(images, False)
        # the codebook rows come back exactly (x + (q - x) differs from q by at most one ulp), so re-quantising them
        # selects the same rows: decode(quantize(encode(x)))
        return self._decode(latent, None)[0], torch.zeros((), device=images.device)

    def launch_count(self) -> int:
        return int(_lib.lib().ddpm_vqvae_launch_count(self._handle)) if self._handle is not None else 0
