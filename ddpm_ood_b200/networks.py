"""Drop-in `DiffusionModelUNet` and `PassthroughVQVAE` for the reconstruction hot path.

`DiffusionModelUNet` mirrors the constructor, `state_dict` keys and call form of monai-generative's class as the
reference uses it (src/trainers/base.py:66-86,145; src/trainers/reconstruct.py:150-153), but its forward is ONE call
into the sm_100a engine (libddpm_ood_b200.so: tcgen05 implicit-GEMM convs, fused GroupNorm/SiLU, attention). There is
no PyTorch fallback: calling it without the CUDA library or on a CPU tensor raises.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _lib


class PassthroughVQVAE(nn.Module):
    """Identity stage-1 model for pixel-space DDPMs (src/networks/passthrough_vqvae.py:4-26)."""

    def __init__(self):
        super().__init__()
        self.latent_channels = 1

    def reconstruct(self, x):
        return x

    def decode(self, x):
        return x

    def forward(self, x):
        return x

    def encode_stage_2_inputs(self, x):
        return x

    def decode_stage_2_outputs(self, x):
        return x


def _conv_init(shape, gen=None):
    w = torch.empty(shape)
    nn.init.kaiming_uniform_(w, a=math.sqrt(5))
    fan_in = w[0].numel()
    bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
    b = torch.empty(shape[0]).uniform_(-bound, bound)
    return w, b


class DiffusionModelUNet(nn.Module):
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
        transformer_num_layers: int = 1,
        cross_attention_dim: Optional[int] = None,
        num_class_embeds: Optional[int] = None,
        upcast_attention: bool = False,
        use_flash_attention: bool = False,
    ) -> None:
        super().__init__()
        if with_conditioning or cross_attention_dim is not None or num_class_embeds is not None:
            raise NotImplementedError("conditioning is not on the reconstruction hot path (base.py:74,85)")
        if resblock_updown:
            raise NotImplementedError("resblock_updown=True is not used by the reference")
        if isinstance(num_res_blocks, int):
            num_res_blocks = (num_res_blocks,) * len(num_channels)
        if isinstance(num_head_channels, int):
            num_head_channels = (num_head_channels,) * len(attention_levels)
        if len(num_channels) != len(attention_levels) or len(num_channels) != len(num_res_blocks):
            raise ValueError("num_channels, attention_levels and num_res_blocks must have the same length")
        if any(c % norm_num_groups for c in num_channels):
            raise ValueError("DiffusionModelUNet expects all num_channels being multiple of norm_num_groups")
        self.spatial_dims = spatial_dims
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.block_out_channels = tuple(num_channels)
        self.num_res_blocks = tuple(num_res_blocks)
        self.attention_levels = tuple(bool(a) for a in attention_levels)
        self.num_head_channels = tuple(num_head_channels)
        self.norm_num_groups = norm_num_groups
        self.norm_eps = norm_eps
        self.with_conditioning = False
        self._handle: Optional[C.c_void_p] = None
        self._handle_device: Optional[torch.device] = None
        self._synced_versions: Optional[Tuple[int, ...]] = None
        self._workspaces: Dict[Tuple[int, ...], torch.Tensor] = {}
        self._build_parameters()

    # ------------------------------------------------------------------ parameter tree (MONAI key names)
    def _add(self, path: str, tensor: torch.Tensor) -> None:
        parts = path.split(".")
        mod: nn.Module = self
        for p in parts[:-1]:
            if p not in mod._modules:
                mod.add_module(p, nn.Module())
            mod = mod._modules[p]
        mod.register_parameter(parts[-1], nn.Parameter(tensor))

    def _add_conv(self, path: str, cout: int, cin: int, k: int, zero: bool = False) -> None:
        w, b = _conv_init((cout, cin) + (k,) * self.spatial_dims)
        if zero:
            w.zero_()
            b.zero_()
        self._add(path + ".weight", w)
        self._add(path + ".bias", b)

    def _add_linear(self, path: str, cout: int, cin: int) -> None:
        w, b = _conv_init((cout, cin))
        self._add(path + ".weight", w)
        self._add(path + ".bias", b)

    def _add_norm(self, path: str, c: int) -> None:
        self._add(path + ".weight", torch.ones(c))
        self._add(path + ".bias", torch.zeros(c))

    def _add_resnet(self, path: str, cin: int, cout: int, temb: int) -> None:
        self._add_norm(path + ".norm1", cin)
        self._add_conv(path + ".conv1.conv", cout, cin, 3)
        self._add_linear(path + ".time_emb_proj", cout, temb)
        self._add_norm(path + ".norm2", cout)
        self._add_conv(path + ".conv2.conv", cout, cout, 3, zero=True)  # zero_module
        if cin != cout:
            self._add_conv(path + ".skip_connection.conv", cout, cin, 1)

    def _add_attn(self, path: str, c: int) -> None:
        self._add_norm(path + ".norm", c)
        for n in ("to_q", "to_k", "to_v", "proj_attn"):
            self._add_linear(path + "." + n, c, c)

    def _build_parameters(self) -> None:
        ch = self.block_out_channels
        L = len(ch)
        ted = ch[0] * 4
        self._add_conv("conv_in.conv", ch[0], self.in_channels, 3)
        self._add_linear("time_embed.0", ted, ch[0])
        self._add_linear("time_embed.2", ted, ted)
        oc = ch[0]
        for i in range(L):
            ic, oc = oc, ch[i]
            for j in range(self.num_res_blocks[i]):
                self._add_resnet(f"down_blocks.{i}.resnets.{j}", ic if j == 0 else oc, oc, ted)
                if self.attention_levels[i]:
                    self._add_attn(f"down_blocks.{i}.attentions.{j}", oc)
            if i != L - 1:
                self._add_conv(f"down_blocks.{i}.downsampler.op.conv", oc, oc, 3)
        self._add_resnet("middle_block.resnet_1", ch[-1], ch[-1], ted)
        self._add_attn("middle_block.attention", ch[-1])
        self._add_resnet("middle_block.resnet_2", ch[-1], ch[-1], ted)
        rc = list(reversed(ch))
        rr = list(reversed(self.num_res_blocks))
        ra = list(reversed(self.attention_levels))
        oc = rc[0]
        for i in range(L):
            prev, oc = oc, rc[i]
            ic = rc[min(i + 1, L - 1)]
            n = rr[i] + 1
            for j in range(n):
                res_skip = ic if j == n - 1 else oc
                res_in = prev if j == 0 else oc
                self._add_resnet(f"up_blocks.{i}.resnets.{j}", res_in + res_skip, oc, ted)
                if ra[i]:
                    self._add_attn(f"up_blocks.{i}.attentions.{j}", oc)
            if i != L - 1:
                self._add_conv(f"up_blocks.{i}.upsampler.conv.conv", oc, oc, 3)
        self._add_norm("out.0", ch[0])
        self._add_conv("out.2.conv", self.out_channels, ch[0], 3, zero=True)  # zero_module

    # ------------------------------------------------------------------ engine handle
    def _config_struct(self) -> _lib.UNetConfig:
        cfg = _lib.UNetConfig()
        cfg.spatial_dims = self.spatial_dims
        cfg.in_channels = self.in_channels
        cfg.out_channels = self.out_channels
        cfg.num_levels = len(self.block_out_channels)
        for i, c in enumerate(self.block_out_channels):
            cfg.num_channels[i] = c
            cfg.attention_levels[i] = int(self.attention_levels[i])
            cfg.num_res_blocks[i] = self.num_res_blocks[i]
            cfg.num_head_channels[i] = self.num_head_channels[i]
        cfg.norm_num_groups = self.norm_num_groups
        cfg.norm_eps = self.norm_eps
        return cfg

    def _release(self) -> None:
        if self._handle is not None:
            _lib.lib().ddpm_unet_destroy(self._handle)
            self._handle = None
            self._workspaces.clear()
            self._synced_versions = None

    def __del__(self):  # pragma: no cover - interpreter shutdown ordering
        try:
            self._release()
        except Exception:
            pass

    def _versions(self) -> Tuple[int, ...]:
        return tuple((p._version, p.data_ptr()) for p in self.parameters())

    def sync_weights(self, force: bool = False) -> None:
        """(Re)upload parameters into the engine's packed fp16/fp32 arenas if they changed."""
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise _lib.DdpmError("DiffusionModelUNet (B200 engine) needs its parameters on a CUDA device; "
                                 "there is no CPU fallback")
        L = _lib.lib()
        if self._handle is not None and self._handle_device != dev:
            self._release()
        with torch.cuda.device(dev):
            if self._handle is None:
                h = C.c_void_p()
                cfg = self._config_struct()
                _lib.check(L.ddpm_unet_create(C.byref(cfg), C.byref(h)), "ddpm_unet_create")
                self._handle = h
                self._handle_device = dev
                force = True
            ver = self._versions()
            if not force and ver == self._synced_versions:
                return
            stream = torch.cuda.current_stream().cuda_stream
            for name, p in self.named_parameters():
                t = p.detach()
                if t.dtype != torch.float32 or not t.is_contiguous():
                    t = t.float().contiguous()
                _lib.check(L.ddpm_unet_set_param(self._handle, name.encode(), t.data_ptr(), t.numel(), stream),
                           f"ddpm_unet_set_param({name})")
            _lib.check(L.ddpm_unet_finalize(self._handle, stream), "ddpm_unet_finalize")
            # temporaries created above (dtype casts) must outlive the async copies
            torch.cuda.current_stream().synchronize()
            self._synced_versions = ver

    def _workspace(self, n: int, d: int, h: int, w: int, device) -> torch.Tensor:
        # one workspace (and engine plan) per shape and per concurrent lane: two chains in flight on two streams must
        # not share activations (BatchReconstructor sets workspace_lane around its calls)
        key = (n, d, h, w, getattr(self, "workspace_lane", 0))
        ws = self._workspaces.get(key)
        if ws is None:
            need = _lib.lib().ddpm_unet_workspace_bytes(self._handle, n, d, h, w)
            if need <= 0:
                _lib.check(1, "ddpm_unet_workspace_bytes")
            ws = torch.empty(need, dtype=torch.uint8, device=device)
            self._workspaces[key] = ws
        return ws

    def _dims(self, x: torch.Tensor) -> Tuple[int, int, int, int]:
        if x.dim() != self.spatial_dims + 2:
            raise ValueError(f"expected a {self.spatial_dims + 2}-D input, got {tuple(x.shape)}")
        if x.shape[1] != self.in_channels:
            raise ValueError(f"expected {self.in_channels} input channels, got {x.shape[1]}")
        if self.spatial_dims == 2:
            return x.shape[0], 1, x.shape[2], x.shape[3]
        return x.shape[0], x.shape[2], x.shape[3], x.shape[4]

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, x: torch.Tensor, timesteps: torch.Tensor, context: Optional[torch.Tensor] = None,
                class_labels: Optional[torch.Tensor] = None) -> torch.Tensor:
        if context is not None or class_labels is not None:
            raise NotImplementedError("conditioning is not on the reconstruction hot path")
        if not x.is_cuda:
            raise _lib.DdpmError("DiffusionModelUNet.forward needs a CUDA tensor; there is no CPU fallback")
        self.sync_weights()
        n, d, h, w = self._dims(x)
        xx = x.detach().float().contiguous()
        ts = timesteps.to(device=x.device, dtype=torch.int64).contiguous()
        if ts.numel() != n:
            raise ValueError("timesteps must have one entry per batch item")
        out_shape = (n, self.out_channels) + tuple(x.shape[2:])
        out = torch.empty(out_shape, dtype=torch.float32, device=x.device)
        if n == 0:  # an empty batch is an empty result (what torch modules return), not a launch
            return out
        with torch.cuda.device(x.device):
            ws = self._workspace(n, d, h, w, x.device)
            _lib.check(
                _lib.lib().ddpm_unet_forward(self._handle, xx.data_ptr(), ts.data_ptr(), out.data_ptr(), n, d, h, w,
                                             ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream),
                "ddpm_unet_forward")
        return out

    # ------------------------------------------------------------------ fused chain (used by the Reconstruct trainer)
    @torch.no_grad()
    def run_chain(self, sample: torch.Tensor, timesteps: Sequence[int], steps, ring: torch.Tensor,
                  stash: torch.Tensor) -> None:
        """In-place: for each timestep, eps = UNet(sample, t); sample = PLMS(eps). `steps` is a ctypes array of
        ddpm_plms_step prepared by PNDMScheduler.plan_chain()."""
        self.sync_weights()
        n, d, h, w = self._dims(sample)
        assert sample.dtype == torch.float32 and sample.is_contiguous()
        ts = (C.c_int * len(timesteps))(*[int(t) for t in timesteps])
        with torch.cuda.device(sample.device):
            ws = self._workspace(n, d, h, w, sample.device)
            _lib.check(
                _lib.lib().ddpm_unet_run_chain(self._handle, len(timesteps), ts, steps, sample.data_ptr(),
                                               ring.data_ptr(), stash.data_ptr(), n, d, h, w, ws.data_ptr(), ws.numel(),
                                               torch.cuda.current_stream().cuda_stream),
                "ddpm_unet_run_chain")

    def set_profile(self, every: int) -> None:
        """Time every `every`-th forward per op type with CUDA events (0 = off); see read_profile()."""
        self.sync_weights()
        _lib.check(_lib.lib().ddpm_unet_set_profile(self._handle, int(every)), "ddpm_unet_set_profile")

    def read_profile(self, reset: bool = True) -> Dict[str, Dict[str, float]]:
        p = _lib.OpProfile()
        _lib.check(_lib.lib().ddpm_unet_read_profile(self._handle, C.byref(p), int(reset)), "ddpm_unet_read_profile")
        out: Dict[str, Dict[str, float]] = {"_total": {"forwards": int(p.forwards), "ms": float(p.forward_ms)}}
        for i, name in enumerate(_lib.OP_TYPE_NAMES):
            if p.launches[i]:
                out[name] = {"ms": float(p.ms[i]), "flops": float(p.flops[i]), "bytes": float(p.bytes[i]),
                             "launches": int(p.launches[i])}
        return out

    def launch_count(self) -> int:
        return int(_lib.lib().ddpm_unet_launch_count(self._handle)) if self._handle is not None else 0
