"""Drop-in `PerceptualLoss` (src/losses/perceptual_loss.py) whose LPIPS runs in libddpm_ood_b200.so.

Same constructor and call form as the reference wrapper (`PerceptualLoss(dimensions, include_pixel_loss, is_fake_3d,
lpips_normalize, spatial)`, `pl(y, y_pred)`; src/trainers/reconstruct.py:82-88,172-187). Parameter names follow
`lpips.LPIPS.state_dict()` under the `perceptual_function.` prefix, so real LPIPS weights load with `load_state_dict`.

Weights: the `lpips` package (which bundles the linear heads and pulls torchvision's pretrained AlexNet) is used when it
is importable, or a saved state dict (lpips_kwargs["model_path"] / DDPM_LPIPS_STATE_DICT). Without either the constructor
RAISES: a seeded synthetic initialisation exists only behind an explicit opt-in (allow_synthetic_weights=True or
DDPM_LPIPS_ALLOW_SYNTHETIC=1) for tests, bench.py and smoke(), whose scores are self-consistent but not comparable with
the reference's.
"""
from __future__ import annotations

import ctypes as C
import os
import warnings
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib

_CONVS = (("net.slice1.0", 64, 3, 11), ("net.slice2.3", 192, 64, 5), ("net.slice3.6", 384, 192, 3),
          ("net.slice4.8", 256, 384, 3), ("net.slice5.10", 256, 256, 3))


class _LPIPSAlex(nn.Module):
    """Parameter container + engine handle for LPIPS(net='alex', version='0.1', lpips=True, spatial=False)."""

    def __init__(self, seed: int = 1234):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        for name, cout, cin, k in _CONVS:
            w = torch.randn((cout, cin, k, k), generator=g) * (2.0 / (cin * k * k)) ** 0.5
            b = 0.05 * torch.randn((cout,), generator=g)
            self._add(name + ".weight", w)
            self._add(name + ".bias", b)
        for i, (_, cout, _, _) in enumerate(_CONVS):
            self._add(f"lin{i}.model.1.weight", torch.rand((1, cout, 1, 1), generator=g) * (2.0 / cout))
        self.register_buffer("shift", torch.tensor([-0.030, -0.088, -0.188])[None, :, None, None], persistent=False)
        self.register_buffer("scale", torch.tensor([0.458, 0.448, 0.450])[None, :, None, None], persistent=False)
        self.synthetic = True
        self._handle: Optional[C.c_void_p] = None
        self._handle_device = None
        self._synced = None
        self._ws: Dict[Tuple[int, int, int], torch.Tensor] = {}

    def _add(self, path: str, tensor: torch.Tensor) -> None:
        parts = path.split(".")
        mod: nn.Module = self
        for p in parts[:-1]:
            if p not in mod._modules:
                mod.add_module(p, nn.Module())
            mod = mod._modules[p]
        mod.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=False))

    def load_lpips_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        """Accepts lpips.LPIPS(...).state_dict() (extra keys such as lins.* / scaling_layer.* are ignored)."""
        own = dict(self.named_parameters())
        missing = [k for k in own if k not in sd]
        if missing:
            raise KeyError(f"LPIPS state dict lacks {missing}")
        with torch.no_grad():
            for k, p in own.items():
                p.copy_(sd[k].reshape(p.shape))
        self.synthetic = False

    def _release(self):
        if self._handle is not None:
            _lib.lib().ddpm_lpips_destroy(self._handle)
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
            raise _lib.DdpmError("PerceptualLoss (B200 engine) needs .to('cuda'); there is no CPU fallback")
        L = _lib.lib()
        if self._handle is not None and self._handle_device != dev:
            self._release()
        ver = tuple((p._version, p.data_ptr()) for p in self.parameters())
        if self._handle is not None and ver == self._synced:
            return
        with torch.cuda.device(dev):
            if self._handle is None:
                h = C.c_void_p()
                _lib.check(L.ddpm_lpips_create(C.byref(h)), "ddpm_lpips_create")
                self._handle = h
                self._handle_device = dev
            stream = torch.cuda.current_stream().cuda_stream
            for name, p in self.named_parameters():
                t = p.detach().float().contiguous()
                _lib.check(L.ddpm_lpips_set_param(self._handle, name.encode(), t.data_ptr(), t.numel(), stream),
                           f"ddpm_lpips_set_param({name})")
            _lib.check(L.ddpm_lpips_finalize(self._handle), "ddpm_lpips_finalize")
            torch.cuda.current_stream().synchronize()
        self._synced = ver

    @torch.no_grad()
    def forward(self, in0: torch.Tensor, in1: torch.Tensor, normalize: bool = False) -> torch.Tensor:
        if not in0.is_cuda or not in1.is_cuda:
            raise _lib.DdpmError("LPIPS needs CUDA tensors; there is no CPU fallback")
        if in0.shape != in1.shape or in0.dim() != 4:
            raise ValueError(f"LPIPS expects two [B, C, H, W] tensors of equal shape, got {tuple(in0.shape)} "
                             f"and {tuple(in1.shape)}")
        self._sync()
        a = in0.detach().float().contiguous()
        b = in1.detach().float().contiguous()
        B, Cc, H, W = a.shape
        out = torch.empty((B,), dtype=torch.float32, device=a.device)
        L = _lib.lib()
        with torch.cuda.device(a.device):
            key = (B, H, W)
            ws = self._ws.get(key)
            if ws is None:
                need = L.ddpm_lpips_workspace_bytes(self._handle, B, H, W)
                if need <= 0:
                    raise _lib.DdpmError(f"LPIPS: {H}x{W} input is too small for AlexNet")
                ws = torch.empty(need, dtype=torch.uint8, device=a.device)
                self._ws[key] = ws
            _lib.check(L.ddpm_lpips_forward(self._handle, a.data_ptr(), b.data_ptr(), out.data_ptr(), B, Cc, H, W,
                                            1 if normalize else 0, ws.data_ptr(), ws.numel(),
                                            torch.cuda.current_stream().cuda_stream), "ddpm_lpips_forward")
        return out.view(B, 1, 1, 1)

    def launch_count(self) -> int:
        return int(_lib.lib().ddpm_lpips_launch_count(self._handle)) if self._handle is not None else 0


def _try_real_lpips_weights(model: _LPIPSAlex, lpips_kwargs: Dict) -> bool:
    path = lpips_kwargs.get("model_path") or os.environ.get("DDPM_LPIPS_STATE_DICT")
    if path and os.path.exists(path):
        model.load_lpips_state_dict(torch.load(path, map_location="cpu"))
        return True
    try:
        import lpips  # noqa: F401  (third-party; absent in the build container)
    except Exception:
        return False
    kw = dict(lpips_kwargs)
    kw["verbose"] = False
    ref = lpips.LPIPS(**kw)
    model.load_lpips_state_dict(ref.state_dict())
    return True


class PerceptualLoss(nn.Module):
    def __init__(self, dimensions: int, include_pixel_loss: bool = True, is_fake_3d: bool = True,
                 drop_ratio: float = 0.0, fake_3d_axis: Tuple[int, ...] = (2, 3, 4), lpips_kwargs: Dict = None,
                 lpips_normalize: bool = True, spatial: bool = False, allow_synthetic_weights: bool = False):
        super().__init__()
        if dimensions not in (2, 3):
            raise NotImplementedError("Perceptual loss is implemented only in 2D and 3D.")
        if dimensions == 3 and is_fake_3d is False:
            raise NotImplementedError("True 3D perceptual loss is not implemented yet.")
        if spatial:
            raise NotImplementedError("spatial=True is not used on the reconstruction path (trainers/reconstruct.py:87)")
        self.dimensions = dimensions
        self.include_pixel_loss = include_pixel_loss
        self.lpips_kwargs = (
            {"pretrained": True, "net": "alex", "version": "0.1", "lpips": True, "spatial": spatial,
             "pnet_rand": False, "pnet_tune": False, "use_dropout": True, "model_path": None, "eval_mode": True,
             "verbose": False} if lpips_kwargs is None else lpips_kwargs)
        if self.lpips_kwargs.get("net", "alex") != "alex":
            raise NotImplementedError("only net='alex' is implemented (the reference default)")
        self.fake_3D_views = (
            ([((0, 2, 1, 3, 4), (1, 3, 4))] if 2 in fake_3d_axis else [])
            + ([((0, 3, 1, 2, 4), (1, 2, 4))] if 3 in fake_3d_axis else [])
            + ([((0, 4, 1, 2, 3), (1, 2, 3))] if 4 in fake_3d_axis else [])
        ) if is_fake_3d else None
        self.keep_ratio = 1 - drop_ratio
        self.lpips_normalize = lpips_normalize
        self.perceptual_function = _LPIPSAlex()
        if not _try_real_lpips_weights(self.perceptual_function, self.lpips_kwargs):
            # Production (trainers.Reconstruct, the CLI) must not write `perceptual_difference` columns computed with
            # made-up weights into reference-shaped CSVs: refuse unless the caller opts in (tests, bench, smoke - which
            # then load the oracle's weights or only time the kernels).
            if not (allow_synthetic_weights or os.environ.get("DDPM_LPIPS_ALLOW_SYNTHETIC") == "1"):
                raise _lib.DdpmError(
                    "LPIPS weights not found: install the `lpips` package, or point lpips_kwargs['model_path'] / "
                    "DDPM_LPIPS_STATE_DICT at a saved lpips.LPIPS(net='alex').state_dict(). A seeded synthetic "
                    "initialisation exists for tests and benchmarks only (allow_synthetic_weights=True or "
                    "DDPM_LPIPS_ALLOW_SYNTHETIC=1); its scores are not comparable with the reference's.")
            warnings.warn("LPIPS: synthetic AlexNet/linear-head weights (explicit opt-in); perceptual_difference values "
                          "are self-consistent but not comparable with the reference's.")
        self.perceptual_factor = 1

    @torch.no_grad()
    def forward(self, y: torch.Tensor, y_pred: torch.Tensor) -> torch.Tensor:
        y = y.float()
        y_pred = y_pred.float()
        if self.dimensions == 3 and self.fake_3D_views:
            # The reference loop assigns (not accumulates) per view (perceptual_loss.py:113-122): the last view is the
            # result, so only that one is evaluated.
            permute_dims, view_dims = self.fake_3D_views[-1]
            loss = self._calculate_fake_3d_loss(y, y_pred, permute_dims, view_dims) * self.perceptual_factor
        else:
            loss = self.perceptual_function(y, y_pred, normalize=self.lpips_normalize) * self.perceptual_factor
        return loss

    def _calculate_fake_3d_loss(self, y, y_pred, permute_dims, view_dims):
        ys = y.permute(*permute_dims).contiguous().view(-1, y.shape[view_dims[0]], y.shape[view_dims[1]],
                                                        y.shape[view_dims[2]])
        ps = y_pred.permute(*permute_dims).contiguous().view(-1, y_pred.shape[view_dims[0]],
                                                             y_pred.shape[view_dims[1]], y_pred.shape[view_dims[2]])
        if self.keep_ratio < 1:
            idx = torch.randperm(ps.shape[0], device=ps.device)[: int(ps.shape[0] * self.keep_ratio)]
            ys, ps = ys[idx], ps[idx]
        return torch.mean(self.perceptual_function(ys, ps, normalize=self.lpips_normalize))

    @torch.no_grad()
    def per_item(self, y: torch.Tensor, y_pred: torch.Tensor, max_slices: int = 4096) -> torch.Tensor:
        """[B] scores, one per batch item: what the reference's 3-D scoring loop computes by calling the loss once per
        item (`self.perceptual_loss(images_original[b, None, ...], reconstructions[b, None, ...])`,
        src/trainers/reconstruct.py:181-187), with the slices of several items in ONE LPIPS launch sequence
        (`max_slices` slice pairs per call) and the mean taken per item."""
        B = y.shape[0]
        if not (self.dimensions == 3 and self.fake_3D_views and self.keep_ratio == 1):
            return torch.stack([self(y[b:b + 1], y_pred[b:b + 1]).reshape(()) for b in range(B)])
        permute_dims, view_dims = self.fake_3D_views[-1]  # the reference keeps the last view's value
        y, y_pred = y.float(), y_pred.float()
        slices = y.shape[permute_dims[1]]
        step = max(1, max_slices // slices)
        out = []
        for b0 in range(0, B, step):
            ys = y[b0:b0 + step].permute(*permute_dims).contiguous()
            ps = y_pred[b0:b0 + step].permute(*permute_dims).contiguous()
            n = ys.shape[0]
            ys = ys.view(-1, ys.shape[2], ys.shape[3], ys.shape[4])
            ps = ps.view(-1, ps.shape[2], ps.shape[3], ps.shape[4])
            d = self.perceptual_function(ys, ps, normalize=self.lpips_normalize)
            out.append(d.reshape(n, slices).mean(dim=1) * self.perceptual_factor)
        return torch.cat(out)

    def get_perceptual_factor(self) -> float:
        return self.perceptual_factor

    def set_perceptual_factor(self, perceptual_factor: float) -> float:
        self.perceptual_factor = perceptual_factor
        return self.get_perceptual_factor()
