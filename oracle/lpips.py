"""ORACLE (test infrastructure, not product code): restatement of `lpips.LPIPS(net='alex', version='0.1', lpips=True,
spatial=False)` (lpips==0.1.4, requirements.txt:1) and of the reference's wrapper src/losses/perceptual_loss.py:105-186.

PARITY UNPINNED for the third-party part: the `lpips` package and its pretrained AlexNet / linear-head weights are not
available; weights here are synthetic and shared with the CUDA path. The wrapper part
(PerceptualLoss) follows the reference file line by line, including the 2.5-D loop that overwrites instead of
accumulating (perceptual_loss.py:113-122), so only the last view counts; it IS pinned: the reference's own
perceptual_loss.py, executed with this file's LPIPS in place of the absent package, gives the same values in 2-D and
2.5-D (tests/test_reference_loop_pin_cpu.py).

Parameter names mirror lpips' (`net.slice{1..5}.{idx}.weight`, `lin{k}.model.1.weight`) so real weights would load.
"""
from __future__ import annotations

import torch
import torch.nn as nn

ALEX_CHNS = (64, 192, 384, 256, 256)


class ScalingLayer(nn.Module):
    def __init__(self):
        super().__init__()
        self.register_buffer("shift", torch.tensor([-0.030, -0.088, -0.188])[None, :, None, None])
        self.register_buffer("scale", torch.tensor([0.458, 0.448, 0.450])[None, :, None, None])

    def forward(self, x):
        return (x - self.shift) / self.scale


class AlexSlices(nn.Module):
    """torchvision AlexNet `features`, cut after ReLU 1..5 (indices 1, 4, 7, 9, 11)."""

    def __init__(self):
        super().__init__()
        self.slice1 = nn.Sequential()
        self.slice2 = nn.Sequential()
        self.slice3 = nn.Sequential()
        self.slice4 = nn.Sequential()
        self.slice5 = nn.Sequential()
        self.slice1.add_module("0", nn.Conv2d(3, 64, kernel_size=11, stride=4, padding=2))
        self.slice1.add_module("1", nn.ReLU(inplace=False))
        self.slice2.add_module("2", nn.MaxPool2d(kernel_size=3, stride=2))
        self.slice2.add_module("3", nn.Conv2d(64, 192, kernel_size=5, padding=2))
        self.slice2.add_module("4", nn.ReLU(inplace=False))
        self.slice3.add_module("5", nn.MaxPool2d(kernel_size=3, stride=2))
        self.slice3.add_module("6", nn.Conv2d(192, 384, kernel_size=3, padding=1))
        self.slice3.add_module("7", nn.ReLU(inplace=False))
        self.slice4.add_module("8", nn.Conv2d(384, 256, kernel_size=3, padding=1))
        self.slice4.add_module("9", nn.ReLU(inplace=False))
        self.slice5.add_module("10", nn.Conv2d(256, 256, kernel_size=3, padding=1))
        self.slice5.add_module("11", nn.ReLU(inplace=False))

    def forward(self, x):
        h1 = self.slice1(x)
        h2 = self.slice2(h1)
        h3 = self.slice3(h2)
        h4 = self.slice4(h3)
        h5 = self.slice5(h4)
        return [h1, h2, h3, h4, h5]


class NetLinLayer(nn.Module):
    def __init__(self, chn_in):
        super().__init__()
        self.model = nn.Sequential(nn.Dropout(), nn.Conv2d(chn_in, 1, 1, stride=1, padding=0, bias=False))

    def forward(self, x):
        return self.model(x)


def normalize_tensor(f, eps=1e-10):
    return f / (torch.sqrt(torch.sum(f ** 2, dim=1, keepdim=True)) + eps)


class LPIPS(nn.Module):
    def __init__(self, seed: int = 1234):
        super().__init__()
        self.scaling_layer = ScalingLayer()
        self.net = AlexSlices()
        self.lin0, self.lin1, self.lin2, self.lin3, self.lin4 = (NetLinLayer(c) for c in ALEX_CHNS)
        self.lins = nn.ModuleList([self.lin0, self.lin1, self.lin2, self.lin3, self.lin4])
        self.synthetic_init_(seed)
        self.eval()

    def synthetic_init_(self, seed: int) -> None:
        """No pretrained weights exist in this environment: seeded He-style convs, non-negative linear heads."""
        g = torch.Generator().manual_seed(seed)
        with torch.no_grad():
            for name, p in sorted(self.net.named_parameters()):
                if p.dim() == 4:
                    fan_in = p[0].numel()
                    p.copy_(torch.randn(p.shape, generator=g) * (2.0 / fan_in) ** 0.5)
                else:
                    p.copy_(0.05 * torch.randn(p.shape, generator=g))
            for lin in self.lins:
                w = lin.model[1].weight
                w.copy_(torch.rand(w.shape, generator=g) * (2.0 / w.shape[1]))

    def forward(self, in0, in1, normalize=False):
        if normalize:
            in0 = 2 * in0 - 1
            in1 = 2 * in1 - 1
        o0 = self.net(self.scaling_layer(in0))
        o1 = self.net(self.scaling_layer(in1))
        val = 0
        for k in range(5):
            d = (normalize_tensor(o0[k]) - normalize_tensor(o1[k])) ** 2
            val = val + self.lins[k](d).mean([2, 3], keepdim=True)
        return val


class PerceptualLoss(nn.Module):
    """src/losses/perceptual_loss.py, restated for include_pixel_loss=False, drop_ratio=0."""

    def __init__(self, dimensions: int, include_pixel_loss: bool = True, is_fake_3d: bool = True,
                 drop_ratio: float = 0.0, fake_3d_axis=(2, 3, 4), lpips_kwargs=None, lpips_normalize: bool = True,
                 spatial: bool = False, seed: int = 1234):
        super().__init__()
        if dimensions not in (2, 3):
            raise NotImplementedError("Perceptual loss is implemented only in 2D and 3D.")
        if dimensions == 3 and is_fake_3d is False:
            raise NotImplementedError("True 3D perceptual loss is not implemented yet.")
        self.dimensions = dimensions
        self.fake_3D_views = (
            ([((0, 2, 1, 3, 4), (1, 3, 4))] if 2 in fake_3d_axis else [])
            + ([((0, 3, 1, 2, 4), (1, 2, 4))] if 3 in fake_3d_axis else [])
            + ([((0, 4, 1, 2, 3), (1, 2, 3))] if 4 in fake_3d_axis else [])
        ) if is_fake_3d else None
        self.keep_ratio = 1 - drop_ratio
        self.lpips_normalize = lpips_normalize
        self.perceptual_function = LPIPS(seed=seed)
        self.perceptual_factor = 1

    def forward(self, y, y_pred):
        y = y.float()
        y_pred = y_pred.float()
        if self.dimensions == 3 and self.fake_3D_views:
            loss = torch.zeros(())
            for permute_dims, view_dims in self.fake_3D_views:  # overwrites: last view wins (reference :113-122)
                loss = self._calculate_fake_3d_loss(y, y_pred, permute_dims, view_dims) * self.perceptual_factor
        else:
            loss = self.perceptual_function(y, y_pred, normalize=self.lpips_normalize) * self.perceptual_factor
        return loss

    def _calculate_fake_3d_loss(self, y, y_pred, permute_dims, view_dims):
        ys = y.permute(*permute_dims).contiguous().view(-1, y.shape[view_dims[0]], y.shape[view_dims[1]],
                                                        y.shape[view_dims[2]])
        ps = y_pred.permute(*permute_dims).contiguous().view(-1, y_pred.shape[view_dims[0]],
                                                             y_pred.shape[view_dims[1]], y_pred.shape[view_dims[2]])
        # keep_ratio == 1: the random permutation (reference :171-177) does not change the mean.
        return torch.mean(self.perceptual_function(ys, ps, normalize=self.lpips_normalize))
