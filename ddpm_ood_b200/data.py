"""Minimal host-side data ingest for the reconstruction CLI.

The reference's loader (src/data/get_train_and_val_dataloader.py) is MONAI-based host I/O and is OUT OF SCOPE for the
accelerated path; this module only restates enough of it for `reconstruct.py` to run end to end
on the reference's own `.npy` datasets (src/data/get_computer_vision_datasets.py writes [H,W] uint8 for grayscale and
[3,H,W] uint8 for colour) and on single-file NIfTI-1 volumes (the 3-D medical datasets; `nifti.py`), and to reproduce its multi-GPU image partition (get_train_and_val_dataloader.py:21-31).
Batches have the reference's structure: {"image": Tensor[B,C,...], "image_meta_dict": {"filename_or_obj": [...]}}.
"""
from __future__ import annotations

import math
from typing import Dict, Iterator, List, Optional, Sequence

import numpy as np
import pandas as pd
import torch
import torch.nn.functional as F


def partition_indices(data_len: int, num_partitions: int, rank: int, shuffle: bool = True, seed: int = 0,
                      drop_last: bool = False, even_divisible: bool = True) -> List[int]:
    """Restatement of monai.data.partition_dataset(num_partitions=..., shuffle=True, seed=0, drop_last=False,
    even_divisible=True)[rank] as called at get_train_and_val_dataloader.py:24-31 [3P-RECALL]: seeded shuffle, wrap-pad
    to a multiple of the world size, then a strided slice per rank."""
    if data_len < num_partitions:
        raise RuntimeError(f"there is no enough data to be split into {num_partitions} partitions.")
    indices = list(range(data_len))
    if shuffle:
        np.random.RandomState(seed).shuffle(indices)
    if drop_last and data_len % num_partitions != 0:
        num_samples = math.ceil((data_len - num_partitions) / num_partitions)
    else:
        num_samples = math.ceil(data_len / num_partitions)
    total_size = num_samples * num_partitions
    if even_divisible:
        if not drop_last and total_size - data_len > 0:
            indices += indices[: (total_size - data_len)]
        else:
            indices = indices[:total_size]
    return indices[rank:total_size:num_partitions]


def get_data_dicts(ids_path: str, first_n=False, rank: Optional[int] = None, world_size: Optional[int] = None):
    # the split files are ONE csv row of paths; pandas reads it as the header (get_train_and_val_dataloader.py:10-17)
    paths = list(pd.read_csv(ids_path, sep=","))
    dicts = [{"image": p} for p in paths]
    if first_n is not False and first_n is not None:
        dicts = dicts[: int(first_n)]
    print(f"Found {len(dicts)} subjects.")
    if world_size is not None and world_size > 1:
        idx = partition_indices(len(dicts), world_size, rank)
        dicts = [dicts[i] for i in idx]
    return dicts


def _load(path: str) -> np.ndarray:
    if path.endswith(".npy"):
        return np.load(path)
    if path.endswith(".pt"):
        return torch.load(path, map_location="cpu").numpy()
    if path.endswith(".nii") or path.endswith(".nii.gz"):
        # what monai's LoadImaged + EnsureChannelFirstd give for a NIfTI volume (get_train_and_val_dataloader.py:68-69):
        # the file's own axis order, a 4th (modality) axis - BraTS stores four modalities in one file - moved first
        from .nifti import read_nifti

        arr = read_nifti(path)
        return np.moveaxis(arr, -1, 0) if arr.ndim == 4 else arr
    raise NotImplementedError(f"the minimal loader reads .npy / .pt / .nii / .nii.gz images (got {path})")


def _transform(arr: np.ndarray, is_grayscale: bool, spatial_dimension: int, image_size, image_roi, add_vflip: bool,
               add_hflip: bool) -> torch.Tensor:
    x = _select(torch.from_numpy(np.ascontiguousarray(arr)), is_grayscale, spatial_dimension, image_roi).float()
    if image_size:
        size = (int(image_size),) * spatial_dimension
        x = F.interpolate(x[None], size=size, mode="area")[0]  # monai Resize default mode
    mn, mx = x.min(), x.max()
    x = (x - mn) / (mx - mn) if mx > mn else x - mn  # ScaleIntensityd(minv=0, maxv=1)
    if add_vflip:
        x = torch.flip(x, dims=(1,))
    if add_hflip:
        x = torch.flip(x, dims=(2,))
    return x.contiguous()


def _select(x: torch.Tensor, is_grayscale: bool, spatial_dimension: int, image_roi) -> torch.Tensor:
    """The index-only head of the transform chain (channel selection + centre crop), in the stored dtype."""
    if is_grayscale:
        if x.dim() == spatial_dimension:  # EnsureChannelFirstd
            x = x[None]
        x = x[0, None, ...]  # "needed for BRATs data with 4 modalities in 1"
    if image_roi:
        roi = [int(r) for r in image_roi]
        sl = [slice(None)]
        for d, r in enumerate(roi):
            size = x.shape[1 + d]
            if r < 0 or r >= size:
                sl.append(slice(None))
            else:
                start = (size - r) // 2
                sl.append(slice(start, start + r))
        x = x[tuple(sl)]
    return x


def scale_intensity_on_device(raw: torch.Tensor) -> torch.Tensor:
    """Per-image min-max scaling to [0, 1] (ScaleIntensityd(minv=0, maxv=1), get_train_and_val_dataloader.py:76) of a
    batch of raw images already on the GPU, uint8 or float32 [N, ...]: the GPU-resident ingest path."""
    from . import _lib

    if not raw.is_cuda or raw.dtype not in (torch.uint8, torch.float32):
        raise _lib.DdpmError("scale_intensity_on_device needs a CUDA uint8 / float32 tensor; there is no CPU fallback")
    raw = raw.contiguous()
    n = raw.shape[0]
    out = torch.empty(raw.shape, dtype=torch.float32, device=raw.device)
    with torch.cuda.device(raw.device):
        _lib.check(_lib.lib().ddpm_scale_intensity(raw.data_ptr(), int(raw.dtype == torch.uint8), out.data_ptr(), n,
                                                   raw.numel() // max(n, 1), torch.cuda.current_stream().cuda_stream),
                   "ddpm_scale_intensity")
    return out


class SimpleLoader:
    """Sequential, un-shuffled batches (the reference's val loader: ThreadDataLoader(shuffle=False))."""

    def __init__(self, dicts: Sequence[Dict[str, str]], batch_size: int, drop_last: bool, device=None, **tf):
        """device: a CUDA device turns on GPU-resident ingest: images are copied as stored (uint8 stays
        uint8: a quarter of the PCIe bytes) in one transfer per batch and scaled per image by ddpm_scale_intensity;
        batches come out on that device. Not available with image_size (area resize happens before the scaling)."""
        self.device = torch.device(device) if device is not None else None
        if self.device is not None and (self.device.type != "cuda" or tf.get("image_size")):
            raise ValueError("device ingest needs a CUDA device and no image_size resize")
        self.dicts = list(dicts)
        self.batch_size = batch_size
        self.drop_last = drop_last
        self.tf = tf
        self._cache: Dict[int, torch.Tensor] = {}

    def __len__(self) -> int:
        n = len(self.dicts)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self) -> Iterator[Dict]:
        n = len(self.dicts)
        for s in range(0, n, self.batch_size):
            idx = list(range(s, min(s + self.batch_size, n)))
            if self.drop_last and len(idx) < self.batch_size:
                return
            names = {"filename_or_obj": [self.dicts[i]["image"] for i in idx]}
            if self.device is not None:
                yield {"image": self._device_batch(idx), "image_meta_dict": names}
                continue
            imgs = []
            for i in idx:
                if i not in self._cache:
                    self._cache[i] = _transform(_load(self.dicts[i]["image"]), **self.tf)
                imgs.append(self._cache[i])
            yield {"image": torch.stack(imgs), "image_meta_dict": names}

    def _device_batch(self, idx) -> torch.Tensor:
        tf = self.tf
        raws = []
        for i in idx:
            if i not in self._cache:
                x = _select(torch.from_numpy(np.ascontiguousarray(_load(self.dicts[i]["image"]))), tf["is_grayscale"],
                            tf["spatial_dimension"], tf["image_roi"])
                self._cache[i] = x.contiguous() if x.dtype == torch.uint8 else x.float().contiguous()
            raws.append(self._cache[i])
        if len({r.dtype for r in raws}) > 1:
            raws = [r.float() for r in raws]
        x = scale_intensity_on_device(torch.stack(raws).pin_memory().to(self.device, non_blocking=True))
        flips = [d for d, on in ((2, tf["add_vflip"]), (3, tf["add_hflip"])) if on]
        return torch.flip(x, dims=flips) if flips else x


def get_training_data_loader(batch_size: int, training_ids: str, validation_ids: str, only_val: bool = False,
                             augmentation: bool = True, drop_last: bool = False, num_workers: int = 8,
                             num_val_workers: int = 3, cache_data=True, first_n=None, is_grayscale=False,
                             add_vflip=False, add_hflip=False, image_size=None, image_roi=None, spatial_dimension=2,
                             rank: Optional[int] = None, world_size: Optional[int] = None, device=None):
    """Same signature as the reference's get_training_data_loader (get_train_and_val_dataloader.py:36-53); only the
    validation-style loader (only_val=True) is on the reconstruction path."""
    if not only_val:
        raise NotImplementedError("training loaders are out of scope for the reconstruction path")
    dicts = get_data_dicts(validation_ids, first_n=first_n if first_n else False, rank=rank, world_size=world_size)
    return SimpleLoader(dicts, batch_size, bool(drop_last), device=device, is_grayscale=bool(is_grayscale),
                        spatial_dimension=spatial_dimension, image_size=image_size, image_roi=image_roi,
                        add_vflip=add_vflip, add_hflip=add_hflip)
