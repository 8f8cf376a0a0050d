"""Minimal NIfTI-1 reader for the reconstruction CLI's 3-D datasets.

The reference loads volumes with `monai.transforms.LoadImaged` (src/data/get_train_and_val_dataloader.py:68), which
for `.nii` / `.nii.gz` goes through nibabel (`monai[nibabel]==1.2.0`, requirements.txt:4) - absent from this
environment. This module reads what that route returns for a single-file NIfTI-1 image [3P-RECALL: nibabel's
`np.asanyarray(img.dataobj)` as used by monai's NibabelReader, no reorientation]:

* the voxel array in the file's own axis order - index [i, j, k(, t)], i fastest on disk (Fortran order);
* `scl_slope` / `scl_inter` applied when the header sets them (slope != 0 and not the identity pair);
* as float32 (LoadImage's default dtype).

Out of scope: NIfTI-2, the two-file `.hdr` / `.img` form, header extensions' content, affines (the path never reads them).
"""
from __future__ import annotations

import gzip
import struct

import numpy as np

_DTYPES = {2: "u1", 4: "i2", 8: "i4", 16: "f4", 64: "f8", 256: "i1", 512: "u2", 768: "u4", 1024: "i8", 1280: "u8"}


class NiftiError(ValueError):
    pass


def read_nifti(path: str) -> np.ndarray:
    """Voxel data of a single-file NIfTI-1 image (`.nii` or `.nii.gz`) as float32, shape `dim[1 : ndim + 1]`."""
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rb") as f:
        raw = f.read()
    if len(raw) < 348:
        raise NiftiError(f"{path}: shorter than a NIfTI-1 header")
    # endianness: sizeof_hdr must read 348
    if struct.unpack_from("<i", raw, 0)[0] == 348:
        bo = "<"
    elif struct.unpack_from(">i", raw, 0)[0] == 348:
        bo = ">"
    elif struct.unpack_from("<i", raw, 0)[0] == 540 or struct.unpack_from(">i", raw, 0)[0] == 540:
        raise NiftiError(f"{path}: NIfTI-2 is not supported by the minimal reader")
    else:
        raise NiftiError(f"{path}: not a NIfTI-1 file (sizeof_hdr != 348)")
    magic = raw[344:348]
    if magic == b"ni1\x00":
        raise NiftiError(f"{path}: two-file NIfTI (.hdr/.img) is not supported by the minimal reader")
    if magic != b"n+1\x00":
        raise NiftiError(f"{path}: bad NIfTI-1 magic {magic!r}")
    dim = struct.unpack_from(bo + "8h", raw, 40)
    ndim = dim[0]
    if not 1 <= ndim <= 7:
        raise NiftiError(f"{path}: dim[0] = {ndim}")
    shape = tuple(int(d) for d in dim[1:ndim + 1])
    if any(d < 1 for d in shape):
        raise NiftiError(f"{path}: non-positive extent in dim {dim}")
    # Trailing singleton dims beyond the three spatial ones carry no data. monai keeps them as a channel of size one and
    # EnsureChannelFirstd moves it to the front; dropping them here gives the loader's `x[None]` the same [1, X, Y, Z].
    while len(shape) > 3 and shape[-1] == 1:
        shape = shape[:-1]
    datatype, bitpix = struct.unpack_from(bo + "2h", raw, 70)
    if datatype not in _DTYPES:
        raise NiftiError(f"{path}: unsupported datatype code {datatype}")
    dt = np.dtype(bo + _DTYPES[datatype])
    if dt.itemsize * 8 != bitpix:
        raise NiftiError(f"{path}: bitpix {bitpix} does not match datatype {datatype}")
    vox_offset = int(struct.unpack_from(bo + "f", raw, 108)[0])
    if vox_offset < 352:
        vox_offset = 352  # single-file images start their data after the 4-byte extension flag at the earliest
    slope, inter = struct.unpack_from(bo + "2f", raw, 112)
    count = int(np.prod(shape))
    if len(raw) < vox_offset + count * dt.itemsize:
        raise NiftiError(f"{path}: file holds fewer than {count} voxels")
    arr = np.frombuffer(raw, dtype=dt, count=count, offset=vox_offset).reshape(shape, order="F")
    if slope != 0 and not (slope == 1.0 and inter == 0.0) and np.isfinite(slope) and np.isfinite(inter):
        arr = arr.astype(np.float64) * float(slope) + float(inter)
    return np.ascontiguousarray(arr, dtype=np.float32)
