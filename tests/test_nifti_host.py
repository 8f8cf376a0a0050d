"""Host-side NIfTI-1 ingest (ddpm_ood_b200/nifti.py, data.py): the files are written here, byte by byte from the NIfTI-1
header layout (an independent writer: nibabel is not part of this environment), and read back through the reader and
through the reference-shaped loader (src/data/get_train_and_val_dataloader.py:60-84: LoadImaged, EnsureChannelFirstd,
first modality of a 4-D file, centre crop, ScaleIntensityd)."""
import gzip
import struct

import numpy as np
import pytest
import torch

from ddpm_ood_b200 import data as D
from ddpm_ood_b200.nifti import NiftiError, read_nifti

CODES = {"uint8": (2, 8), "int16": (4, 16), "int32": (8, 32), "float32": (16, 32), "float64": (64, 64),
         "uint16": (512, 16)}


def write_nifti(path, arr, slope=0.0, inter=0.0, big_endian=False, vox_offset=352, magic=b"n+1\x00", sizeof_hdr=348):
    bo = ">" if big_endian else "<"
    code, bitpix = CODES[arr.dtype.name]
    hdr = bytearray(348)
    struct.pack_into(bo + "i", hdr, 0, sizeof_hdr)
    dim = [arr.ndim] + list(arr.shape) + [1] * (7 - arr.ndim)
    struct.pack_into(bo + "8h", hdr, 40, *dim)
    struct.pack_into(bo + "2h", hdr, 70, code, bitpix)
    struct.pack_into(bo + "8f", hdr, 76, 1.0, *([1.0] * 7))
    struct.pack_into(bo + "f", hdr, 108, float(vox_offset))
    struct.pack_into(bo + "2f", hdr, 112, slope, inter)
    hdr[344:348] = magic
    body = bytes(hdr) + b"\x00" * (vox_offset - 348) + arr.astype(arr.dtype.newbyteorder(bo)).tobytes(order="F")
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "wb") as f:
        f.write(body)


@pytest.mark.parametrize("dtype", list(CODES))
@pytest.mark.parametrize("big_endian", [False, True], ids=["le", "be"])
def test_reads_every_datatype_in_file_axis_order(tmp_path, dtype, big_endian):
    rng = np.random.RandomState(0)
    arr = (rng.rand(5, 4, 3) * 200).astype(dtype)
    p = tmp_path / "v.nii"
    write_nifti(p, arr, big_endian=big_endian)
    got = read_nifti(str(p))
    assert got.dtype == np.float32 and got.shape == (5, 4, 3)
    assert np.array_equal(got, arr.astype(np.float32))  # index [i, j, k] as stored, i fastest on disk


def test_gzip_scaling_offset_and_4d(tmp_path):
    rng = np.random.RandomState(1)
    arr = rng.randint(-500, 500, size=(6, 5, 4, 4)).astype(np.int16)  # BraTS-like: four modalities in one file
    p = tmp_path / "brats.nii.gz"
    write_nifti(p, arr, slope=0.5, inter=-3.0, vox_offset=416)  # a header extension pushes the data back
    got = read_nifti(str(p))
    assert got.shape == (6, 5, 4, 4)
    assert np.allclose(got, arr.astype(np.float64) * 0.5 - 3.0)
    # slope 0 (and the identity pair) mean "no scaling"
    write_nifti(p, arr, slope=0.0, inter=7.0)
    assert np.array_equal(read_nifti(str(p)), arr.astype(np.float32))
    write_nifti(p, arr, slope=1.0, inter=0.0)
    assert np.array_equal(read_nifti(str(p)), arr.astype(np.float32))
    # a trailing singleton 4th dim is a plain volume
    write_nifti(p, arr[..., :1])
    assert read_nifti(str(p)).shape == (6, 5, 4)


def test_rejects_what_it_does_not_read(tmp_path):
    arr = np.zeros((2, 2, 2), dtype=np.float32)
    p = tmp_path / "x.nii"
    write_nifti(p, arr, sizeof_hdr=540)
    with pytest.raises(NiftiError, match="NIfTI-2"):
        read_nifti(str(p))
    write_nifti(p, arr, magic=b"ni1\x00")
    with pytest.raises(NiftiError, match="two-file"):
        read_nifti(str(p))
    write_nifti(p, arr, magic=b"abcd")
    with pytest.raises(NiftiError, match="magic"):
        read_nifti(str(p))
    write_nifti(p, arr)
    p.write_bytes(p.read_bytes()[:-4])
    with pytest.raises(NiftiError, match="fewer"):
        read_nifti(str(p))
    p.write_bytes(b"\x00" * 100)
    with pytest.raises(NiftiError, match="shorter"):
        read_nifti(str(p))


def test_loader_gives_the_reference_transform_chain_on_nifti(tmp_path):
    """is_grayscale=True, spatial_dimension=3 on a 4-D file: channel-first, first modality, centre crop, min-max to
    [0, 1] (get_train_and_val_dataloader.py:66-76) - batches shaped [B, 1, X, Y, Z] with the file names alongside."""
    rng = np.random.RandomState(2)
    paths = []
    vols = []
    for i in range(3):
        arr = rng.randint(0, 2000, size=(10, 12, 8, 4)).astype(np.int16)
        p = tmp_path / f"sub{i}.nii.gz"
        write_nifti(p, arr)
        paths.append(str(p))
        vols.append(arr)
    ids = tmp_path / "val.csv"
    ids.write_text(",".join(paths) + "\n")
    loader = D.get_training_data_loader(batch_size=2, training_ids=str(ids), validation_ids=str(ids), only_val=True,
                                        is_grayscale=True, spatial_dimension=3, image_roi=[8, 8, 8])
    batches = list(loader)
    assert [b["image"].shape for b in batches] == [torch.Size([2, 1, 8, 8, 8]), torch.Size([1, 1, 8, 8, 8])]
    assert batches[0]["image_meta_dict"]["filename_or_obj"] == paths[:2]
    want = torch.from_numpy(vols[1][1:9, 2:10, :, 0].astype(np.float32))
    want = (want - want.min()) / (want.max() - want.min())
    assert torch.allclose(batches[0]["image"][1, 0], want)
