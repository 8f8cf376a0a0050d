"""Score post-processing on the device: the z-score / per-file mean / ROC-AUC arithmetic of the
reference's `ood_detection.py:150-206` for score tensors that are still on the GPU (`BatchReconstructor.score_batch`
returns `[n_t, B]` fp32 per target). The CSV / pandas route of the reference keeps working: `trainers.Reconstruct`
writes the same files."""
from __future__ import annotations

import torch

from . import _lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _f32(x: torch.Tensor) -> torch.Tensor:
    if not x.is_cuda:
        raise _lib.DdpmError("ood scoring needs CUDA tensors; there is no CPU fallback")
    return x.detach().float().contiguous()


def val_stats(val: torch.Tensor):
    """val: [n_t, n_val] -> (mean [n_t], sample std [n_t]) on the device."""
    val = _f32(val)
    n_t, n = val.shape
    mean = torch.empty(n_t, dtype=torch.float32, device=val.device)
    std = torch.empty_like(mean)
    with torch.cuda.device(val.device):
        _lib.check(_lib.lib().ddpm_val_stats(val.data_ptr(), n_t, n, mean.data_ptr(), std.data_ptr(), _stream()),
                   "ddpm_val_stats")
    return mean, std


def mean_z(scores: torch.Tensor, mean: torch.Tensor, std: torch.Tensor) -> torch.Tensor:
    """scores: [n_t, n] -> [n] mean z-score per image."""
    scores = _f32(scores)
    n_t, n = scores.shape
    out = torch.empty(n, dtype=torch.float32, device=scores.device)
    with torch.cuda.device(scores.device):
        _lib.check(_lib.lib().ddpm_mean_z(scores.data_ptr(), _f32(mean).data_ptr(), _f32(std).data_ptr(), n_t, n,
                                          out.data_ptr(), _stream()), "ddpm_mean_z")
    return out


def roc_auc(in_scores: torch.Tensor, out_scores: torch.Tensor) -> float:
    """ROC-AUC with `out` as the positive class (sklearn.metrics.roc_auc_score on the concatenation)."""
    a, b = _f32(in_scores), _f32(out_scores)
    counts = torch.zeros(2, dtype=torch.int64, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(_lib.lib().ddpm_auc_counts(a.data_ptr(), a.numel(), b.data_ptr(), b.numel(), counts.data_ptr(),
                                              _stream()), "ddpm_auc_counts")
    gt, eq = (int(v) for v in counts.tolist())
    return (gt + 0.5 * eq) / (a.numel() * b.numel())


def select_t(t_values, min_t: int = 0, max_t: int = 1000):
    """Row indices of the t-start grid the reference's post-processing keeps (host integer work): the first occurrence of
    every t (`drop_duplicates(subset=["filename", "t"], keep="first")`, ood_detection.py:53-54,143-145 - a skip-1 grid
    holds t = 980 twice) with MIN_T < t < MAX_T (:56-62,146-147)."""
    keep, seen = [], set()
    for i, t in enumerate(int(v) for v in t_values):
        if t in seen:
            continue
        seen.add(t)
        if min_t < t < max_t:
            keep.append(i)
    return keep


def ood_auc(val: torch.Tensor, ins: torch.Tensor, outs: torch.Tensor, t=None, min_t: int = 0, max_t: int = 1000) -> float:
    """val / ins / outs: [n_t, n_*] scores of one target for the validation, in-distribution and out-of-distribution
    sets (same t grid). Returns the AUC the reference prints (`Zscore_<target>`). t: the grid's t values ([n_t], what
    `score_batch` returns as "t"); when given, duplicate t rows are dropped (first kept) and the reference's
    min_t < t < max_t filter is applied before averaging, as ood_detection.py does on the CSV rows."""
    if t is not None:
        rows = select_t(t.tolist() if hasattr(t, "tolist") else t, min_t, max_t)
        if len(rows) != val.shape[0]:
            idx = torch.tensor(rows, dtype=torch.long, device=val.device)
            val, ins, outs = val.index_select(0, idx), ins.index_select(0, idx), outs.index_select(0, idx)
    m, s = val_stats(val)
    return roc_auc(mean_z(ins, m, s), mean_z(outs, m, s))
