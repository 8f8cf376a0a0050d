"""TEST INFRASTRUCTURE (checker only): numpy restatement of the score post-processing of the reference's
`ood_detection.py`: per-t z-scores against the validation set, mean z per file, ROC-AUC of
out-of-distribution vs in-distribution files.

Follows /root/reference/ood_detection.py:
  :150-161  val mean / std per t (pandas `.agg(["mean", "std"])`: sample std, ddof = 1), z = (x - mean_t) / std_t
  :174      per-file mean over the t values (`groupby(["filename", "type"]).mean()`)
  :191-206  `roc_auc_score(class, score)` with class 0 = "in", 1 = "out"
Inputs are dense arrays [n_t, n_images] per dataset (the CSV rows of one target, `mse` or `perceptual_difference`,
after the reference's duplicate removal and t filtering at :53-62,143-147).
"""
from __future__ import annotations

import numpy as np


def select_t(t_values, min_t: int = 0, max_t: int = 1000) -> np.ndarray:
    """Indices (first occurrences, in order) of the t-start columns the reference keeps: `drop_duplicates(subset=
    ["filename", "t"], keep="first")` removes the second t = 980 column of a skip-1 grid (ood_detection.py:53-54,143-145)
    and the strict filter MIN_T < t < MAX_T drops the rest (:56-62,146-147)."""
    keep, seen = [], set()
    for i, t in enumerate(np.asarray(t_values).tolist()):
        if t in seen:
            continue
        seen.add(t)
        if min_t < t < max_t:
            keep.append(i)
    return np.asarray(keep, dtype=np.int64)


def val_stats(val: np.ndarray):
    """val: [n_t, n_val] -> (mean [n_t], std [n_t]) with ddof = 1 (ood_detection.py:153-158)."""
    val = np.asarray(val, dtype=np.float64)
    return val.mean(axis=1), val.std(axis=1, ddof=1)


def mean_z(scores: np.ndarray, mean: np.ndarray, std: np.ndarray) -> np.ndarray:
    """scores: [n_t, n] -> per-image mean over t of (x - mean_t) / std_t (ood_detection.py:159-161,174)."""
    scores = np.asarray(scores, dtype=np.float64)
    return ((scores - mean[:, None]) / std[:, None]).mean(axis=0)


def roc_auc(in_scores: np.ndarray, out_scores: np.ndarray) -> float:
    """Area under the ROC curve with `out` as the positive class: the Mann-Whitney statistic
    P(out > in) + 0.5 P(out == in), which is what sklearn's roc_auc_score computes (ood_detection.py:206)."""
    a = np.asarray(in_scores, dtype=np.float64)[None, :]
    b = np.asarray(out_scores, dtype=np.float64)[:, None]
    return float(((b > a).sum() + 0.5 * (b == a).sum()) / (a.size * b.size))


def ood_auc(val: np.ndarray, ins: np.ndarray, outs: np.ndarray) -> float:
    m, s = val_stats(val)
    return roc_auc(mean_z(ins, m, s), mean_z(outs, m, s))
