"""Drop-in for the reference's `ood_detection.py` (read the results_*.csv files a reconstruction run wrote, z-score every
(file, t) row against the validation set, average per file, print the ROC-AUC of every out-of-distribution set) with the
arithmetic of :150-206 done by the device kernels of ddpm_ood_b200/ood.py (`ddpm_val_stats`, `ddpm_mean_z`,
`ddpm_auc_counts`) instead of pandas merges and sklearn.

Same flags (--seed, --output_dir, --model_name, --max_t, --min_t, --t_skip), same dataset table keyed on the model name
(ood_detection.py:92-134), same duplicate removal (:53-54,143-145), t filter (:56-62) and printed lines (:217-223).
pandas is used to READ the CSVs only. Plotting is not part of the path.
"""
from __future__ import annotations

import argparse
from pathlib import Path

import numpy as np
import pandas as pd
import torch

from ddpm_ood_b200 import ood

MEDNIST = ["AbdomenCT", "BreastMRI", "ChestCT", "CXR", "Hand", "HeadCT"]


def parse_args(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--seed", type=int, default=2, help="Random seed to use.")
    parser.add_argument("--output_dir", help="Location for models.")
    parser.add_argument("--model_name", help="Name of model.")
    parser.add_argument("--max_t", type=int, default=1000, help="Maximum T to consider reconstructions from.")
    parser.add_argument("--min_t", type=int, default=0, help="Minimum T to consider reconstructions from.")
    parser.add_argument("--t_skip", type=int, default=1, help="Only use every n reconstructions.")
    parser.add_argument("--plot_target", default="mse", choices=["mse", "perceptual_difference"],
                        help="score column (the reference hard-codes 'mse', ood_detection.py:74)")
    parser.add_argument("--device", default="cuda:0")
    return parser.parse_args(argv)


def out_datasets(model: str):
    """The model-name table of ood_detection.py:92-134."""
    m = model
    if "fashionmnist" in m:
        return ["MNIST", "FashionMNIST_vflip", "FashionMNIST_hflip"]
    if "mnist" in m:
        return ["FashionMNIST", "MNIST_vflip", "MNIST_hflip"]
    if "cifar10" in m:
        return ["SVHN", "CelebA", "CIFAR10_vflip", "CIFAR10_hflip"]
    if "celeba" in m.lower():
        return ["CIFAR10", "SVHN", "CelebA_vflip", "CelebA_hflip"]
    if "svhn" in m:
        return ["CIFAR10", "CelebA", "SVHN_vflip", "SVHN_hflip"]
    for key, name in (("abdomenct", "AbdomenCT"), ("breastmri", "BreastMRI"), ("cxr", "CXR"), ("chestct", "ChestCT"),
                      ("hand", "Hand"), ("headct", "HeadCT")):
        if key in m:
            return [d for d in MEDNIST if d != name]
    if "decathlon" in m or "Task01" in m:
        return [f"Task{i:02d}" for i in range(2, 11)]
    raise ValueError(f"Unknown dataset to select for run_dir {model}")


def _dense(df: pd.DataFrame, t_values, target: str, device) -> tuple[torch.Tensor, int]:
    """CSV rows -> [n_t, n_files] fp32 on the device (first row of every (filename, t) pair, the kept t values only)."""
    df = df.drop_duplicates(subset=["filename", "t"], keep="first")
    df = df[df["t"].isin(t_values)]
    table = df.pivot(index="t", columns="filename", values=target).reindex(index=list(t_values))
    if table.isna().any().any():
        raise ValueError("every file needs one row per kept t value")
    return torch.from_numpy(table.to_numpy(dtype=np.float32)).to(device), table.shape[1]


def main(args) -> dict:
    torch.manual_seed(args.seed)
    model = args.model_name
    run_dir = Path(args.output_dir) / model
    print(f"Run directory: {str(run_dir)}")
    out_dir = run_dir / "ood"
    df_val = pd.read_csv(out_dir / "results_val.csv")
    df_val = df_val.drop_duplicates(subset=["filename", "t"], keep="first")
    all_t = df_val["t"].unique()
    t_values = [int(t) for t in all_t[ood.select_t(all_t.tolist(), args.min_t, args.max_t)]]
    target = args.plot_target
    print(f"SETTING MAX_T to {args.max_t} and T_SKIP to 1 with a total of {len(t_values)} starting points")
    print(f"Plot target is {target}")
    val, n_val = _dense(df_val, t_values, target, args.device)
    mean, std = ood.val_stats(val)
    df_in = pd.read_csv(out_dir / "results_in.csv")
    ins, n_in = _dense(df_in, t_values, target, args.device)
    z_in = ood.mean_z(ins, mean, std)
    results = {"ood_data": [], f"Zscore_{target}": []}
    for name in out_datasets(model):
        outs, n_out = _dense(pd.read_csv(out_dir / f"results_{name}.csv"), t_values, target, args.device)
        auc = ood.roc_auc(z_in, ood.mean_z(outs, mean, std))
        print(f"n_val={n_val} n_in={n_in} n_out={n_out}")
        results["ood_data"].append(name)
        results[f"Zscore_{target}"].append(auc)
    for o, s in zip(results["ood_data"], results[f"Zscore_{target}"]):
        print(f"AUC for {model} vs {o}: {s * 100:.1f}")
    print(f"Average AUC: {np.mean(results[f'Zscore_{target}']) * 100:.1f}")
    return results


if __name__ == "__main__":
    main(parse_args())
