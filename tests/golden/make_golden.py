"""Generate the committed golden vectors from the oracle (run here, on CPU):  python tests/golden/make_golden.py

The reference itself cannot be imported (its model/scheduler/LPIPS live in packages that are not installable here,
SURVEY.md §8c), so these are outputs of oracle/, the restatement; parity with the third-party originals stays unpinned.
The one reference-side pin, README.md:118-120 (skip factor -> number of reconstructions), is stored in t_grid.json.
"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import unet as ou  # noqa: E402
from oracle.lpips import PerceptualLoss  # noqa: E402
from oracle.pndm import PNDMScheduler, t_start_grid  # noqa: E402
from oracle.recon_loop import LoopConfig, reconstruct_batch  # noqa: E402

OUT = Path(__file__).resolve().parent


def main():
    torch.set_num_threads(8)
    # 1. t-start grids for every skip factor of the README table
    s = PNDMScheduler(num_train_timesteps=1000, skip_prk_steps=True)
    s.set_timesteps(100)
    readme = {1: 100, 2: 50, 3: 34, 4: 25, 5: 20, 8: 13, 16: 7, 32: 4, 64: 2}
    grids = {str(k): t_start_grid(s.timesteps, k).tolist() for k in readme}
    for k, n in readme.items():
        assert len(grids[str(k)]) == n
    (OUT / "t_grid.json").write_text(json.dumps({"timesteps": s.timesteps.tolist(), "readme_counts": readme,
                                                  "grids": grids}))

    # 2. one UNet forward (small, 1x32x32, B=2)
    seed = 0
    model = ou.randomize_(ou.make_small(2, 1), seed=seed).eval()
    g = torch.Generator().manual_seed(42)
    x = torch.randn((2, 1, 32, 32), generator=g)
    t = torch.tensor([990, 10])
    with torch.no_grad():
        y = model(x, t)
    torch.save({"weight_seed": seed, "x": x, "t": t, "y": y}, OUT / "unet_small_1x32x32.pt")

    # 3. reconstruction loop, BASELINE config-1 shape at skip 32 (4 t-starts), B=2, both PLMS state modes
    pl = PerceptualLoss(dimensions=2, include_pixel_loss=False, is_fake_3d=False, lpips_normalize=True, spatial=False)
    x0 = torch.rand((2, 1, 32, 32), generator=torch.Generator().manual_seed(100))
    noise = [torch.randn((2, 1, 32, 32), generator=torch.Generator().manual_seed(1000 + i)) for i in range(4)]
    gold = {"weight_seed": seed, "x0": x0, "noise": noise, "skip": 32}
    for mode in ("carry", "reset"):
        cfg = LoopConfig(inference_skip_factor=32, plms_state=mode)
        r = reconstruct_batch(model, pl, x0, lambda i, t_: noise[i], cfg)
        gold["t"] = r["t"]
        gold[mode] = {"mse": r["mse"], "perceptual_difference": r["perceptual_difference"]}
        print(mode, r["mse"].flatten().tolist(), r["perceptual_difference"].flatten().tolist())
    torch.save(gold, OUT / "recon_fmnist_skip32.pt")


if __name__ == "__main__":
    main()
