"""Golden vector for the latent-diffusion route of BASELINE config 5 (image -> VQ-VAE encode -> 3-D latent chains ->
VQ-VAE decode -> MSE + per-item 2.5-D LPIPS), generated from the fp32 oracle on CPU:

    python tests/golden/make_golden_ldm.py            (a few minutes on 8 cores)

Scaled so the oracle finishes: image [2,1,32,32,32], a two-level VQ-VAE (128/256 channels, 256 x 128 codebook) -> latent
[2,128,8,8,8] = config 5's latent shape, the small 3-D UNet, skip 32 -> 4 t-starts (200 UNet evaluations). Like the other
goldens these are outputs of oracle/ (restatements; parity with the third-party originals is unpinned).
"""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import unet as ou  # noqa: E402
from oracle import vqvae as ov  # noqa: E402
from oracle.lpips import PerceptualLoss  # noqa: E402
from oracle.recon_loop import LoopConfig, reconstruct_batch  # noqa: E402

OUT = Path(__file__).resolve().parent
VQ_CFG = dict(spatial_dims=3, in_channels=1, out_channels=1, num_channels=(128, 256), num_res_layers=1,
              num_res_channels=(128, 256), downsample_parameters=((2, 4, 1, 1),) * 2,
              upsample_parameters=((2, 4, 1, 1, 0),) * 2, num_embeddings=256, embedding_dim=128)


def main():
    torch.set_num_threads(8)
    t0 = time.time()
    vq = ov.randomize_(ov.VQVAE(**VQ_CFG), seed=0).eval()
    model = ou.randomize_(ou.make_small(3, 128), seed=0).eval()
    pl = PerceptualLoss(dimensions=3, include_pixel_loss=False, is_fake_3d=True, lpips_normalize=True, spatial=False)
    x0 = torch.rand((2, 1, 32, 32, 32), generator=torch.Generator().manual_seed(100))
    lat_shape = (2, 128, 8, 8, 8)
    noise = [torch.randn(lat_shape, generator=torch.Generator().manual_seed(1000 + i)) for i in range(4)]
    cfg = LoopConfig(inference_skip_factor=32, spatial_dimension=3)
    r = reconstruct_batch(model, pl, x0, lambda i, t_: noise[i], cfg, vqvae=vq, keep_indices=True)
    with torch.no_grad():  # positions whose two best rows are equidistant to within fp32 summation-order noise
        _, d, _ = vq.quantizer.quantizer.quantize(vq.encode(x0))
        top2 = torch.topk(-d, 2, dim=1).values
        ties = ((top2[:, 0] - top2[:, 1]).abs() <= 2e-5 * d.abs().max(dim=1).values).view(r["enc_indices"].shape)
    gold = {"enc_ties": ties, "weight_seed": 0, "vq_cfg": VQ_CFG, "x0": x0, "noise_seeds": [1000 + i for i in range(4)],
            "noise_shape": lat_shape, "skip": 32, "num_inference_steps": 100, "plms_state": "carry", "t": r["t"],
            "mse": r["mse"], "perceptual_difference": r["perceptual_difference"],
            "enc_indices": r["enc_indices"].to(torch.int16), "dec_indices": r["dec_indices"].to(torch.int16)}
    torch.save(gold, OUT / "recon_cfg5_ldm.pt")
    print(f"recon_cfg5_ldm.pt: {time.time() - t0:.1f}s t={r['t'].tolist()} mse={r['mse'].tolist()} "
          f"pd={r['perceptual_difference'].tolist()}")


if __name__ == "__main__":
    main()
