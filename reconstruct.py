"""Reconstruction CLI: the flag surface of the reference's `reconstruct.py` (same names, types and defaults, so existing
launch scripts work unchanged) driving the B200 engine.

Three extra, optional flags that the reference does not have:
  --plms_state carry|reset            carry (default) keeps the PNDM scheduler state across t-starts of a batch, exactly
                                      like the reference; reset gives every t-start chain a fresh PLMS history.
  --honour_num_inference_steps 0|1    the reference parses --num_inference_steps but always uses 100
                                      (src/trainers/reconstruct.py:118); 1 makes the flag effective.
  --shard images|t_starts             under torchrun, what the ranks divide: the images (default, what the reference
                                      does) or the t-start grid (every rank runs its share of the grid on every image;
                                      exact only with --plms_state reset, refused otherwise).
"""
import argparse
import ast

# (flag, type, default, what it does) - names, types and defaults are the reference's (reconstruct.py:7-141); the wording is ours
_lit = ast.literal_eval
FLAGS = [
    ("seed", int, 2, "RNG seed"),
    ("output_dir", None, None, "directory that holds the run directories"),
    ("model_name", None, None, "run directory under --output_dir (checkpoint in, ood/results_*.csv out)"),
    ("validation_ids", None, None, "csv with the validation image paths"),
    ("in_ids", None, None, "csv with the in-distribution image paths"),
    ("out_ids", None, None, "comma-separated csv files, one per out-of-distribution set (suffix _vflip / _hflip: flipped)"),
    ("spatial_dimension", int, 2, "2 or 3"),
    ("image_size", None, None, "resize every image to this edge length"),
    ("image_roi", _lit, None, "centre crop, a tuple; -1 keeps a dimension whole"),
    ("latent_pad", _lit, None, "F.pad tuple applied to the latent so that the UNet's downsamplings divide it"),
    ("vqvae_checkpoint", None, None, "stage-1 checkpoint of a latent diffusion model (vqvae_config.json next to it)"),
    ("ddpm_checkpoint_epoch", None, None, "use checkpoint_<epoch>.pth instead of checkpoint.pth"),
    ("prediction_type", None, "epsilon", "epsilon | sample | v_prediction"),
    ("model_type", None, "small", "small | big"),
    ("beta_schedule", None, "linear", "linear | scaled_linear_beta"),
    ("beta_start", float, 1e-4, "first beta"),
    ("beta_end", float, 2e-2, "last beta"),
    ("b_scale", float, 1, "data scale applied before noising"),
    ("snr_shift", float, 1, "SNR shift of the noise schedule"),
    ("simplex_noise", int, 0, "1: simplex noise instead of Gaussian"),
    ("batch_size", int, 256, "images per batch"),
    ("augmentation", int, 0, "accepted for launch-script compatibility (training only)"),
    ("cache_data", int, 1, "accepted for launch-script compatibility"),
    ("num_workers", int, 8, "accepted for launch-script compatibility"),
    ("first_n_val", None, None, "only the first n validation images"),
    ("first_n", None, None, "only the first n images of every other set"),
    ("eval_checkpoint", None, None, "accepted for launch-script compatibility"),
    ("drop_last", None, False, "drop a trailing partial batch"),
    ("is_grayscale", int, 0, "1: single-channel images"),
    ("run_val", int, 1, "score the validation set"),
    ("run_in", int, 1, "score the in-distribution set"),
    ("run_out", int, 1, "score the out-of-distribution sets"),
    ("num_inference_steps", int, 100, "PLMS steps (ignored unless --honour_num_inference_steps 1, like the reference)"),
    ("inference_skip_factor", int, 1, "use every k-th timestep as a reconstruction starting point"),
]
EXTENSIONS = [  # see the module docstring
    ("plms_state", dict(default="carry", choices=["carry", "reset"], help="PLMS history across the t-starts of a batch")),
    ("honour_num_inference_steps", dict(type=int, default=0, help="1: --num_inference_steps takes effect")),
    ("shard", dict(default="images", choices=["images", "t_starts"],
                   help="what torchrun ranks divide: images, or the t-start grid (needs --plms_state reset)")),
]


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    for name, typ, default, text in FLAGS:
        kw = dict(default=default, help=text)
        if typ is not None:
            kw["type"] = typ
        parser.add_argument("--" + name, **kw)
    for name, kw in EXTENSIONS:
        parser.add_argument("--" + name, **kw)
    return parser


def parse_args(argv=None):
    return build_parser().parse_args(argv)


if __name__ == "__main__":
    from ddpm_ood_b200.trainers import Reconstruct

    args = parse_args()
    recon = Reconstruct(args)
    recon.reconstruct(args)
