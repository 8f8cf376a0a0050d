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

FLAGS = [
    # name, kwargs
    ("--seed", dict(type=int, default=2, help="Random seed to use.")),
    ("--output_dir", dict(help="Location for models.")),
    ("--model_name", dict(help="Name of model.")),
    ("--validation_ids", dict(help="Location of file with validation ids.")),
    ("--in_ids", dict(help="Location of file with inlier ids.")),
    ("--out_ids", dict(help="List of location of file with outlier ids.")),
    ("--spatial_dimension", dict(default=2, type=int, help="Dimension of images: 2d or 3d.")),
    ("--image_size", dict(default=None, help="Resize images.")),
    ("--image_roi", dict(default=None, type=ast.literal_eval,
                         help="Central ROI crop of inputs, as a tuple, with -1 to not crop a dimension.")),
    ("--latent_pad", dict(default=None, type=ast.literal_eval,
                          help="Padding applied to a latent so the U-net's two downsamplings divide it; a tuple in "
                               "torch.nn.functional.pad order.")),
    ("--vqvae_checkpoint", dict(default=None, help="Path to a VQ-VAE checkpoint, to reconstruct with an LDM.")),
    ("--ddpm_checkpoint_epoch", dict(default=None, help="Epoch of a specific checkpoint; default is the best one.")),
    ("--prediction_type", dict(default="epsilon", help="Scheduler prediction type: epsilon, sample or v_prediction.")),
    ("--model_type", dict(default="small", help="Small or big model.")),
    ("--beta_schedule", dict(default="linear", help="Linear or scaled linear")),
    ("--beta_start", dict(type=float, default=1e-4, help="Beta start.")),
    ("--beta_end", dict(type=float, default=2e-2, help="Beta end.")),
    ("--b_scale", dict(type=float, default=1, help="Scale the data by a factor b before noising.")),
    ("--snr_shift", dict(type=float, default=1, help="Shift the SNR of the noise scheduler by a factor.")),
    ("--simplex_noise", dict(type=int, default=0, help="Use simplex instead of Gaussian noise.")),
    ("--batch_size", dict(type=int, default=256, help="Batch size.")),
    ("--augmentation", dict(type=int, default=0, help="Use of augmentation, 1 (True) or 0 (False).")),
    ("--cache_data", dict(type=int, default=1, help="Whether or not to cache data in dataloaders.")),
    ("--num_workers", dict(type=int, default=8, help="Number of loader workers")),
    ("--first_n_val", dict(default=None, help="Only run on the first n samples from the val dataset.")),
    ("--first_n", dict(default=None, help="Only run on the first n samples from each dataset.")),
    ("--eval_checkpoint", dict(default=None, help="Select a specific checkpoint to evaluate on.")),
    ("--drop_last", dict(default=False, help="Drop last non-complete batch..")),
    ("--is_grayscale", dict(type=int, default=0, help="Is data grayscale.")),
    ("--run_val", dict(type=int, default=1, help="Run reconstructions on val set.")),
    ("--run_in", dict(type=int, default=1, help="Run reconstructions on in set.")),
    ("--run_out", dict(type=int, default=1, help="Run reconstructions on out set.")),
    ("--num_inference_steps", dict(type=int, default=100, help="Number of inference steps to use with the PLMS sampler.")),
    ("--inference_skip_factor", dict(type=int, default=1,
                                     help="Perform fewer reconstructions by skipping some t-values as starting points.")),
    # extensions (see module docstring)
    ("--plms_state", dict(default="carry", choices=["carry", "reset"], help="PLMS history across t-starts.")),
    ("--honour_num_inference_steps", dict(type=int, default=0, help="Make --num_inference_steps effective.")),
    ("--shard", dict(default="images", choices=["images", "t_starts"],
                     help="Under torchrun: split images over ranks (the reference) or the t-start grid (needs "
                          "--plms_state reset).")),
]


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    for name, kw in FLAGS:
        parser.add_argument(name, **kw)
    return parser


def parse_args(argv=None):
    return build_parser().parse_args(argv)


if __name__ == "__main__":
    from ddpm_ood_b200.trainers import Reconstruct

    args = parse_args()
    recon = Reconstruct(args)
    recon.reconstruct(args)
