"""Host-side contract of the drop-in VQVAE class (no GPU): constructor surface and state_dict keys equal the oracle
restatement of monai-generative's VQVAE (so a reference-trained vqvae checkpoint loads strict=True), unsupported
configurations are refused at construction, and there is no CPU compute path."""
import pytest
import torch

CFG = dict(spatial_dims=3, in_channels=1, out_channels=1, num_channels=(128, 256), num_res_layers=2,
           num_res_channels=(128, 256), downsample_parameters=((2, 4, 1, 1),) * 2,
           upsample_parameters=((2, 4, 1, 1, 0),) * 2, num_embeddings=64, embedding_dim=128)


def test_state_dict_keys_and_shapes_match_the_oracle():
    from ddpm_ood_b200.vqvae import VQVAE
    from oracle import vqvae as ov

    ref = ov.VQVAE(**CFG)
    ours = VQVAE(**CFG)
    a, b = ref.state_dict(), ours.state_dict()
    assert list(sorted(a)) == list(sorted(b))
    for k in a:
        assert a[k].shape == b[k].shape, k
    ours.load_state_dict(a, strict=True)
    # the README configuration (README.md:153-159) is accepted, with the reference's extra trainer kwargs
    VQVAE(spatial_dims=3, in_channels=1, out_channels=1, num_channels=[256] * 4, num_res_layers=3,
          num_res_channels=[256] * 4, downsample_parameters=[[2, 4, 1, 1]] * 4, upsample_parameters=[[2, 4, 1, 1, 0]] * 4,
          num_embeddings=2048, embedding_dim=128, decay=0.99, commitment_cost=0.25, epsilon=1e-5, dropout=0.0,
          ddp_sync=True)


def test_unsupported_configurations_are_refused():
    from ddpm_ood_b200.vqvae import VQVAE

    with pytest.raises(NotImplementedError):
        VQVAE(**{**CFG, "downsample_parameters": ((1, 3, 1, 1),) * 2})
    with pytest.raises(NotImplementedError):
        VQVAE(**{**CFG, "upsample_parameters": ((2, 4, 1, 1, 1),) * 2})
    with pytest.raises(NotImplementedError):
        VQVAE(**CFG, dropout=0.1)
    with pytest.raises(ValueError):
        VQVAE(**{**CFG, "num_res_channels": (128,)})


def test_no_cpu_path():
    from ddpm_ood_b200._lib import DdpmError
    from ddpm_ood_b200.vqvae import VQVAE

    m = VQVAE(**CFG)
    with pytest.raises(DdpmError):
        m.encode_stage_2_inputs(torch.zeros(1, 1, 16, 16, 16))
    with pytest.raises(DdpmError):
        m.decode_stage_2_outputs(torch.zeros(1, 128, 4, 4, 4))
