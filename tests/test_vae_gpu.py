"""GPU parity of the native VAE decoder (latentblending_b200/vae.py) against the CPU oracle decoder
(oracle/vae.py, fp32) on identical seeded weights.  Tolerance (stated): uint8 frames differ by <= 1.0 levels on
average and <= 12 levels anywhere (fp16 storage vs fp32)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("h,w", [(16, 16), (32, 16)])
def test_vae_decoder_matches_oracle(h, w):
    from latentblending_b200.vae import VAEDecoderB200
    from oracle.vae import VAEConfig, VAEDecoder, latent2image_np, synthetic_vae_init_
    cfg = VAEConfig(block_out_channels=(64, 64, 128, 128))
    ov = synthetic_vae_init_(VAEDecoder(cfg), seed=4).eval()
    with torch.no_grad():
        for p in ov.parameters():
            p.copy_(p.half().float())
    g = torch.Generator().manual_seed(1)
    lat = (torch.randn(1, 4, h, w, generator=g) * 0.8).half()
    with torch.no_grad():
        ref = latent2image_np(ov, lat)
    vae = VAEDecoderB200(ov.state_dict(), cfg.block_out_channels, cfg.scaling_factor, "cuda:0")
    got = vae.decode_to_u8(lat.cuda()).cpu().numpy()
    assert got.shape == ref.shape == (8 * h, 8 * w, 3)
    d = np.abs(got.astype(np.int32) - ref.astype(np.int32))
    assert d.mean() <= 1.0 and d.max() <= 12, (d.mean(), d.max())
    assert ref.std() > 5            # the frame is not degenerate
    from latentblending_b200 import ops
    assert ops.error_flag() == 0
