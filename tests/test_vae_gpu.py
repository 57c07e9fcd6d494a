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


def test_sdxl_width_vae_matches_fixture():
    """The SDXL-width decoder (128,256,512,512) at 64x64 latents vs the fp32 oracle frame committed as a fixture
    (tests/golden/vae_sdxl_64.npz, tests/golden/make_fullsize_fixtures.py).  Same tolerance as above."""
    from latentblending_b200 import ops
    from latentblending_b200.vae import VAEDecoderB200
    from make_fullsize_fixtures import VAE_FIXTURE, oracle_vae, vae_latent, weights_checksum
    fx = np.load(VAE_FIXTURE)
    ov, cfg = oracle_vae()
    assert weights_checksum(ov.state_dict()) == str(fx["weights_sha1"]), "seeded VAE recipe drifted"
    vae = VAEDecoderB200(ov.state_dict(), cfg.block_out_channels, cfg.scaling_factor, "cuda:0")
    got = vae.decode_to_u8(vae_latent(64, 64).cuda()).cpu().numpy()
    ref = fx["frame"]
    assert got.shape == ref.shape == (512, 512, 3)
    d = np.abs(got.astype(np.int32) - ref.astype(np.int32))
    print(f"SDXL-width VAE @64x64: mean |d|={d.mean():.3f} max={d.max()} levels")
    assert d.mean() <= 1.0 and d.max() <= 12, (d.mean(), d.max())
    assert ops.error_flag() == 0
    assert vae.overflow_count() == 0
