"""Generate the full-size parity fixtures from the fp32 CPU oracle (run in the build container; minutes of CPU).

    python tests/golden/make_fullsize_fixtures.py [unet] [vae]

* unet_sdxl_b2_128.npz -- ONE CFG-batch-2 forward of the full SDXL-base UNet (2 567 463 684 parameters,
                          oracle/sdxl_unet.py, fp32, weights rounded to fp16 first) at the BENCHMARKED shape
                          128x128 latents (1024 px): ``eps`` [2,4,128,128] fp32 plus a checksum of the seeded
                          weights so a drift of the init recipe is detected before the comparison.
* vae_sdxl_64.npz      -- one decode of the SDXL-width VAE decoder (128,256,512,512; oracle/vae.py, fp32) at
                          64x64 latents -> uint8 frame [512,512,3].

The seeded inputs are rebuilt by the tests with the functions below (same generator calls), so only the oracle's
OUTPUT is stored.  The oracle restates un-vendored diffusers 0.25.0 (parity unpinned by the reference itself,
see DESIGN.md section 2); these fixtures pin the CUDA path to the oracle at the shape bench.py measures.
"""
import hashlib
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

UNET_FIXTURE = os.path.join(HERE, "unet_sdxl_b2_128.npz")
VAE_FIXTURE = os.path.join(HERE, "vae_sdxl_64.npz")
UNET_T = 499.0
UNET_SEED = 0
VAE_SEED = 4
VAE_CHANNELS = (128, 256, 512, 512)


def unet_inputs(cfg, B, h, w, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 4, h, w, generator=g).half()
    ctx = (torch.randn(B, 77, cfg.cross_attention_dim, generator=g) * 0.5).half()
    pooled = torch.randn(B, cfg.pooled_dim, generator=g).half()
    tids = torch.tensor([[8.0 * h, 8.0 * w, 0, 0, 8.0 * h, 8.0 * w]] * B).half()
    return x, ctx, pooled, tids


def vae_latent(h, w, seed=1):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(1, 4, h, w, generator=g) * 0.8).half()


def weights_checksum(state_dict, stride=4099):
    """sha1 over a strided sample of every parameter (fp16 bytes) -- cheap, catches any init drift."""
    h = hashlib.sha1()
    for k in sorted(state_dict):
        v = state_dict[k].detach().reshape(-1)
        h.update(k.encode())
        h.update(v[::stride].to(torch.float16).cpu().numpy().tobytes())
    return h.hexdigest()


def oracle_unet():
    from oracle.sdxl_unet import SDXL_BASE, SDXLUNet, synthetic_init_
    net = synthetic_init_(SDXLUNet(SDXL_BASE), seed=UNET_SEED).eval()
    with torch.no_grad():
        for p in net.parameters():
            p.copy_(p.half().float())
    return net


def oracle_vae():
    from oracle.vae import VAEConfig, VAEDecoder, synthetic_vae_init_
    cfg = VAEConfig(block_out_channels=VAE_CHANNELS)
    ov = synthetic_vae_init_(VAEDecoder(cfg), seed=VAE_SEED).eval()
    with torch.no_grad():
        for p in ov.parameters():
            p.copy_(p.half().float())
    return ov, cfg


def make_unet():
    from oracle.sdxl_unet import SDXL_BASE
    t0 = time.time()
    net = oracle_unet()
    x, ctx, pooled, tids = unet_inputs(SDXL_BASE, 2, 128, 128, UNET_SEED)
    t1 = time.time()
    with torch.no_grad():
        eps = net(x.float(), UNET_T, ctx.float(), pooled.float(), tids.float())
    t2 = time.time()
    np.savez_compressed(UNET_FIXTURE, eps=eps.numpy().astype(np.float32), t=np.float32(UNET_T),
                        weights_sha1=np.array(weights_checksum(net.state_dict())),
                        threads=np.int32(torch.get_num_threads()), seconds=np.float32(t2 - t1))
    print(f"unet fixture: init {t1 - t0:.0f}s forward {t2 - t1:.0f}s  |eps|={eps.norm():.4f} "
          f"finite={bool(torch.isfinite(eps).all())} -> {UNET_FIXTURE}")


def make_vae():
    from oracle.vae import latent2image_np
    ov, cfg = oracle_vae()
    lat = vae_latent(64, 64)
    t0 = time.time()
    with torch.no_grad():
        frame = latent2image_np(ov, lat)
    np.savez_compressed(VAE_FIXTURE, frame=frame, weights_sha1=np.array(weights_checksum(ov.state_dict())),
                        seconds=np.float32(time.time() - t0))
    print(f"vae fixture: {time.time() - t0:.0f}s frame {frame.shape} std {frame.std():.1f} -> {VAE_FIXTURE}")


if __name__ == "__main__":
    what = sys.argv[1:] or ["vae", "unet"]
    if "vae" in what:
        make_vae()
    if "unet" in what:
        make_unet()
