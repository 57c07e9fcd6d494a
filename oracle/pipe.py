"""Oracle: a CPU stand-in for the ``StableDiffusionXLPipeline`` surface the
reference's holder touches (test infrastructure).

The reference reads 44 attributes off ``pipe``; diffusers is
not installed, there are no checkpoints, so this supplies the same arithmetic
with seeded random-init weights and seeded synthetic prompt embeddings:

* ``prompt_embeds  ~ 0.5*N(0,1) [1,77,ctx]``, ``pooled ~ N(0,1) [1,pooled]`` from a
  CPU generator seeded with ``crc32(prompt)`` (stands in for
  ``pipe.encode_prompt``, diffusers_holder.py:81-95);
* noise ``randn([1,4,h,w]) * init_noise_sigma`` in fp16 from a CPU generator
  (``pipe.prepare_latents``, diffusers_holder.py:100-109; the reference's CUDA
  Philox stream cannot be reproduced on CPU, so parity tests inject latents);
* ``add_time_ids = [H,W,0,0,H,W]`` with H,W from ``default_sample_size * 8``
  (diffusers_holder.py:216-220, 264-270).
"""
import zlib

import torch

from .schedulers import EulerAncestralDiscrete, EulerDiscrete
from .sdxl_unet import SDXL_BASE, SDXL_TURBO, SDXLUNet, UNetConfig, synthetic_init_
from .vae import SDXL_VAE, VAEDecoder, synthetic_vae_init_


def prompt_seed(prompt: str) -> int:
    return zlib.crc32(prompt.encode("utf-8")) & 0x7FFFFFFF


def synthetic_text_embedding(prompt, ctx_dim, pooled_dim, dtype=torch.float16):
    g = torch.Generator().manual_seed(prompt_seed(prompt))
    emb = (torch.randn(1, 77, ctx_dim, generator=g) * 0.5).to(dtype)
    pooled = torch.randn(1, pooled_dim, generator=g).to(dtype)
    return emb, pooled


class OraclePipe:
    vae_scale_factor = 8

    def __init__(self, name="stabilityai/stable-diffusion-xl-base-1.0", unet_cfg: UNetConfig = None,
                 vae_cfg=None, seed=0, build_vae=True, unet=None, vae=None):
        self._name_or_path = name
        self.is_turbo = "turbo" in name
        if unet_cfg is None:
            unet_cfg = SDXL_TURBO if self.is_turbo else SDXL_BASE
        self.unet_cfg = unet_cfg
        self.unet = unet if unet is not None else synthetic_init_(SDXLUNet(unet_cfg), seed=seed).eval()
        self.default_sample_size = unet_cfg.sample_size
        self.scheduler = EulerAncestralDiscrete() if self.is_turbo else EulerDiscrete()
        self.vae = vae
        if vae is None and build_vae:
            self.vae = synthetic_vae_init_(VAEDecoder(vae_cfg or SDXL_VAE), seed=seed + 1).eval()

    def encode_prompt(self, prompt, negative_prompt, do_cfg, dtype=torch.float16):
        """4-tuple (prompt_embeds, negative_prompt_embeds, pooled, negative_pooled);
        negatives are None without CFG (diffusers_holder.py:80-96)."""
        c = self.unet_cfg
        pe, pp = synthetic_text_embedding(prompt, c.cross_attention_dim, c.pooled_dim, dtype)
        if not do_cfg:
            return pe, None, pp, None
        neg = negative_prompt[0] if isinstance(negative_prompt, (list, tuple)) else (negative_prompt or "")
        ne, np_ = synthetic_text_embedding("<neg>" + neg, c.cross_attention_dim, c.pooled_dim, dtype)
        return pe, ne, pp, np_

    def add_time_ids(self, dtype=torch.float16):
        hw = self.default_sample_size * self.vae_scale_factor
        return torch.tensor([[hw, hw, 0, 0, hw, hw]], dtype=dtype)

    def prepare_latents(self, h_lat, w_lat, seed, dtype=torch.float16):
        g = torch.Generator().manual_seed(int(seed))
        lat = torch.randn(1, self.unet_cfg.in_channels, h_lat, w_lat, generator=g).to(dtype)
        return lat * self.scheduler.init_noise_sigma.to(dtype)
