"""Oracle: AutoencoderKL decoder + image post-processing (test infrastructure).

Call site: latentblending/diffusers_holder.py:114-143 (``latent2image``):
``vae.decode(latents / scaling_factor)`` in fp32 (force_upcast), then
``image_processor.postprocess`` -> PIL.  Code behind it: diffusers==0.25.0
``models/autoencoder_kl.py`` / ``vae.py`` (NOT vendored).  Restated from the
published SDXL VAE config: latent 4ch, block_out_channels [128,256,512,512],
layers_per_block 2 (decoder uses 3 resnets per up block), GroupNorm(32) eps
1e-6, one single-head attention in the mid block, scaling_factor 0.13025.
"""
from dataclasses import dataclass
from typing import Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .sdxl_unet import ResnetBlock2D, Upsample2D


@dataclass
class VAEConfig:
    latent_channels: int = 4
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    scaling_factor: float = 0.13025
    force_upcast: bool = True


SDXL_VAE = VAEConfig()


def tiny_vae_config():
    return VAEConfig(block_out_channels=(32, 32, 64, 64))


class VAEAttention(nn.Module):
    def __init__(self, c, groups):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, c, eps=1e-6)
        self.to_q = nn.Linear(c, c)
        self.to_k = nn.Linear(c, c)
        self.to_v = nn.Linear(c, c)
        self.to_out = nn.ModuleList([nn.Linear(c, c), nn.Identity()])

    def forward(self, x):
        B, C, H, W = x.shape
        h = self.group_norm(x).reshape(B, C, H * W).transpose(1, 2)
        q, k, v = self.to_q(h), self.to_k(h), self.to_v(h)
        w = torch.softmax(q @ k.transpose(1, 2) * C ** -0.5, dim=-1)
        o = self.to_out[0](w @ v).transpose(1, 2).reshape(B, C, H, W)
        return o + x


class VAEUpBlock(nn.Module):
    def __init__(self, cin, cout, n, groups, add_up):
        super().__init__()
        self.resnets = nn.ModuleList(
            [ResnetBlock2D(cin if i == 0 else cout, cout, 0, groups, eps=1e-6) for i in range(n)])
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class VAEMid(nn.Module):
    def __init__(self, c, groups):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, 0, groups, eps=1e-6) for _ in range(2)])
        self.attentions = nn.ModuleList([VAEAttention(c, groups)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class VAEDecoder(nn.Module):
    def __init__(self, cfg: VAEConfig = SDXL_VAE):
        super().__init__()
        self.cfg = cfg
        ch = list(reversed(cfg.block_out_channels))
        g = cfg.norm_num_groups
        self.post_quant_conv = nn.Conv2d(cfg.latent_channels, cfg.latent_channels, 1)
        self.conv_in = nn.Conv2d(cfg.latent_channels, ch[0], 3, padding=1)
        self.mid_block = VAEMid(ch[0], g)
        ups, cout = [], ch[0]
        for i, c in enumerate(ch):
            cin, cout = cout, c
            ups.append(VAEUpBlock(cin, cout, cfg.layers_per_block + 1, g, add_up=i < len(ch) - 1))
        self.up_blocks = nn.ModuleList(ups)
        self.conv_norm_out = nn.GroupNorm(g, ch[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(ch[-1], 3, 3, padding=1)

    def forward(self, z):
        x = self.conv_in(self.post_quant_conv(z))
        x = self.mid_block(x)
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


def postprocess_to_uint8(image):
    """VaeImageProcessor.postprocess(output_type='pil'): denormalize, clamp,
    NHWC, *255 round -> uint8 (diffusers_holder.py:141)."""
    image = (image / 2 + 0.5).clamp(0, 1)
    arr = image.detach().cpu().permute(0, 2, 3, 1).float().numpy()
    return (arr * 255).round().astype("uint8")


def latent2image_np(vae: VAEDecoder, latents):
    """diffusers_holder.py:129-141 with output as uint8 HxWx3 array."""
    z = latents.to(torch.float32) / vae.cfg.scaling_factor
    return postprocess_to_uint8(vae(z))[0]


def synthetic_vae_init_(vae, seed=1):
    from .sdxl_unet import synthetic_init_
    return synthetic_init_(vae, seed=seed, damp=0.3)
