"""SyntheticSDXLPipe: the ``pipe`` object handed to BlendingEngine when no diffusers
checkpoint is available (there is no network here): SDXL-shaped random-init weights
generated on the device, synthetic prompt embeddings, Euler scheduler tables.

It supplies what the reference's holder reads off a StableDiffusionXLPipeline
(latentblending/diffusers_holder.py): ``_name_or_path``,
``unet`` config, ``vae_scale_factor``, ``default_sample_size``, ``scheduler``,
``encode_prompt`` and the weights.  Bench-definition recipes:
  * weights: uniform(-1/sqrt(fan_in), 1/sqrt(fan_in)); residual-branch output
    projections (resnet conv2, attention to_out, FF out, transformer proj_out)
    scaled by 0.1; norm gains ~ 1 +- 0.1;
  * prompt embeddings: 0.5*N(0,1) [1,77,ctx] and N(0,1) [1,pooled] from a CPU
    generator seeded with crc32(prompt).
"""
import math
import zlib
from collections import OrderedDict

import torch

from .schedulers import EulerTables
from .unet import UNetConfig

SDXL_BASE = UNetConfig()
SDXL_TURBO = UNetConfig(sample_size=64)
VAE_CHANNELS = (128, 256, 512, 512)


def unet_param_shapes(cfg: UNetConfig):
    """diffusers state_dict names -> shapes of the SDXL UNet2DConditionModel."""
    P = OrderedDict()
    ch, T, X = list(cfg.block_out_channels), cfg.time_embed_dim, cfg.cross_attention_dim

    def lin(n, i, o, bias=True):
        P[n + ".weight"] = (o, i)
        if bias:
            P[n + ".bias"] = (o,)

    def conv(n, i, o, k):
        P[n + ".weight"] = (o, i, k, k)
        P[n + ".bias"] = (o,)

    def norm(n, c):
        P[n + ".weight"] = (c,)
        P[n + ".bias"] = (c,)

    def resnet(n, i, o):
        norm(n + ".norm1", i)
        conv(n + ".conv1", i, o, 3)
        lin(n + ".time_emb_proj", T, o)
        norm(n + ".norm2", o)
        conv(n + ".conv2", o, o, 3)
        if i != o:
            conv(n + ".conv_shortcut", i, o, 1)

    def transformer(n, c, depth):
        norm(n + ".norm", c)
        lin(n + ".proj_in", c, c)
        for d in range(depth):
            t = f"{n}.transformer_blocks.{d}"
            norm(t + ".norm1", c)
            for a, kdim in (("attn1", c), ("attn2", X)):
                lin(f"{t}.{a}.to_q", c, c, False)
                lin(f"{t}.{a}.to_k", kdim, c, False)
                lin(f"{t}.{a}.to_v", kdim, c, False)
                lin(f"{t}.{a}.to_out.0", c, c)
                if a == "attn1":
                    norm(t + ".norm2", c)
            norm(t + ".norm3", c)
            lin(t + ".ff.net.0.proj", c, 8 * c)
            lin(t + ".ff.net.2", 4 * c, c)
        lin(n + ".proj_out", c, c)

    conv("conv_in", cfg.in_channels, ch[0], 3)
    lin("time_embedding.linear_1", ch[0], T)
    lin("time_embedding.linear_2", T, T)
    lin("add_embedding.linear_1", cfg.add_in_dim, T)
    lin("add_embedding.linear_2", T, T)
    cout = ch[0]
    for i, c in enumerate(ch):
        cin, cout = cout, c
        for j in range(cfg.layers_per_block):
            resnet(f"down_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout)
            if cfg.transformer_layers[i]:
                transformer(f"down_blocks.{i}.attentions.{j}", cout, cfg.transformer_layers[i])
        if i < len(ch) - 1:
            conv(f"down_blocks.{i}.downsamplers.0.conv", cout, cout, 3)
    resnet("mid_block.resnets.0", ch[-1], ch[-1])
    transformer("mid_block.attentions.0", ch[-1], cfg.transformer_layers[-1])
    resnet("mid_block.resnets.1", ch[-1], ch[-1])
    rev, rdepth = list(reversed(ch)), list(reversed(cfg.transformer_layers))
    cout = rev[0]
    for i, c in enumerate(rev):
        cprev, cout = cout, c
        cin = rev[min(i + 1, len(ch) - 1)]
        n = cfg.layers_per_block + 1
        for j in range(n):
            skip_c = cin if j == n - 1 else cout
            res_in = cprev if j == 0 else cout
            resnet(f"up_blocks.{i}.resnets.{j}", res_in + skip_c, cout)
            if rdepth[i]:
                transformer(f"up_blocks.{i}.attentions.{j}", cout, rdepth[i])
        if i < len(ch) - 1:
            conv(f"up_blocks.{i}.upsamplers.0.conv", cout, cout, 3)
    norm("conv_norm_out", ch[0])
    conv("conv_out", ch[0], cfg.out_channels, 3)
    return P


def vae_param_shapes(channels=VAE_CHANNELS, latent=4):
    P = OrderedDict()
    ch = list(reversed(channels))

    def conv(n, i, o, k):
        P[n + ".weight"] = (o, i, k, k)
        P[n + ".bias"] = (o,)

    def norm(n, c):
        P[n + ".weight"] = (c,)
        P[n + ".bias"] = (c,)

    def resnet(n, i, o):
        norm(n + ".norm1", i)
        conv(n + ".conv1", i, o, 3)
        norm(n + ".norm2", o)
        conv(n + ".conv2", o, o, 3)
        if i != o:
            conv(n + ".conv_shortcut", i, o, 1)

    conv("post_quant_conv", latent, latent, 1)
    conv("conv_in", latent, ch[0], 3)
    resnet("mid_block.resnets.0", ch[0], ch[0])
    norm("mid_block.attentions.0.group_norm", ch[0])
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        P[f"mid_block.attentions.0.{n}.weight"] = (ch[0], ch[0])
        P[f"mid_block.attentions.0.{n}.bias"] = (ch[0],)
    resnet("mid_block.resnets.1", ch[0], ch[0])
    cout = ch[0]
    for i, c in enumerate(ch):
        cin, cout = cout, c
        for j in range(3):
            resnet(f"up_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout)
        if i < len(ch) - 1:
            conv(f"up_blocks.{i}.upsamplers.0.conv", cout, cout, 3)
    norm("conv_norm_out", ch[-1])
    conv("conv_out", ch[-1], 3, 3)
    return P


_DAMPED = ("conv2.weight", "conv2.bias", "to_out.0.weight", "to_out.0.bias", "ff.net.2.weight", "ff.net.2.bias",
           "proj_out.weight", "proj_out.bias")


def random_state_dict(shapes, seed, device, damp=0.1, dtype=torch.float16):
    g = torch.Generator(device=device).manual_seed(seed)
    sd = OrderedDict()
    for name, shape in shapes.items():
        if len(shape) >= 2:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = (torch.rand(shape, generator=g, device=device) * 2 - 1) * (1.0 / math.sqrt(fan_in))
        elif "norm" in name and name.endswith("weight"):
            t = 1.0 + 0.1 * (torch.rand(shape, generator=g, device=device) * 2 - 1)
        elif "norm" in name:
            t = 0.05 * (torch.rand(shape, generator=g, device=device) * 2 - 1)
        else:
            t = 0.02 * (torch.rand(shape, generator=g, device=device) * 2 - 1)
        if any(name.endswith(s) for s in _DAMPED):
            t = t * damp
        sd[name] = t.to(dtype)
    return sd


def prompt_seed(prompt: str) -> int:
    return zlib.crc32(prompt.encode("utf-8")) & 0x7FFFFFFF


class SyntheticSDXLPipe:
    vae_scale_factor = 8

    def __init__(self, name="stabilityai/stable-diffusion-xl-base-1.0", device="cuda:0", unet_cfg: UNetConfig = None,
                 seed=0, unet_state_dict=None, vae_state_dict=None, vae_channels=VAE_CHANNELS,
                 lpips_state_dict=None):
        self._name_or_path = name
        self.device = torch.device(device)
        self._execution_device = self.device
        turbo = "turbo" in name
        self.unet_cfg = unet_cfg or (SDXL_TURBO if turbo else SDXL_BASE)
        self.default_sample_size = self.unet_cfg.sample_size
        self.scheduler = EulerTables("euler_ancestral" if turbo else "euler")
        self.unet_state_dict = unet_state_dict if unet_state_dict is not None else \
            random_state_dict(unet_param_shapes(self.unet_cfg), seed, self.device)
        self.vae_channels = tuple(vae_channels)
        self.vae_state_dict = vae_state_dict if vae_state_dict is not None else \
            random_state_dict(vae_param_shapes(vae_channels), seed + 1, self.device, damp=0.3)
        self.vae_scaling_factor = 0.13025
        self.lpips_state_dict = lpips_state_dict
        self.h2d_bytes = 0        # bytes of conditioning copied host->device (bench e2e)

    def encode_prompt(self, prompt, negative_prompt=None, do_classifier_free_guidance=True, dtype=torch.float16):
        """Synthetic stand-in for StableDiffusionXLPipeline.encode_prompt (diffusers_holder.py:81-95)."""
        c = self.unet_cfg

        def emb(text):
            g = torch.Generator().manual_seed(prompt_seed(text))
            e = (torch.randn(1, 77, c.cross_attention_dim, generator=g) * 0.5).to(dtype)
            p = torch.randn(1, c.pooled_dim, generator=g).to(dtype)
            e, p = e.pin_memory(), p.pin_memory()
            self.h2d_bytes += e.numel() * e.element_size() + p.numel() * p.element_size()
            return e.to(self.device, non_blocking=True), p.to(self.device, non_blocking=True)

        pe, pp = emb(prompt)
        if not do_classifier_free_guidance:
            return pe, None, pp, None
        neg = negative_prompt[0] if isinstance(negative_prompt, (list, tuple)) else (negative_prompt or "")
        ne, npool = emb("<neg>" + neg)
        return pe, ne, pp, npool
