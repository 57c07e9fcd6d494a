"""Oracle: SDXL UNet2DConditionModel forward in plain PyTorch (test infrastructure).

Call site in the reference: latentblending/diffusers_holder.py:336-344.
The code behind it is diffusers==0.25.0 (models/unet_2d_condition.py,
unet_2d_blocks.py, resnet.py, transformer_2d.py, attention.py,
attention_processor.py, embeddings.py) -- NOT vendored.  This restates the
published architecture of stabilityai/stable-diffusion-xl-base-1.0:
block_out_channels [320,640,1280], layers_per_block 2,
transformer_layers_per_block [1,2,10] (block 0 is a plain DownBlock2D),
head dim 64, cross_attention_dim 2048, addition_embed_type text_time (256-d
sinusoids of 6 time ids + 1280-d pooled text -> 2816), use_linear_projection,
GroupNorm(32) eps 1e-5 in resnets / 1e-6 at the transformer entry, LayerNorm
eps 1e-5, exact-erf GELU inside GEGLU, flip_sin_to_cos sinusoids.

Parameter names mirror the diffusers state_dict keys so that a real SDXL
checkpoint's ``unet.state_dict()`` loads here and into the CUDA executor alike.
Structural pin: the full config has exactly 2 567 463 684 parameters
(tests/test_oracle_unet.py).
"""
import math
from dataclasses import dataclass, field
from typing import Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280)
    layers_per_block: int = 2
    transformer_layers: Tuple[int, ...] = (0, 2, 10)   # 0 == block without attention
    head_dim: int = 64
    cross_attention_dim: int = 2048
    addition_time_embed_dim: int = 256
    pooled_dim: int = 1280
    norm_num_groups: int = 32
    sample_size: int = 128
    time_cond_proj_dim: object = None

    @property
    def time_embed_dim(self):
        return self.block_out_channels[0] * 4

    @property
    def add_in_dim(self):   # projection_class_embeddings_input_dim (2816 for SDXL)
        return self.pooled_dim + 6 * self.addition_time_embed_dim


SDXL_BASE = UNetConfig()
SDXL_TURBO = UNetConfig(sample_size=64)


def tiny_config(**kw):
    """Same topology, small widths: used by parity tests that must finish in seconds."""
    base = dict(block_out_channels=(64, 128, 256), transformer_layers=(0, 1, 2), head_dim=64,
                cross_attention_dim=128, addition_time_embed_dim=32, pooled_dim=64, sample_size=16)
    base.update(kw)
    return UNetConfig(**base)


def sinusoid(t, dim):
    """diffusers get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0): [cos | sin]."""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half
    arg = t.reshape(-1, 1).float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(arg), torch.sin(arg)], dim=-1)


class TimestepEmbedding(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.linear_1 = nn.Linear(cin, cout)
        self.linear_2 = nn.Linear(cout, cout)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class ResnetBlock2D(nn.Module):
    def __init__(self, cin, cout, temb_dim, groups, eps=1e-5):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_dim, cout) if temb_dim else None
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb=None):
        h = self.conv1(F.silu(self.norm1(x)))
        if self.time_emb_proj is not None:
            h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Attention(nn.Module):
    def __init__(self, dim, ctx_dim, head_dim):
        super().__init__()
        self.heads = dim // head_dim
        self.to_q = nn.Linear(dim, dim, bias=False)
        self.to_k = nn.Linear(ctx_dim, dim, bias=False)
        self.to_v = nn.Linear(ctx_dim, dim, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Identity()])

    def forward(self, x, ctx=None):
        ctx = x if ctx is None else ctx
        B, S, C = x.shape
        h = self.heads
        q = self.to_q(x).view(B, S, h, C // h).transpose(1, 2)
        k = self.to_k(ctx).view(B, ctx.shape[1], h, C // h).transpose(1, 2)
        v = self.to_v(ctx).view(B, ctx.shape[1], h, C // h).transpose(1, 2)
        w = torch.softmax((q @ k.transpose(-1, -2)) * (C // h) ** -0.5, dim=-1)
        o = (w @ v).transpose(1, 2).reshape(B, S, C)
        return self.to_out[0](o)


class GEGLU(nn.Module):
    def __init__(self, dim, inner):
        super().__init__()
        self.proj = nn.Linear(dim, inner * 2)

    def forward(self, x):
        x, gate = self.proj(x).chunk(2, dim=-1)
        return x * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Identity(), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        return self.net[2](self.net[0](x))


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, ctx_dim, head_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, dim, head_dim)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, ctx_dim, head_dim)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, ctx):
        x = x + self.attn1(self.norm1(x))
        x = x + self.attn2(self.norm2(x), ctx)
        return x + self.ff(self.norm3(x))


class Transformer2DModel(nn.Module):
    def __init__(self, dim, depth, ctx_dim, head_dim, groups):
        super().__init__()
        self.norm = nn.GroupNorm(groups, dim, eps=1e-6)
        self.proj_in = nn.Linear(dim, dim)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(dim, ctx_dim, head_dim) for _ in range(depth)])
        self.proj_out = nn.Linear(dim, dim)

    def forward(self, x, ctx):
        B, C, H, W = x.shape
        h = self.norm(x).permute(0, 2, 3, 1).reshape(B, H * W, C)
        h = self.proj_in(h)
        for blk in self.transformer_blocks:
            h = blk(h, ctx)
        h = self.proj_out(h).reshape(B, H, W, C).permute(0, 3, 1, 2)
        return h + x


class Downsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock(nn.Module):
    def __init__(self, cfg, cin, cout, depth, add_down):
        super().__init__()
        g, T = cfg.norm_num_groups, cfg.time_embed_dim
        self.resnets = nn.ModuleList(
            [ResnetBlock2D(cin if i == 0 else cout, cout, T, g) for i in range(cfg.layers_per_block)])
        self.attentions = nn.ModuleList(
            [Transformer2DModel(cout, depth, cfg.cross_attention_dim, cfg.head_dim, g)
             for _ in range(cfg.layers_per_block)]) if depth else None
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_down else None

    def forward(self, x, temb, ctx):
        skips = []
        for i, res in enumerate(self.resnets):
            x = res(x, temb)
            if self.attentions is not None:
                x = self.attentions[i](x, ctx)
            skips.append(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            skips.append(x)
        return x, skips


class UpBlock(nn.Module):
    def __init__(self, cfg, cin, cout, cprev, depth, add_up):
        super().__init__()
        g, T = cfg.norm_num_groups, cfg.time_embed_dim
        n = cfg.layers_per_block + 1
        res = []
        for i in range(n):
            skip_c = cin if i == n - 1 else cout
            res_in = cprev if i == 0 else cout
            res.append(ResnetBlock2D(res_in + skip_c, cout, T, g))
        self.resnets = nn.ModuleList(res)
        self.attentions = nn.ModuleList(
            [Transformer2DModel(cout, depth, cfg.cross_attention_dim, cfg.head_dim, g)
             for _ in range(n)]) if depth else None
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None

    def forward(self, x, skips, temb, ctx):
        for i, res in enumerate(self.resnets):
            x = res(torch.cat([x, skips.pop()], dim=1), temb)
            if self.attentions is not None:
                x = self.attentions[i](x, ctx)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class MidBlock(nn.Module):
    def __init__(self, cfg, c, depth):
        super().__init__()
        g, T = cfg.norm_num_groups, cfg.time_embed_dim
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, T, g), ResnetBlock2D(c, c, T, g)])
        self.attentions = nn.ModuleList([Transformer2DModel(c, depth, cfg.cross_attention_dim, cfg.head_dim, g)])

    def forward(self, x, temb, ctx):
        x = self.resnets[0](x, temb)
        x = self.attentions[0](x, ctx)
        return self.resnets[1](x, temb)


class SDXLUNet(nn.Module):
    def __init__(self, cfg: UNetConfig = SDXL_BASE):
        super().__init__()
        self.cfg = cfg
        ch = cfg.block_out_channels
        T = cfg.time_embed_dim
        self.conv_in = nn.Conv2d(cfg.in_channels, ch[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(ch[0], T)
        self.add_embedding = TimestepEmbedding(cfg.add_in_dim, T)
        downs, cout = [], ch[0]
        for i, c in enumerate(ch):
            cin, cout = cout, c
            downs.append(DownBlock(cfg, cin, cout, cfg.transformer_layers[i], add_down=i < len(ch) - 1))
        self.down_blocks = nn.ModuleList(downs)
        self.mid_block = MidBlock(cfg, ch[-1], cfg.transformer_layers[-1])
        rev, rev_depth = list(reversed(ch)), list(reversed(cfg.transformer_layers))
        ups, cout = [], rev[0]
        for i, c in enumerate(rev):
            cprev, cout = cout, c
            cin = rev[min(i + 1, len(ch) - 1)]
            ups.append(UpBlock(cfg, cin, cout, cprev, rev_depth[i], add_up=i < len(ch) - 1))
        self.up_blocks = nn.ModuleList(ups)
        self.conv_norm_out = nn.GroupNorm(cfg.norm_num_groups, ch[0], eps=1e-5)
        self.conv_out = nn.Conv2d(ch[0], cfg.out_channels, 3, padding=1)

    def embed(self, t, text_embeds, time_ids):
        """emb = time_embedding(sinusoid(t)) + add_embedding([pooled | sinusoid(time_ids)])."""
        B = text_embeds.shape[0]
        cfg = self.cfg
        t = torch.as_tensor(t, dtype=torch.float32).reshape(-1).expand(B)
        temb = self.time_embedding(sinusoid(t, cfg.block_out_channels[0]))
        tid = sinusoid(time_ids.reshape(-1), cfg.addition_time_embed_dim).reshape(B, -1)
        aug = self.add_embedding(torch.cat([text_embeds.float(), tid], dim=-1))
        return temb + aug

    def forward(self, x, t, encoder_hidden_states, text_embeds, time_ids):
        emb = self.embed(t, text_embeds, time_ids)
        ctx = encoder_hidden_states
        x = self.conv_in(x)
        skips = [x]
        for blk in self.down_blocks:
            x, s = blk(x, emb, ctx)
            skips += s
        x = self.mid_block(x, emb, ctx)
        for blk in self.up_blocks:
            x = blk(x, skips, emb, ctx)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


RESIDUAL_DAMP = 0.1


def synthetic_init_(unet: nn.Module, seed=0, damp=RESIDUAL_DAMP):
    """The bench definition's random-init recipe: default
    PyTorch layer inits under ``torch.manual_seed(seed)``, then every
    residual-branch output projection (resnet conv2, attention to_out, FF out,
    transformer proj_out) is scaled by ``damp`` so activations stay finite in
    fp16 through 70 residual transformer blocks x 30 steps."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in unet.named_parameters():
            if p.dim() >= 2:
                fan_in = p[0].numel()
                bound = 1.0 / math.sqrt(fan_in)
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * bound)
            elif name.endswith("bias") and "norm" not in name:
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * 0.02)
            elif "norm" in name and name.endswith("weight"):
                p.copy_(1.0 + 0.1 * (torch.rand(p.shape, generator=g) * 2 - 1))
            elif "norm" in name and name.endswith("bias"):
                p.copy_(0.05 * (torch.rand(p.shape, generator=g) * 2 - 1))
        for name, p in unet.named_parameters():
            if any(name.endswith(s) for s in (
                    "conv2.weight", "conv2.bias", "to_out.0.weight", "to_out.0.bias",
                    "ff.net.2.weight", "ff.net.2.bias", "proj_out.weight", "proj_out.bias")):
                p.mul_(damp)
    return unet


def count_params(cfg: UNetConfig = SDXL_BASE):
    with torch.device("meta"):
        m = SDXLUNet(cfg)
    return sum(p.numel() for p in m.parameters())
