"""GPU numerics for the remaining UNet kernels (attention, GroupNorm/LayerNorm, embeddings,
boundary convs, samplers) against plain PyTorch fp32 references of the same ops."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rand(*shape, seed=0, s=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * s).half()


def _close(out, ref, rtol=2e-3, atol=1e-3):
    out, ref = out.float(), ref.float()
    scale = ref.abs().max().item() + 1e-6
    err = (out - ref).abs().max().item()
    assert err <= rtol * scale + atol, f"max err {err} vs scale {scale}"


@pytest.mark.parametrize("B,heads,S", [(2, 10, 4096), (2, 20, 1024), (1, 2, 256), (2, 4, 64), (1, 1, 16), (1, 3, 200)])
def test_self_attention_fused_qkv(B, heads, S):
    from latentblending_b200 import ops
    C = heads * 64
    qkv = _rand(B * S, 3 * C, seed=1)
    out = torch.zeros(B * S, C, dtype=torch.float16, device="cuda")
    ops.attention(qkv, qkv, qkv, out, B, heads, S, S, q_col0=0, k_col0=C, v_col0=2 * C)
    q, k, v = [t.float().view(B, S, heads, 64).transpose(1, 2) for t in qkv.view(B, S, 3 * C).split(C, dim=-1)]
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * S, C)
    _close(out, ref, rtol=4e-3)
    assert ops.error_flag() == 0


@pytest.mark.parametrize("B,heads,S", [(2, 20, 1024), (2, 10, 4096), (2, 2, 64)])
def test_cross_attention_77(B, heads, S):
    from latentblending_b200 import ops
    C = heads * 64
    q = _rand(B * S, C, seed=2)
    kv = _rand(B * 77, 2 * C, seed=3)
    out = torch.zeros(B * S, C, dtype=torch.float16, device="cuda")
    ops.attention(q, kv, kv, out, B, heads, S, 77, k_col0=0, v_col0=C)
    qf = q.float().view(B, S, heads, 64).transpose(1, 2)
    kf, vf = [t.float().view(B, 77, heads, 64).transpose(1, 2) for t in kv.view(B, 77, 2 * C).split(C, dim=-1)]
    ref = F.scaled_dot_product_attention(qf, kf, vf).transpose(1, 2).reshape(B * S, C)
    _close(out, ref, rtol=4e-3)


@pytest.mark.parametrize("B,HW,C,silu", [(2, 128 * 128, 320, 1), (2, 64 * 64, 960, 1), (2, 32 * 32, 2560, 1),
                                         (2, 32 * 32, 1280, 0), (1, 16, 64, 1), (3, 100, 128, 0)])
def test_groupnorm(B, HW, C, silu):
    from latentblending_b200 import ops
    x = _rand(B * HW, C, seed=4, s=2.0) + 0.5
    g, b = _rand(C, seed=5) * 0.1 + 1.0, _rand(C, seed=6) * 0.1
    eps = 1e-5 if silu else 1e-6
    out = ops.groupnorm(x, B, HW, C, 32, g, b, eps, silu)
    ref = F.group_norm(x.float().view(B, HW, C).transpose(1, 2), 32, g.float(), b.float(), eps)
    if silu:
        ref = F.silu(ref)
    _close(out.view(B, HW, C), ref.transpose(1, 2), rtol=2e-3, atol=2e-3)


def test_groupnorm_strided_concat_buffer():
    from latentblending_b200 import ops
    B, HW, C = 2, 1024, 640
    buf = _rand(B * HW, C + 320, seed=7)
    x = buf[:, :C]
    g, b = _rand(C, seed=8) * 0.1 + 1.0, _rand(C, seed=9) * 0.1
    out = ops.groupnorm(x, B, HW, C, 32, g, b, 1e-5, 1)
    ref = F.silu(F.group_norm(x.float().reshape(B, HW, C).transpose(1, 2), 32, g.float(), b.float(), 1e-5))
    _close(out.view(B, HW, C), ref.transpose(1, 2), atol=2e-3)


@pytest.mark.parametrize("rows,C", [(2048, 1280), (8192, 640), (77, 128), (5, 2048)])
def test_layernorm(rows, C):
    from latentblending_b200 import ops
    x = _rand(rows, C, seed=10, s=3.0) + 1.0
    g, b = _rand(C, seed=11) * 0.1 + 1.0, _rand(C, seed=12) * 0.1
    out = ops.layernorm(x, g, b, 1e-5)
    _close(out, F.layer_norm(x.float(), (C,), g.float(), b.float(), 1e-5), atol=2e-3)


def test_embed_inputs_and_small_linears():
    from latentblending_b200 import ops
    B, dim_t, dim_a, pooled = 2, 320, 256, 1280
    text = _rand(B, pooled, seed=13)
    tids = torch.tensor([[1024., 1024, 0, 0, 1024, 1024]] * B, device="cuda").half()
    for t in (958.0, 499.0, 1.0):
        temb_in, add_in = ops.embed_inputs(t, text, tids, dim_t, dim_a)

        def sinus(v, dim):
            half = dim // 2
            ex = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device="cuda") / half)
            a = v.reshape(-1, 1).float() * ex[None]
            return torch.cat([torch.cos(a), torch.sin(a)], -1)
        _close(temb_in, sinus(torch.full((B,), t, device="cuda"), dim_t), atol=2e-3)
        ref_add = torch.cat([text.float(), sinus(tids.reshape(-1), dim_a).reshape(B, -1)], -1)
        _close(add_in, ref_add, atol=2e-3)
    x = _rand(B, 2816, seed=14)
    w, bias = _rand(1280, 2816, seed=15, s=2816 ** -0.5), _rand(1280, seed=16)
    add = _rand(B, 1280, seed=17)
    out = ops.linear_small(x, w, bias, act_out=1)
    _close(out, F.silu(x.float() @ w.float().t() + bias.float()))
    out = ops.linear_small(x, w, bias, addend=add, act_in=1)
    _close(out, F.silu(x.float()).half().float() @ w.float().t() + bias.float() + add.float())
    x16 = _rand(16, 320, seed=18)
    w2 = _rand(77, 320, seed=19, s=320 ** -0.5)
    _close(ops.linear_small(x16, w2), x16.float() @ w2.float().t())


@pytest.mark.parametrize("B,H,W,C0", [(2, 128, 128, 320), (1, 16, 16, 64), (2, 9, 7, 64)])
def test_conv_in_out(B, H, W, C0):
    from latentblending_b200 import ops
    x = _rand(B, 4, H, W, seed=20)
    w = _rand(C0, 4, 3, 3, seed=21, s=1 / 6)
    b = _rand(C0, seed=22)
    wp = w.permute(2, 3, 1, 0).contiguous()                  # [ky][kx][cin][Cout]
    out = ops.conv_in(x, wp, b, C0)
    ref = F.conv2d(x.float(), w.float(), b.float(), padding=1)
    _close(out.view(B, H, W, C0), ref.permute(0, 2, 3, 1))
    h = _rand(B * H * W, C0, seed=23)
    wo = _rand(4, C0, 3, 3, seed=24, s=(9 * C0) ** -0.5)
    bo = _rand(4, seed=25)
    wop = wo.permute(0, 2, 3, 1).contiguous()                # [co][ky][kx][Cin]
    eps = ops.conv_out(h, B, H, W, C0, wop, bo, 4)
    ref = F.conv2d(h.float().view(B, H, W, C0).permute(0, 3, 1, 2), wo.float(), bo.float(), padding=1)
    _close(eps, ref)


def test_upsample_and_im2col_downsample_conv():
    from latentblending_b200 import ops
    B, H, W, C = 2, 16, 16, 128
    x = _rand(B * H * W, C, seed=26)
    up = ops.upsample2x(x, B, H, W, C)
    ref = F.interpolate(x.float().view(B, H, W, C).permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")
    assert torch.equal(up.view(B, 2 * H, 2 * W, C).float(), ref.permute(0, 2, 3, 1))
    w = _rand(C, C, 3, 3, seed=27, s=(9 * C) ** -0.5)
    bias = _rand(C, seed=28)
    cols = ops.im2col_s2(x, B, H, W, C)
    wp = w.permute(0, 2, 3, 1).reshape(C, 9 * C).contiguous()
    out = ops.gemm(cols, wp, C, 1, 1, B * (H // 2) * (W // 2), bias=bias)
    ref = F.conv2d(x.float().view(B, H, W, C).permute(0, 3, 1, 2), w.float(), bias.float(), stride=2, padding=1)
    _close(out.view(B, H // 2, W // 2, C), ref.permute(0, 2, 3, 1))
