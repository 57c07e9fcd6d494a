"""AutoencoderKL decoder on liblb200 (SURVEY.md section 8f "next #1").

Replaces ``pipe.vae.decode(latents / scaling_factor)`` + ``image_processor.postprocess`` inside
``DiffusersHolder.latent2image`` (latentblending/diffusers_holder.py:114-143; diffusers 0.25.0
autoencoder_kl.py / vae.py, un-vendored).  The reference runs the stock SDXL VAE in fp32
(force_upcast); here the decoder runs in fp16 storage / fp32 accumulation on the same tcgen05
implicit-GEMM conv, GroupNorm and sampler kernels as the UNet -- 10.5 TFLOP per 1024^2 frame.
The mid-block single-head attention (head dim 512, S = h*w) is three GEMMs around a row softmax:
scores = (Wq x)(Wk x)^T (1/sqrt(C) folded into Wq), P = softmax_rows(scores), out = P V with V^T
produced directly by a GEMM with swapped operands; the value bias is folded into the output bias
(softmax rows sum to one).  Weights use the diffusers state_dict names.
"""
import os

import torch

from . import _cabi
from ._cabi import ctx
from .unet import Program, pack_conv_out8


class VAEDecoderB200:
    def __init__(self, state_dict, channels, scaling_factor, device, groups=32):
        self.device = torch.device(device)
        self.dev_index = self.device.index or 0
        self.channels = tuple(channels)
        self.scaling_factor = scaling_factor
        self.groups = groups
        sd = state_dict
        dev = self.device

        def g(n):
            return sd[n].detach().to(device=dev, dtype=torch.float16).contiguous()

        def gf(n):
            return sd[n].detach().to(device=dev, dtype=torch.float32)

        def conv3(n):
            w = g(n + ".weight")
            return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()

        W = self.w = {}
        pq = gf("post_quant_conv.weight")
        C = pq.shape[0]
        W["prep.w"] = (pq.reshape(C, C) / scaling_factor).contiguous()
        W["prep.b"] = gf("post_quant_conv.bias").contiguous()
        W["conv_in.w"] = g("conv_in.weight").permute(2, 3, 1, 0).contiguous()
        W["conv_in.b"] = g("conv_in.bias")
        self.resnets = [k[: -len(".norm1.weight")] for k in sd if k.endswith(".norm1.weight")]
        for r in self.resnets:
            W[r + ".norm1.g"], W[r + ".norm1.b"] = g(r + ".norm1.weight"), g(r + ".norm1.bias")
            W[r + ".norm2.g"], W[r + ".norm2.b"] = g(r + ".norm2.weight"), g(r + ".norm2.bias")
            W[r + ".conv1.w"], W[r + ".conv1.b"] = conv3(r + ".conv1"), g(r + ".conv1.bias")
            w2, b2 = conv3(r + ".conv2"), g(r + ".conv2.bias")
            if (r + ".conv_shortcut.weight") in sd:
                ws = g(r + ".conv_shortcut.weight")
                w2 = torch.cat([w2, ws.reshape(ws.shape[0], -1)], 1).contiguous()
                b2 = (b2.float() + g(r + ".conv_shortcut.bias").float()).half()
                W[r + ".has_shortcut"] = True
            W[r + ".conv2.w"], W[r + ".conv2.b"] = w2, b2
        a = "mid_block.attentions.0"
        Cm = sd[a + ".to_q.weight"].shape[0]
        scale = Cm ** -0.5
        W["attn.norm.g"], W["attn.norm.b"] = g(a + ".group_norm.weight"), g(a + ".group_norm.bias")
        W["attn.qk.w"] = torch.cat([(gf(a + ".to_q.weight") * scale), gf(a + ".to_k.weight")], 0).half().contiguous()
        W["attn.qk.b"] = torch.cat([(gf(a + ".to_q.bias") * scale), gf(a + ".to_k.bias")], 0).half().contiguous()
        W["attn.v.w"] = g(a + ".to_v.weight")
        W["attn.out.w"] = g(a + ".to_out.0.weight")
        W["attn.out.b"] = (gf(a + ".to_out.0.bias") + gf(a + ".to_out.0.weight") @ gf(a + ".to_v.bias")).half().contiguous()
        for k in sd:
            if k.endswith("upsamplers.0.conv.weight"):
                nm = k[: -len(".weight")]
                W[nm + ".w"], W[nm + ".b"] = conv3(nm), g(nm + ".bias")
        W["norm_out.g"], W["norm_out.b"] = g("conv_norm_out.weight"), g("conv_norm_out.bias")
        W["conv_out.w"] = g("conv_out.weight").permute(0, 2, 3, 1).contiguous()
        W["conv_out.b"] = g("conv_out.bias")
        W["conv_out.w8"], W["conv_out.b8"] = pack_conv_out8(W["conv_out.w"], W["conv_out.b"])
        self._plans = {}
        self.nonfinite = torch.zeros(1, dtype=torch.int32, device=dev)
        self.decodes_since_check = 0

    def plan(self, h, w):
        if (h, w) not in self._plans:
            self._plans[(h, w)] = _VAELowering(self, h, w)
        return self._plans[(h, w)]

    @torch.no_grad()
    def decode_to_u8(self, latents):
        """latents [1,4,h,w] fp16 (CUDA) -> uint8 [8h,8w,3] frame on the device."""
        assert latents.is_cuda and latents.shape[0] == 1, "VAE decode: one CUDA latent at a time"
        _, _, h, w = latents.shape
        pl = self.plan(h, w)
        pl.z_in.copy_(latents)
        pl.prog.run()
        self.decodes_since_check += 1
        return pl.frame.clone()

    def overflow_count(self):
        """Non-finite pixels seen by the post-process kernel since the last call (device->host read: call it at a
        point that synchronises anyway).  The decoder stores fp16 where the reference upcasts the stock SDXL VAE to
        fp32 (diffusers_holder.py:128-133); with weights that overflow fp16 this is > 0 and the frames are invalid."""
        n = int(self.nonfinite.item())
        if n:
            self.nonfinite.zero_()
        self.decodes_since_check = 0
        return n

    def check_overflow(self):
        n = self.overflow_count()
        if n:
            raise _cabi.LB200Error(
                f"VAE decode produced {n} non-finite pixels: these VAE weights overflow fp16 (the reference upcasts the "
                "stock SDXL VAE to fp32, diffusers_holder.py:128-133); use the fp16-safe SDXL VAE weights "
                "(madebyollin/sdxl-vae-fp16-fix) with this backend")


class _VAELowering:
    def __init__(self, vae: VAEDecoderB200, h, w):
        Wt, dev, groups = vae.w, vae.device, vae.groups
        f16 = dict(dtype=torch.float16, device=dev)
        ch = list(reversed(vae.channels))            # e.g. [512, 512, 256, 128]
        B = 1
        P = self.prog = Program(vae.dev_index)
        self.z_in = torch.zeros(1, 4, h, w, **f16)
        H, W_ = 8 * h, 8 * w
        self.frame = torch.zeros(H, W_, 3, dtype=torch.uint8, device=dev)
        self.ws = torch.zeros(max(1 << 16, _cabi.load().lb_groupnorm_workspace_bytes(ctx(vae.dev_index), B, H * W_, groups)),
                              dtype=torch.uint8, device=dev)
        scratch = {}

        def sc(name, rows, cols):
            need = rows * cols
            if name not in scratch or scratch[name].numel() < need:
                scratch[name] = torch.empty(need, **f16)
            return scratch[name][:need].view(rows, cols)

        def resnet(rname, x, cin, cout, hh, ww, out):
            M = hh * ww
            n1 = sc("n1", M, cin)
            P.groupnorm(x, B, M, cin, groups, Wt[rname + ".norm1.g"], Wt[rname + ".norm1.b"], 1e-6, 1, n1, self.ws)
            h1 = sc("h1", M, cout)
            P.gemm(n1, Wt[rname + ".conv1.w"], cout, B, hh, ww, h1, taps=9, bias=Wt[rname + ".conv1.b"])
            n2 = sc("n2", M, cout)
            P.groupnorm(h1, B, M, cout, groups, Wt[rname + ".norm2.g"], Wt[rname + ".norm2.b"], 1e-6, 1, n2, self.ws)
            if Wt.get(rname + ".has_shortcut"):
                P.gemm(n2, Wt[rname + ".conv2.w"], cout, B, hh, ww, out, taps=9, a1=x, a1_c=cin, bias=Wt[rname + ".conv2.b"])
            else:
                P.gemm(n2, Wt[rname + ".conv2.w"], cout, B, hh, ww, out, taps=9, bias=Wt[rname + ".conv2.b"], res=x)

        z = torch.empty(1, 4, h, w, **f16)
        P.latent_prep(self.z_in, Wt["prep.w"], Wt["prep.b"], z)
        C0 = ch[0]
        S = h * w
        x = torch.empty(S, C0, **f16)
        P.conv_in(z, Wt["conv_in.w"], Wt["conv_in.b"], C0, x)
        x2 = torch.empty(S, C0, **f16)
        resnet("mid_block.resnets.0", x, C0, C0, h, w, x2)
        # mid-block attention (single head, dim C0)
        hn = sc("n1", S, C0)
        P.groupnorm(x2, B, S, C0, groups, Wt["attn.norm.g"], Wt["attn.norm.b"], 1e-6, 0, hn, self.ws)
        qk = torch.empty(S, 2 * C0, **f16)
        P.gemm(hn, Wt["attn.qk.w"], 2 * C0, 1, 1, S, qk, bias=Wt["attn.qk.b"])
        vT = torch.empty(C0, S, **f16)
        P.gemm(Wt["attn.v.w"], hn, S, 1, 1, C0, vT, static_w=False)                       # V^T = Wv hn^T
        scores = torch.empty(S, S, **f16)
        P.gemm(qk[:, :C0], qk[:, C0:], S, 1, 1, S, scores, static_w=False)               # (scaled q) k^T
        P.softmax_rows(scores, scores)
        att = sc("h1", S, C0)
        P.gemm(scores, vT, C0, 1, 1, S, att, static_w=False)                              # P V
        x3 = torch.empty(S, C0, **f16)
        P.gemm(att, Wt["attn.out.w"], C0, 1, 1, S, x3, bias=Wt["attn.out.b"], res=x2)
        x4 = torch.empty(S, C0, **f16)
        resnet("mid_block.resnets.1", x3, C0, C0, h, w, x4)
        x, cin, hh, ww = x4, C0, h, w
        ping = {}
        for i, cout in enumerate(ch):
            for j in range(3):
                out = torch.empty(hh * ww, cout, **f16) if (i, j) not in ping else ping[(i, j)]
                resnet(f"up_blocks.{i}.resnets.{j}", x, cin, cout, hh, ww, out)
                x, cin = out, cout
            nm = f"up_blocks.{i}.upsamplers.0.conv"
            if (nm + ".w") in Wt:
                up = torch.empty(4 * hh * ww, cout, **f16)
                P.upsample2x(x, B, hh, ww, cout, up)
                hh, ww = 2 * hh, 2 * ww
                nx = torch.empty(hh * ww, cout, **f16)
                P.gemm(up, Wt[nm + ".w"], cout, B, hh, ww, nx, taps=9, bias=Wt[nm + ".b"])
                x = nx
        no = sc("n1", hh * ww, cin)
        P.groupnorm(x, B, hh * ww, cin, groups, Wt["norm_out.g"], Wt["norm_out.b"], 1e-6, 1, no, self.ws)
        img = torch.empty(1, 3, hh, ww, **f16)
        if Wt.get("conv_out.w8") is not None and os.environ.get("LB_CONV_OUT_DIRECT") is None:
            # 14 % of a 1024^2 decode went into the direct 128 -> 3 kernel (2.07 ms, r02k); as an N = 8 GEMM it is ~0.2 ms
            P.conv_out_gemm(no, B, hh, ww, cin, Wt["conv_out.w8"], Wt["conv_out.b8"], 3, img, sc("h1", hh * ww, 8))
        else:
            P.conv_out(no, B, hh, ww, cin, Wt["conv_out.w"], Wt["conv_out.b"], 3, img)
        P.postprocess_u8(img, self.frame, vae.nonfinite)
        self._keep = (scratch, z, qk, vT, scores, x2, x3, x4, img)
        P.finalize()
