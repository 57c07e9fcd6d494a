"""Interim on-device implementations of the two SURVEY.md section 8(f) "next" rows that sit
inside the reference's timed region but outside the denoise hot path proper:

  next #1  VAE decode   (latentblending/diffusers_holder.py:114-143)
  next #2  LPIPS-Alex   (latentblending/blending_engine.py:744-758)

They are PLAIN PYTORCH ON THE CUDA DEVICE (cuDNN/cuBLAS library kernels), not
hand-written sm_100a code, and are labelled as such in DESIGN.md; they exist so
``run_transition`` is functionally complete end to end.  No CPU path: CUDA
tensors are required.  The native VAE decoder (latentblending_b200/vae.py),
when enabled, replaces ``TorchVAEDecoder``.
"""
import torch
import torch.nn.functional as F

_LP_SHIFT = (-0.030, -0.088, -0.188)
_LP_SCALE = (0.458, 0.448, 0.450)
_LP_CH = (64, 192, 384, 256, 256)


def _gn(x, sd, n, eps=1e-6, groups=32):
    return F.group_norm(x, groups, sd[n + ".weight"], sd[n + ".bias"], eps)


def _resnet(x, sd, n):
    h = F.conv2d(F.silu(_gn(x, sd, n + ".norm1")), sd[n + ".conv1.weight"], sd[n + ".conv1.bias"], padding=1)
    h = F.conv2d(F.silu(_gn(h, sd, n + ".norm2")), sd[n + ".conv2.weight"], sd[n + ".conv2.bias"], padding=1)
    if (n + ".conv_shortcut.weight") in sd:
        x = F.conv2d(x, sd[n + ".conv_shortcut.weight"], sd[n + ".conv_shortcut.bias"])
    return x + h


class TorchVAEDecoder:
    """AutoencoderKL decoder (SDXL VAE topology) as functional PyTorch over a diffusers-named state dict."""

    def __init__(self, state_dict, device, n_up_blocks=4, scaling_factor=0.13025, dtype=torch.float16):
        self.sd = {k: v.detach().to(device=device, dtype=dtype) for k, v in state_dict.items()}
        self.n_up = n_up_blocks
        self.scaling_factor = scaling_factor
        self.dtype = dtype

    @torch.no_grad()
    def decode(self, latents):
        """[B,4,h,w] latents -> [B,3,8h,8w] image in [-1,1]-ish (pre-postprocess)."""
        assert latents.is_cuda, "VAE decode needs CUDA tensors (no CPU fallback)"
        sd = self.sd
        z = latents.to(self.dtype) / self.scaling_factor
        x = F.conv2d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
        x = F.conv2d(x, sd["conv_in.weight"], sd["conv_in.bias"], padding=1)
        x = _resnet(x, sd, "mid_block.resnets.0")
        a = "mid_block.attentions.0"
        B, C, H, W = x.shape
        h = _gn(x, sd, a + ".group_norm").reshape(B, C, H * W).transpose(1, 2)
        q = F.linear(h, sd[a + ".to_q.weight"], sd[a + ".to_q.bias"])
        k = F.linear(h, sd[a + ".to_k.weight"], sd[a + ".to_k.bias"])
        v = F.linear(h, sd[a + ".to_v.weight"], sd[a + ".to_v.bias"])
        o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]
        o = F.linear(o, sd[a + ".to_out.0.weight"], sd[a + ".to_out.0.bias"])
        x = x + o.transpose(1, 2).reshape(B, C, H, W)
        x = _resnet(x, sd, "mid_block.resnets.1")
        for i in range(self.n_up):
            for j in range(3):
                x = _resnet(x, sd, f"up_blocks.{i}.resnets.{j}")
            u = f"up_blocks.{i}.upsamplers.0.conv"
            if (u + ".weight") in sd:
                x = F.interpolate(x, scale_factor=2.0, mode="nearest")
                x = F.conv2d(x, sd[u + ".weight"], sd[u + ".bias"], padding=1)
        x = F.silu(_gn(x, sd, "conv_norm_out"))
        return F.conv2d(x, sd["conv_out.weight"], sd["conv_out.bias"], padding=1)


def postprocess_uint8(image):
    """VaeImageProcessor.postprocess: (x/2+0.5).clamp(0,1) -> NHWC -> *255 round -> uint8 (device tensor)."""
    image = (image.float() / 2 + 0.5).clamp(0, 1)
    return (image.permute(0, 2, 3, 1) * 255).round().to(torch.uint8)


def lpips_random_state_dict(seed, device):
    g = torch.Generator(device=device).manual_seed(seed)
    sd = {}
    spec = [(3, 64, 11), (64, 192, 5), (192, 384, 3), (384, 256, 3), (256, 256, 3)]
    for i, (ci, co, k) in enumerate(spec):
        fan_in = ci * k * k
        sd[f"convs.{i}.weight"] = torch.randn(co, ci, k, k, generator=g, device=device) * (2.0 / fan_in) ** 0.5
        sd[f"convs.{i}.bias"] = torch.zeros(co, device=device)
        sd[f"lins.{i}.weight"] = torch.rand(1, co, 1, 1, generator=g, device=device) / co
    return sd


class TorchLPIPSAlex:
    """LPIPS v0.1 / AlexNet distance on device frames (lpips==0.1.4 definition)."""

    def __init__(self, state_dict, device):
        self.sd = {k: v.to(device=device, dtype=torch.float32) for k, v in state_dict.items()}
        self.shift = torch.tensor(_LP_SHIFT, device=device).view(1, 3, 1, 1)
        self.scale = torch.tensor(_LP_SCALE, device=device).view(1, 3, 1, 1)

    def _features(self, x):
        sd = self.sd
        x = (x - self.shift) / self.scale
        taps = []
        strides, pads = (4, 1, 1, 1, 1), (2, 2, 1, 1, 1)
        for i in range(5):
            x = F.relu(F.conv2d(x, sd[f"convs.{i}.weight"], sd[f"convs.{i}.bias"], stride=strides[i], padding=pads[i]))
            taps.append(x)
            if i in (0, 1):
                x = F.max_pool2d(x, 3, 2)
        return taps

    @torch.no_grad()
    def distance(self, frame_a_u8, frame_b_u8):
        """uint8 HxWx3 device frames -> python float (blending_engine.py:750-758)."""
        assert frame_a_u8.is_cuda and frame_b_u8.is_cuda, "LPIPS needs CUDA tensors (no CPU fallback)"

        def prep(f):
            return (2 * f.float() / 255.0 - 1).permute(2, 0, 1).unsqueeze(0)
        total = 0.0
        for fa, fb, i in zip(self._features(prep(frame_a_u8)), self._features(prep(frame_b_u8)), range(5)):
            na = fa / (fa.pow(2).sum(1, keepdim=True).sqrt() + 1e-10)
            nb = fb / (fb.pow(2).sum(1, keepdim=True).sqrt() + 1e-10)
            total = total + F.conv2d((na - nb) ** 2, self.sd[f"lins.{i}.weight"]).mean(dim=(2, 3), keepdim=True)
        return float(total[0, 0, 0, 0])
