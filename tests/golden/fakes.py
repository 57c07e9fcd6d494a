"""Deterministic stand-ins shared by tests/golden/make_golden.py (which drives the
REFERENCE's BlendingEngine host logic with them) and by the tests (which drive
the oracle's and the product's engines with the very same objects).

FakeHolder exposes the 14-attribute ``dh`` seam the reference engine touches
 with cheap closed-form "denoising" on CPU so the whole tree
logic can run in milliseconds.
"""
import numpy as np
import torch


class FakeHolder:
    def __init__(self, turbo=False, h=8, w=8, device="cpu"):
        self.is_sdxl_turbo = turbo
        self.device = device
        self.pipe = None
        self.guidance_scale = 5.0
        self.num_inference_steps = 30
        self.negative_prompt = ""
        self.height_latent, self.width_latent = h, w
        self.height_img, self.width_img = h * 8, w * 8
        self.calls = []          # log of run_diffusion_sd_xl arguments

    def set_dimensions(self, size_output):
        pass

    def set_negative_prompt(self, neg):
        self.negative_prompt = neg

    def set_num_inference_steps(self, n):
        self.num_inference_steps = n

    def get_text_embedding(self, prompt):
        g = torch.Generator().manual_seed(sum(ord(c) for c in prompt) + 7)
        pe = torch.randn(1, 5, 16, generator=g).to(self.device)
        pp = torch.randn(1, 8, generator=g).to(self.device)
        if self.guidance_scale > 1:
            return pe, pe * 0.1, pp, pp * 0.1
        return pe, None, pp, None

    def get_noise(self, seed=420):
        g = torch.Generator().manual_seed(int(seed))
        return torch.randn(1, 4, self.height_latent, self.width_latent, generator=g).half().to(self.device)

    def latent2image(self, latents, output_type="pil"):
        x = latents.float()[0, :3]
        x = torch.nn.functional.interpolate(x[None], scale_factor=8, mode="nearest")[0]
        img = ((torch.tanh(x) * 0.5 + 0.5) * 255).round().clamp(0, 255).byte()
        return img.permute(1, 2, 0).cpu().numpy()

    def run_diffusion_sd_xl(self, text_embeddings, latents_start, idx_start=0,
                            list_latents_mixing=None, mixing_coeffs=0.0, return_image=False):
        N = self.num_inference_steps
        coeffs = list(mixing_coeffs) if isinstance(mixing_coeffs, list) else float(mixing_coeffs)
        self.calls.append(dict(idx_start=int(idx_start), coeffs=coeffs,
                               guidance=float(self.guidance_scale),
                               start_sum=float(latents_start.float().sum()),
                               cond_sum=float(text_embeddings[0].float().sum()),
                               n_mix_none=None if list_latents_mixing is None
                               else sum(1 for m in list_latents_mixing if m is None)))
        bias = float(text_embeddings[0].float().mean())
        lat = latents_start.clone()
        out = []
        for i in range(N):
            if i < idx_start:
                out.append(None)
                continue
            lat = (lat.float() * 0.93 + 0.05 * torch.sin(lat.float() * 3 + i) + 0.02 * bias).half()
            out.append(lat.clone())
        return out


def fake_similarity(img_a, img_b):
    a = np.asarray(img_a).astype(np.float64)
    b = np.asarray(img_b).astype(np.float64)
    return float(np.mean(np.abs(a - b)) / 255.0)
