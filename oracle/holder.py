"""Oracle: the denoise operator (test infrastructure).

Restates latentblending/diffusers_holder.py:
  :51-66   set_num_inference_steps / set_dimensions
  :79-96   get_text_embedding (CFG decided from guidance_scale > 1)
  :98-111  get_noise
  :146-156 prepare_mixing
  :172-366 run_diffusion_sd_xl  -- loop order per step:
           skip(None) -> inject start -> crossfeed slerp -> CFG cat ->
           scale_model_input -> UNet -> CFG combine -> scheduler.step -> clone
The UNet runs in fp32 on the (fp16-rounded) model input and its output is
rounded back to the latent dtype, i.e. an ideal fp16-storage / fp32-accumulate
pipeline; every elementwise op around it runs in the latent dtype exactly as
the reference's torch ops do.
"""
import numpy as np
import torch

from .mixing import interpolate_spherical
from .vae import latent2image_np


class OracleHolder:
    def __init__(self, pipe, latent_dtype=torch.float16):
        self.pipe = pipe
        self.device = "cpu"
        self.dtype = latent_dtype
        self.negative_prompt = ""                      # diffusers_holder.py:23
        self.guidance_scale = 5.0                      # :24
        self.num_inference_steps = 30                  # :25
        self.is_sdxl_turbo = "turbo" in pipe._name_or_path          # :48
        self.pipe.scheduler.set_timesteps(self.num_inference_steps)  # :42
        s = pipe.unet_cfg.sample_size
        self.width_latent = self.height_latent = s                    # :32-33
        self.width_img = self.height_img = s * pipe.vae_scale_factor  # :34-35
        self.noise_fn = None   # ancestral-step noise injection hook: noise_fn(i, shape) -> tensor

    # -- configuration ----------------------------------------------------
    def set_num_inference_steps(self, n):
        self.num_inference_steps = n
        self.pipe.scheduler.set_timesteps(n)

    def set_dimensions(self, size_output):
        s = self.pipe.vae_scale_factor
        if size_output is None:
            w = h = self.pipe.unet_cfg.sample_size
        else:
            w, h = size_output
        self.width_img = int(round(w / s) * s)
        self.width_latent = int(self.width_img / s)
        self.height_img = int(round(h / s) * s)
        self.height_latent = int(self.height_img / s)

    def set_negative_prompt(self, negative_prompt):
        self.negative_prompt = [negative_prompt] if isinstance(negative_prompt, str) else negative_prompt
        if len(self.negative_prompt) > 1:
            self.negative_prompt = [self.negative_prompt[0]]

    # -- inputs -----------------------------------------------------------
    def do_cfg(self):
        return self.guidance_scale > 1 and self.pipe.unet_cfg.time_cond_proj_dim is None

    def get_text_embedding(self, prompt):
        return self.pipe.encode_prompt(prompt, self.negative_prompt, self.do_cfg(), self.dtype)

    def get_noise(self, seed=420):
        return self.pipe.prepare_latents(self.height_latent, self.width_latent, seed, self.dtype)

    def latent2image(self, latents, output_type="np"):
        return latent2image_np(self.pipe.vae, latents)

    # -- the loop ---------------------------------------------------------
    def prepare_mixing(self, mixing_coeffs, list_latents_mixing):
        N = self.num_inference_steps
        if type(mixing_coeffs) == float:
            coeffs = (1 + N) * [mixing_coeffs]
        elif type(mixing_coeffs) == list:
            assert len(mixing_coeffs) == N, f"len(mixing_coeffs) {len(mixing_coeffs)} != {N}"
            coeffs = mixing_coeffs
        else:
            raise ValueError("mixing_coeffs should be float or list with len=num_inference_steps")
        if np.sum(coeffs) > 0:
            assert len(list_latents_mixing) == N, f"len(list_latents_mixing) {len(list_latents_mixing)} != {N}"
        return coeffs

    def unet_eps(self, model_input, t, ctx, text_embeds, time_ids):
        out = self.pipe.unet(model_input.float(), t, ctx.float(), text_embeds.float(), time_ids.float())
        return out.to(model_input.dtype)

    @torch.no_grad()
    def run_diffusion_sd_xl(self, text_embeddings, latents_start, idx_start=0,
                            list_latents_mixing=None, mixing_coeffs=0.0, return_image=False):
        sched = self.pipe.scheduler
        coeffs = self.prepare_mixing(mixing_coeffs, list_latents_mixing)
        pe, ne, pp, npool = text_embeddings
        sched.set_timesteps(self.num_inference_steps)                 # retrieve_timesteps, :247
        latents = latents_start.clone()
        cfg_on = self.guidance_scale > 1                              # pipe.do_classifier_free_guidance, :282
        tid = self.pipe.add_time_ids(pe.dtype)
        if cfg_on:
            ctx = torch.cat([ne, pe], dim=0)
            text = torch.cat([npool, pp], dim=0)
            tids = torch.cat([tid, tid], dim=0)
        else:
            ctx, text, tids = pe, pp, tid
        out = []
        for i, t in enumerate(sched.timesteps):
            if i < idx_start:
                out.append(None)
                continue
            elif i == idx_start:
                latents = latents_start.clone()
            if i > 0 and coeffs[i] > 0:                               # :322-324
                latents = interpolate_spherical(latents, list_latents_mixing[i - 1].clone(), coeffs[i])
            x = torch.cat([latents] * 2) if cfg_on else latents
            x = sched.scale_model_input(x, i)
            eps = self.unet_eps(x, float(t), ctx, text, tids)
            if cfg_on:
                e_u, e_t = eps.chunk(2)
                eps = e_u + self.guidance_scale * (e_t - e_u)         # :347-349
            noise = None
            if sched.ancestral and self.noise_fn is not None:
                noise = self.noise_fn(i, eps.shape).to(eps.dtype)
            latents = sched.step(eps, i, latents, noise=noise)
            out.append(latents.clone())
        if return_image:
            return self.latent2image(latents)
        return out
