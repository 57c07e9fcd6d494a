"""DiffusersHolder: the denoise operator behind BlendingEngine, on liblb200.

Same attributes and methods the reference engine touches on ``self.dh``
(latentblending/diffusers_holder.py:20-366; the 14-item seam in SURVEY.md
section 8b): device, pipe, get_text_embedding, get_noise, run_diffusion_sd_xl,
latent2image, is_sdxl_turbo, set_dimensions, guidance_scale, set_negative_prompt,
set_num_inference_steps, height_img / width_img.

B200-first differences (results are the same, layout and launch structure are not):
  * a trajectory is ONE contiguous [N,4,h,w] fp16 slab in HBM; the returned
    ``list_latents_out`` holds views into it (None for i < idx_start), so the
    parental mix of a whole branch is a single batched lb_slerp_rows launch;
  * per step: (optional crossfeed slerp) -> lb_scale_model_input straight into the
    UNet program's input buffer -> lb_program_run (~930 pre-planned launches) ->
    lb_cfg_euler_step (CFG + Euler + trajectory store) -- no torch arithmetic;
  * the cross-attention K/V projections depend only on the conditioning and are
    computed once per branch, not once per step.
There is no CPU path: everything raises without CUDA + liblb200.so.
"""
import numpy as np
import torch

from . import ops
from .vae import VAEDecoderB200
from .unet import UNetB200


class DiffusersHolder:
    def __init__(self, pipe):
        self.negative_prompt = ""                 # reference defaults, diffusers_holder.py:23-25
        self.guidance_scale = 5.0
        self.num_inference_steps = 30
        self.pipe = pipe
        self.device = str(pipe._execution_device)
        if not torch.cuda.is_available() or not self.device.startswith("cuda"):
            raise RuntimeError("latentblending_b200 needs a CUDA device (no CPU fallback)")
        self.dtype = torch.float16
        self.is_sdxl_turbo = "turbo" in pipe._name_or_path
        self.pipe.scheduler.set_timesteps(self.num_inference_steps, device=self.device)
        s = pipe.unet_cfg.sample_size
        self.width_latent = self.height_latent = s
        self.width_img = self.height_img = s * pipe.vae_scale_factor
        self.unet = UNetB200(pipe.unet_cfg, pipe.unet_state_dict, self.device)
        self.vae = VAEDecoderB200(pipe.vae_state_dict, pipe.vae_channels, pipe.vae_scaling_factor, self.device)
        self.noise_fn = None          # tests: inject the ancestral-step noise, noise_fn(i, shape)
        self._cond_key = None
        self.n_unet_calls = 0

    # ---- configuration --------------------------------------------------------------------
    def set_num_inference_steps(self, num_inference_steps):
        self.num_inference_steps = num_inference_steps
        self.pipe.scheduler.set_timesteps(num_inference_steps, device=self.device)

    def set_dimensions(self, size_output):
        s = self.pipe.vae_scale_factor
        if size_output is None:
            width = height = self.pipe.unet_cfg.sample_size
        else:
            width, height = size_output
        self.width_img = int(round(width / s) * s)
        self.width_latent = int(self.width_img / s)
        self.height_img = int(round(height / s) * s)
        self.height_latent = int(self.height_img / s)

    def set_negative_prompt(self, negative_prompt):
        self.negative_prompt = [negative_prompt] if isinstance(negative_prompt, str) else negative_prompt
        if len(self.negative_prompt) > 1:
            self.negative_prompt = [self.negative_prompt[0]]

    # ---- inputs -----------------------------------------------------------------------------
    def get_text_embedding(self, prompt):
        do_cfg = self.guidance_scale > 1 and self.pipe.unet_cfg.time_cond_proj_dim is None
        return self.pipe.encode_prompt(prompt, negative_prompt=self.negative_prompt,
                                       do_classifier_free_guidance=do_cfg)

    def get_noise(self, seed=420):
        """randn([1,4,h,w], fp16, CUDA generator(seed)) * init_noise_sigma -- as pipe.prepare_latents does."""
        g = torch.Generator(device=self.device).manual_seed(int(seed))
        lat = torch.randn((1, self.pipe.unet_cfg.in_channels, self.height_latent, self.width_latent), generator=g,
                          device=self.device, dtype=torch.float16)
        return lat * self.pipe.scheduler.init_noise_sigma.to(device=self.device, dtype=torch.float16)

    @torch.no_grad()
    def decode_to_device(self, latents):
        """latents [1,4,h,w] -> uint8 [H,W,3] frame on the device."""
        return self.vae.decode_to_u8(latents.to(torch.float16))

    @torch.no_grad()
    def latent2image(self, latents, output_type="pil"):
        assert output_type in ["pil", "np"]
        arr = self.decode_to_device(latents).cpu().numpy()
        if output_type == "np":
            return arr.astype(np.float32) / 255.0
        from PIL import Image
        return Image.fromarray(arr)

    # ---- the loop -----------------------------------------------------------------------------
    def prepare_mixing(self, mixing_coeffs, list_latents_mixing):
        N = self.num_inference_steps
        if type(mixing_coeffs) == float:
            list_mixing_coeffs = (1 + N) * [mixing_coeffs]
        elif type(mixing_coeffs) == list:
            assert len(mixing_coeffs) == N, f"len(mixing_coeffs) {len(mixing_coeffs)} != self.num_inference_steps {N}"
            list_mixing_coeffs = mixing_coeffs
        else:
            raise ValueError("mixing_coeffs should be float or list with len=num_inference_steps")
        if np.sum(list_mixing_coeffs) > 0:
            assert len(list_latents_mixing) == N, \
                f"len(list_latents_mixing) {len(list_latents_mixing)} != self.num_inference_steps {N}"
        return list_mixing_coeffs

    def run_diffusion(self, text_embeddings, latents_start, idx_start=0, list_latents_mixing=None, mixing_coeffs=0.0,
                      return_image=False):
        return self.run_diffusion_sd_xl(text_embeddings, latents_start, idx_start, list_latents_mixing, mixing_coeffs,
                                        return_image)

    @torch.no_grad()
    def run_diffusion_sd_xl(self, text_embeddings, latents_start, idx_start=0, list_latents_mixing=None,
                            mixing_coeffs=0.0, return_image=False):
        sched = self.pipe.scheduler
        N = self.num_inference_steps
        coeffs = self.prepare_mixing(mixing_coeffs, list_latents_mixing)
        pe, ne, pp, npool = text_embeddings
        sched.set_timesteps(N, device=self.device)
        cfg_on = self.guidance_scale > 1                       # pipe.do_classifier_free_guidance
        hw = self.pipe.default_sample_size * self.pipe.vae_scale_factor   # original/target size, :216-220
        tid = torch.tensor([[hw, hw, 0, 0, hw, hw]], dtype=torch.float16, device=self.device)
        if cfg_on:
            ctx = torch.cat([ne, pe], dim=0)
            text = torch.cat([npool, pp], dim=0)
            tids = torch.cat([tid, tid], dim=0)
        else:
            ctx, text, tids = pe, pp, tid
        B = 2 if cfg_on else 1
        _, C, h, w = latents_start.shape
        plan = self.unet.plan(B, h, w)
        plan.ctx.copy_(ctx.reshape(plan.ctx.shape))
        plan.text.copy_(text)
        plan.tids.copy_(tids)
        plan.prog_ctx.run()                                    # cross-attention K/V: once per conditioning
        traj = torch.empty((N, C, h, w), dtype=torch.float16, device=self.device)
        latents = latents_start.clone().contiguous()
        out = [None] * N
        n = latents.numel()
        for i in range(N):
            if i < idx_start:
                continue
            elif i == idx_start:
                latents = latents_start.clone().contiguous()
            if i > 0 and coeffs[i] > 0:
                latents = ops.slerp_rows(latents.view(1, n), list_latents_mixing[i - 1].reshape(1, n),
                                         float(coeffs[i])).view(1, C, h, w)
            sc = sched.step_scalars[i]
            ops.scale_model_input(latents, B, sc["divisor"], out=plan.x_in)
            plan.prog_step.run(sc["t"])
            self.n_unet_calls += 1
            noise = None
            if sched.ancestral:
                if self.noise_fn is not None:
                    noise = self.noise_fn(i, latents.shape).to(device=self.device, dtype=torch.float16).contiguous()
                else:
                    noise = torch.randn(latents.shape, device=self.device, dtype=torch.float16)
            new = traj[i:i + 1]
            ops.cfg_euler_step(latents, plan.eps, self.guidance_scale, sc["sigma"], sc["dt"], sc["sigma_up"],
                               noise=noise, out=new)
            latents = new
            out[i] = new
        if return_image:
            return self.latent2image(latents)
        return out
