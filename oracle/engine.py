"""Oracle: branch-tree host logic (test infrastructure).

Restates latentblending/blending_engine.py host logic, quirks included:
  :120-207  setters / model-keyed defaults (base model ignores user parental
            crossfeed args, :200-203)
  :155-164  set_guidance_mid_dampening
  :258-293  set_branching           :467-529  get_time_based_branching
  :295-365  run_transition          :370-465  compute_latents1/2/_mix
  :531-562  get_mixing_parameters   :564-588  insert_into_tree
  :643-654  get_mixed_conditioning  :731-742  swap_forward
  :767-789  get_closest_idx
``tree_similarities`` starts as a one-element list holding a non-number
(:349 stores the bound method); ``np.argmax`` of a 1-list is 0, which is all the
first insertion needs -- reproduced here with a ``None`` placeholder semantics
(``_argmax_gap``).
"""
import numpy as np
import torch

from .lpips_alex import LPIPSAlex, lpips_distance
from .mixing import interpolate_linear, interpolate_spherical


def time_based_branching(num_inference_steps, depth_strength, dt_unet_step, dt_vae,
                         t_compute_max_allowed=None, nmb_max_branches=None):
    """blending_engine.py:467-529."""
    N = num_inference_steps
    idx_base = int(np.floor(N * depth_strength))
    steps = int(np.ceil(N / 10))
    list_idx = np.arange(idx_base, N, steps)
    stems = np.ones(len(list_idx), dtype=np.int32)
    if nmb_max_branches is None:
        assert t_compute_max_allowed is not None
        criterion = "time"
    elif t_compute_max_allowed is None:
        criterion = "count"
        nmb_max_branches -= 2                                   # :498 outer frames discounted
    else:
        raise ValueError("Either specify t_compute_max_allowed or nmb_max_branches")
    done, first = False, True
    while not done:
        compute_steps = (N - list_idx) * stems
        t_compute = np.sum(compute_steps) * dt_unet_step + dt_vae * np.sum(stems)
        t_compute += 2 * (N * dt_unet_step + dt_vae)
        grown = False
        for s in range(len(stems) - 1):
            if stems[s + 1] / stems[s] >= 1:
                stems[s] += 1
                grown = True
                break
        if not grown:
            stems[-1] += 1
        if criterion == "time" and t_compute > t_compute_max_allowed:
            done = True
        elif criterion == "count" and np.sum(stems) >= nmb_max_branches:
            done = True
            if first:                                           # :521-524 undersample
                list_idx = np.linspace(list_idx[0], list_idx[-1], nmb_max_branches).astype(np.int32)
                stems = np.ones(len(list_idx), dtype=np.int32)
        else:
            first = False
    return list_idx, stems


def crossfeed_coeffs_branch1(N, power, rng, decay):
    """blending_engine.py:406-408."""
    stop = int(round(N * rng))
    c = list(np.linspace(power, power * decay, stop))
    c.extend((N - stop) * [0])
    return c


def crossfeed_coeffs_parental(N, idx_injection, power, rng, decay):
    """blending_engine.py:452-457."""
    stop = int(round(N * rng))
    c = idx_injection * [power]
    n_mix = stop - idx_injection
    if n_mix > 0:
        c.extend(list(np.linspace(power, power * decay, n_mix)))
    c.extend((N - len(c)) * [0])
    return c


def guidance_mid_dampening(guidance_scale_base, damper, fract):
    """blending_engine.py:155-164."""
    mid = 1 - np.abs(fract - 0.5) / 0.5
    max_red = guidance_scale_base * (1 - damper) - 1
    return guidance_scale_base - max_red * mid


def closest_idx(tree_fracts, fract):
    """blending_engine.py:767-789."""
    d = fract - np.asarray(tree_fracts)
    pos = d.copy()
    pos[pos < 0] = np.inf
    b1 = int(np.argmin(pos))
    neg = -d.copy()
    neg[neg <= 0] = np.inf
    b2 = int(np.argmin(neg))
    return (b2, b1) if b1 > b2 else (b1, b2)


class OracleEngine:
    def __init__(self, holder, guidance_scale_mid_damper=0.5, lpips_net=None):
        assert 0 < guidance_scale_mid_damper <= 1.0
        self.dh = holder
        self.damper = guidance_scale_mid_damper
        self.seed1 = self.seed2 = 0
        self.tree_latents = [None, None]
        self.tree_fracts = None
        self.tree_final_imgs = []
        self.negative_prompt = None
        self.set_dimensions()
        self.set_guidance_scale()
        self.lpips = lpips_net or LPIPSAlex()
        self.set_prompt1("")
        self.set_prompt2("")
        self.set_branch1_crossfeed()
        self.set_parental_crossfeed()
        self.set_num_inference_steps()
        self.dt_unet_step, self.dt_vae = 0.05, 0.1      # benchmark_speed() stand-ins, overridable

    # setters ---------------------------------------------------------------
    def set_dimensions(self, size_output=None):
        if size_output is None:                         # blending_engine.py:128-132
            size_output = (512, 512) if self.dh.is_sdxl_turbo else (1024, 1024)
        self.dh.set_dimensions(size_output)

    def set_guidance_scale(self, g=None):
        if g is None:
            g = 0.0 if self.dh.is_sdxl_turbo else 4.0
        self.guidance_scale_base = self.guidance_scale = self.dh.guidance_scale = g

    def set_negative_prompt(self, neg):
        self.negative_prompt = neg
        self.dh.set_negative_prompt(neg)

    def set_branch1_crossfeed(self, power=0, rng=0, decay=0):
        self.b1_power, self.b1_range, self.b1_decay = (np.clip(v, 0, 1) for v in (power, rng, decay))

    def set_parental_crossfeed(self, power=None, rng=None, decay=None):
        if self.dh.is_sdxl_turbo:
            power = 1.0 if power is None else power
            rng = 1.0 if rng is None else rng
            decay = 1.0 if decay is None else decay
        else:
            power, rng, decay = 0.3, 0.6, 0.9
        self.p_power, self.p_range, self.p_decay = (np.clip(v, 0, 1) for v in (power, rng, decay))

    def set_prompt1(self, p):
        self.prompt1 = p.replace("_", " ")
        self.text_embedding1 = self.dh.get_text_embedding(self.prompt1)

    def set_prompt2(self, p):
        self.prompt2 = p.replace("_", " ")
        self.text_embedding2 = self.dh.get_text_embedding(self.prompt2)

    def set_num_inference_steps(self, n=None):
        if n is None:
            n = 4 if self.dh.is_sdxl_turbo else 30
        self.num_inference_steps = n
        self.dh.set_num_inference_steps(n)

    def set_branching(self, depth_strength=None, t_compute_max_allowed=None, nmb_max_branches=None):
        N = self.num_inference_steps
        if self.dh.is_sdxl_turbo:
            assert t_compute_max_allowed is None
            idx = int(round(N * depth_strength)) if depth_strength is not None else 2
            self.list_idx_injection = [idx]
            self.list_nmb_stems = [10 if nmb_max_branches is None else nmb_max_branches]
        else:
            if depth_strength is None:
                depth_strength = 0.5
            if t_compute_max_allowed is None and nmb_max_branches is None:
                t_compute_max_allowed = 20
            self.list_idx_injection, self.list_nmb_stems = time_based_branching(
                N, depth_strength, self.dt_unet_step, self.dt_vae, t_compute_max_allowed, nmb_max_branches)

    # tree ------------------------------------------------------------------
    def mixed_conditioning(self, fract):
        return [None if a is None else interpolate_linear(a, b, fract)
                for a, b in zip(self.text_embedding1, self.text_embedding2)]

    def similarity(self, img_a, img_b):
        return lpips_distance(self.lpips, img_a, img_b)

    def compute_latents1(self):
        cond = self.mixed_conditioning(0)
        traj = self.dh.run_diffusion_sd_xl(cond, self.dh.get_noise(self.seed1), idx_start=0)
        self.tree_latents[0] = traj
        return traj

    def compute_latents2(self):
        cond = self.mixed_conditioning(1)
        start = self.dh.get_noise(self.seed2)
        if self.b1_power > 0.0:
            coeffs = crossfeed_coeffs_branch1(self.num_inference_steps, self.b1_power, self.b1_range, self.b1_decay)
            traj = self.dh.run_diffusion_sd_xl(cond, start, 0, self.tree_latents[0], coeffs)
        else:
            traj = self.dh.run_diffusion_sd_xl(cond, start)
        self.tree_latents[-1] = traj
        return traj

    def compute_latents_mix(self, fract, b1, b2, idx_injection):
        cond = self.mixed_conditioning(fract)
        f_par = (fract - self.tree_fracts[b1]) / (self.tree_fracts[b2] - self.tree_fracts[b1])
        mix = []
        for i in range(self.num_inference_steps):
            a, b = self.tree_latents[b1][i], self.tree_latents[b2][i]
            mix.append(None if (a is None or b is None) else interpolate_spherical(a, b, f_par))
        coeffs = crossfeed_coeffs_parental(self.num_inference_steps, idx_injection,
                                           self.p_power, self.p_range, self.p_decay)
        return self.dh.run_diffusion_sd_xl(cond, mix[idx_injection - 1], idx_injection, mix, coeffs)

    def _argmax_gap(self):
        sims = self.tree_similarities
        if len(sims) == 1:
            return 0                      # :349 quirk -- argmax of a 1-list
        return int(np.argmax(sims))

    def get_mixing_parameters(self, idx_injection):
        c1 = self._argmax_gap()
        c2 = c1 + 1
        fract = (self.tree_fracts[c1] + self.tree_fracts[c2]) / 2
        p1 = c1
        while self.tree_idx_injection[p1] >= idx_injection:
            p1 -= 1
        p2 = c2
        while self.tree_idx_injection[p2] >= idx_injection:
            p2 += 1
        return fract, p1, p2

    def insert_into_tree(self, fract, idx_injection, traj):
        img = self.dh.latent2image(traj[-1])
        b1, b2 = closest_idx(self.tree_fracts, fract)
        left = self.similarity(img, self.tree_final_imgs[b1])
        right = self.similarity(img, self.tree_final_imgs[b2])
        k = b1 + 1
        self.tree_latents.insert(k, traj)
        self.tree_final_imgs.insert(k, img)
        self.tree_fracts.insert(k, fract)
        self.tree_idx_injection.insert(k, idx_injection)
        self.tree_similarities[b1] = left
        self.tree_similarities.insert(k, right)

    def run_transition(self, recycle_img1=False, recycle_img2=False, fixed_seeds=None):
        N = self.num_inference_steps
        if fixed_seeds is not None:
            if isinstance(fixed_seeds, str) and fixed_seeds == "randomize":
                fixed_seeds = list(np.random.randint(0, 1000000, 2).astype(np.int32))
            else:
                assert len(fixed_seeds) == 2
            self.seed1, self.seed2 = fixed_seeds
        t1 = self.tree_latents[0] if (recycle_img1 and self.tree_latents[0] is not None
                                      and len(self.tree_latents[0]) == N) else self.compute_latents1()
        t2 = self.tree_latents[-1] if (recycle_img2 and self.tree_latents[-1] is not None
                                       and len(self.tree_latents[-1]) == N) else self.compute_latents2()
        self.tree_latents = [t1, t2]
        self.tree_fracts = [0.0, 1.0]
        self.tree_final_imgs = [self.dh.latent2image(t1[-1]), self.dh.latent2image(t2[-1])]
        self.tree_idx_injection = [0, 0]
        self.tree_similarities = [None]
        for idx_injection, n_stems in zip(self.list_idx_injection, self.list_nmb_stems):
            for _ in range(int(n_stems)):
                fract, p1, p2 = self.get_mixing_parameters(int(idx_injection))
                g = guidance_mid_dampening(self.guidance_scale_base, self.damper, fract)
                self.guidance_scale = self.dh.guidance_scale = g
                traj = self.compute_latents_mix(fract, p1, p2, int(idx_injection))
                self.insert_into_tree(fract, int(idx_injection), traj)
        return self.tree_final_imgs

    def swap_forward(self):
        self.tree_latents[0] = self.tree_latents[-1]
        self.prompt1, self.text_embedding1 = self.prompt2, self.text_embedding2
        self.tree_final_imgs = []
