"""Oracle: latent mixing arithmetic (test infrastructure, see oracle/__init__.py).

Follows latentblending/utils.py:29-71 (interpolate_spherical) and :74-102
(interpolate_linear) of the reference.
"""
import numpy as np
import torch

SLERP_CLAMP_EPS = 1e-7  # utils.py:55


def slerp_scalars(p0, p1, fract):
    """The two fp64 scalars (s0, s1) of the whole-tensor slerp, utils.py:52-63.

    Returned as python floats so tests can compare the device-side reduction
    separately from the axpby.
    """
    a = p0.detach().to(torch.float64).reshape(-1)
    b = p1.detach().to(torch.float64).reshape(-1)
    norm = torch.linalg.norm(a) * torch.linalg.norm(b)                # utils.py:54
    dot = torch.sum(a * b) / norm                                     # utils.py:56
    dot = dot.clamp(-1 + SLERP_CLAMP_EPS, 1 - SLERP_CLAMP_EPS)        # utils.py:57
    theta0 = torch.arccos(dot)                                        # utils.py:59
    sin0 = torch.sin(theta0)
    theta_t = theta0 * fract
    s0 = torch.sin(theta0 - theta_t) / sin0                           # utils.py:62
    s1 = torch.sin(theta_t) / sin0                                    # utils.py:63
    return float(s0), float(s1)


def interpolate_spherical(p0, p1, fract_mixing):
    """Whole-tensor slerp in fp64, result recast to fp16 if p0 is fp16 else fp32
    (utils.py:47-71)."""
    out_dtype = torch.float16 if p0.dtype == torch.float16 else torch.float32
    s0, s1 = slerp_scalars(p0, p1, fract_mixing)
    interp = p0.to(torch.float64) * s0 + p1.to(torch.float64) * s1    # utils.py:64
    return interp.to(out_dtype)


def interpolate_linear(p0, p1, fract_mixing):
    """(1-f)*p0 + f*p1; uint8 numpy inputs go through fp64 and are clipped
    (utils.py:88-102)."""
    back_to_u8 = False
    if isinstance(p0, np.ndarray) and p0.dtype == np.uint8:
        back_to_u8, p0 = True, p0.astype(np.float64)
    if isinstance(p1, np.ndarray) and p1.dtype == np.uint8:
        back_to_u8, p1 = True, p1.astype(np.float64)
    out = (1 - fract_mixing) * p0 + fract_mixing * p1
    if back_to_u8:
        out = np.clip(out, 0, 255).astype(np.uint8)
    return out


def parental_mix(traj1, traj2, fract):
    """blending_engine.py:442-450: per-step slerp of two parent trajectories,
    None wherever either parent has no latent for that step."""
    out = []
    for a, b in zip(traj1, traj2):
        out.append(None if (a is None or b is None) else interpolate_spherical(a, b, fract))
    return out
