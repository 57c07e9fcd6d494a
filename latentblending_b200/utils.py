"""Math utilities with the reference's names (latentblending/utils.py), backed
by the CUDA kernels for device tensors.

interpolate_spherical  <- utils.py:29-71   (K1 kernel, lb_slerp_rows)
interpolate_linear     <- utils.py:74-102  (lb_lerp for CUDA tensors; numpy for
                                            uint8 frames, which never touch the GPU)
"""
import numpy as np
import torch

from . import ops


@torch.no_grad()
def interpolate_spherical(p0, p1, fract_mixing: float):
    """Whole-tensor slerp (fp64 arithmetic on device); returns fp16 if p0 is fp16 else fp32."""
    if not (torch.is_tensor(p0) and p0.is_cuda):
        raise RuntimeError("latentblending_b200.interpolate_spherical needs CUDA tensors (no CPU fallback)")
    dt = torch.float16 if p0.dtype == torch.float16 else torch.float32
    a = p0.to(dt).contiguous().view(1, -1)
    b = p1.to(dt).contiguous().view(1, -1)
    return ops.slerp_rows(a, b, float(fract_mixing)).view(p0.shape)


def interpolate_linear(p0, p1, fract_mixing):
    """(1-f)*p0 + f*p1.  CUDA tensors go through lb_lerp; uint8 numpy frames are
    mixed in fp64 and clipped like the reference."""
    if torch.is_tensor(p0) and p0.is_cuda and p0.dtype in (torch.float16, torch.float32):
        return ops.lerp(p0.contiguous(), p1.contiguous(), float(fract_mixing))
    back = False
    if isinstance(p0, np.ndarray) and p0.dtype == np.uint8:
        back, p0 = True, p0.astype(np.float64)
    if isinstance(p1, np.ndarray) and p1.dtype == np.uint8:
        back, p1 = True, p1.astype(np.float64)
    out = (1 - fract_mixing) * p0 + fract_mixing * p1
    return np.clip(out, 0, 255).astype(np.uint8) if back else out


def _insert_counts(gaps, nmb_frames_target, seed=None):
    """Frames to insert into each gap so the total is exactly ``nmb_frames_target`` (utils.py:143-160: floor of the
    mean plus a random 0/1 per gap, redrawn until the sum fits).  None when nothing is missing."""
    missing = nmb_frames_target - gaps - 1
    if missing < 1:
        return None
    mean_ins = missing / gaps
    base = np.floor(mean_ins)
    thresh = 1 - (mean_ins - base)
    rng = np.random if seed is None else np.random.RandomState(seed)
    for _ in range(100001):
        ins = (rng.rand(gaps) > thresh).astype(np.float64) + base
        if np.sum(ins) == missing:
            break
    return ins.astype(np.int32)


def plan_frame_fill(n_key_frames, nmb_frames_target, seed=None):
    """The frame fill of add_frames_linear_interp as index/weight arrays for lb_frames_lerp_u8: output frame t is
    uint8(float32(w0[t]) * key[left[t]] + float32(w1[t]) * key[left[t]+1]) -- numpy's float32 blend
    ``(1 - f) * img0 + f * img1`` (utils.py:97,170) of the reference's float32 images."""
    gaps = n_key_frames - 1
    ins = _insert_counts(gaps, nmb_frames_target, seed)
    if ins is None:
        ins = np.zeros(gaps, dtype=np.int32)
    left, w0, w1 = [], [], []
    for i in range(gaps):
        left.append(i); w0.append(1.0); w1.append(0.0)
        for f in np.linspace(0, 1, ins[i] + 2)[1:-1]:
            left.append(i); w0.append(1 - f); w1.append(f)
    left.append(gaps - 1 if gaps > 0 else 0); w0.append(0.0 if gaps > 0 else 1.0); w1.append(1.0 if gaps > 0 else 0.0)
    return (np.asarray(left, dtype=np.int32), np.asarray(w0, dtype=np.float64).astype(np.float32),
            np.asarray(w1, dtype=np.float64).astype(np.float32))


def add_frames_linear_interp(list_imgs, fps_target=None, duration_target=None, nmb_frames_target=None, seed=None):
    """Fill a frame list up to an exact frame count with linear blends
    (utils.py:105-178).  The per-gap insert counts are drawn at random until they
    sum to the target; ``seed`` makes that reproducible.  Host (numpy) version for images that are not on the
    device; BlendingEngine.get_movie_frames runs the same plan through lb_frames_lerp_u8."""
    if nmb_frames_target is not None and fps_target is not None:
        raise ValueError("You cannot specify both fps_target and nmb_frames_target")
    if fps_target is None:
        assert nmb_frames_target is not None, "Either specify nmb_frames_target or fps_target+duration_target"
    if nmb_frames_target is None:
        assert fps_target is not None and duration_target is not None
        nmb_frames_target = fps_target * duration_target
    gaps = len(list_imgs) - 1
    ins = _insert_counts(gaps, nmb_frames_target, seed)
    if ins is None:
        return list_imgs
    frames = [np.asarray(im).astype(np.float32) for im in list_imgs]
    out = []
    for i in range(gaps):
        out.append(frames[i].astype(np.uint8))
        for f in np.linspace(0, 1, ins[i] + 2)[1:-1]:
            # float32 arithmetic like the reference's numpy (value-based casting of the float64 scalar)
            blend = np.float32(1 - f) * frames[i] + np.float32(f) * frames[i + 1]
            out.append(blend.astype(np.uint8))
    out.append(frames[-1].astype(np.uint8))
    return out
