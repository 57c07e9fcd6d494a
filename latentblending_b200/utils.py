"""Math utilities with the reference's names (latentblending/utils.py), backed
by the CUDA kernels for device tensors.

interpolate_spherical  <- utils.py:29-71   (K1 kernel, lb_slerp_rows)
interpolate_linear     <- utils.py:74-102  (lb_lerp for CUDA tensors; numpy for
                                            uint8 frames, which never touch the GPU)
"""
import time

import numpy as np
import torch
import yaml

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


def add_frames_linear_interp(list_imgs, fps_target=None, duration_target=None, nmb_frames_target=None, seed=None):
    """Fill a frame list up to an exact frame count with linear blends
    (utils.py:105-178).  The per-gap insert counts are drawn at random until they
    sum to the target; ``seed`` makes that reproducible."""
    if nmb_frames_target is not None and fps_target is not None:
        raise ValueError("You cannot specify both fps_target and nmb_frames_target")
    if fps_target is None:
        assert nmb_frames_target is not None, "Either specify nmb_frames_target or fps_target+duration_target"
    if nmb_frames_target is None:
        assert fps_target is not None and duration_target is not None
        nmb_frames_target = fps_target * duration_target
    gaps = len(list_imgs) - 1
    missing = nmb_frames_target - gaps - 1
    if missing < 1:
        return list_imgs
    frames = [np.asarray(im).astype(np.float32) for im in list_imgs]
    mean_ins = missing / gaps
    base = np.floor(mean_ins)
    thresh = 1 - (mean_ins - base)
    rng = np.random if seed is None else np.random.RandomState(seed)
    for _ in range(100001):
        ins = (rng.rand(gaps) > thresh).astype(np.float64) + base
        if np.sum(ins) == missing:
            break
    ins = ins.astype(np.int32)
    out = []
    for i in range(gaps):
        out.append(frames[i].astype(np.uint8))
        for f in np.linspace(0, 1, ins[i] + 2)[1:-1]:
            out.append(interpolate_linear(frames[i], frames[i + 1], f).astype(np.uint8))
    out.append(frames[-1].astype(np.uint8))
    return out


def get_spacing(nmb_points: int, scaling: float):
    """Non-linear spacing on [0,1], denser around 0.5 (utils.py:181-200)."""
    if scaling < 1.7:
        return np.linspace(0, 1, nmb_points)
    per_side = nmb_points // 2 + 1
    left = np.abs(np.linspace(1, 0, per_side) ** scaling / 2 - 0.5)
    if nmb_points % 2 != 0:
        right = 1 - left[::-1][1:]
    else:
        left = left[:-1]
        right = 1 - left[::-1]
    return np.hstack([left, right])


def get_time(resolution=None):
    """Time string like 221117_1620 (utils.py:203-221)."""
    fmt = {None: "%y%m%d_%H%M%S", "second": "%y%m%d_%H%M%S", "minute": "%y%m%d_%H%M", "day": "%y%m%d"}
    if resolution == "millisecond":
        return time.strftime("%y%m%d_%H%M%S", time.localtime()) + "_{:03d}".format(int((time.time() % 1) * 1000))
    if resolution not in fmt:
        raise ValueError("bad resolution provided: %s" % resolution)
    return time.strftime(fmt[resolution], time.localtime())


def yml_load(fp_yml, print_fields=False):
    with open(fp_yml) as f:
        return dict(yaml.load(f, Loader=yaml.loader.SafeLoader))


def yml_save(fp_yml, dict_stuff):
    with open(fp_yml, "w") as f:
        yaml.dump(dict_stuff, f, sort_keys=False, default_flow_style=False)
