"""Tensor-level wrappers over the C ABI (device memory and streams come from
PyTorch; all arithmetic happens in liblb200.so).  Every function requires CUDA
tensors and raises otherwise -- there is no CPU path."""
import numpy as np
import torch

from . import _cabi
from ._cabi import check, ctx, ptr, stream_ptr

_DT = {torch.float16: 0, torch.float32: 1}


def _dev(t):
    if not t.is_cuda:
        raise _cabi.LB200Error("latentblending_b200 ops need CUDA tensors (no CPU fallback)")
    return t.device.index or 0


_ws_cache = {}


def _workspace(dev, nbytes):
    buf = _ws_cache.get(dev)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8, device=f"cuda:{dev}")
        _ws_cache[dev] = buf
    return buf


def slerp_rows(p0, p1, fract, out=None, fract_rows=None):
    """rows x n whole-row slerp: p0, p1 are [rows, n] (row stride arbitrary, inner
    contiguous).  utils.py:29-71 per row."""
    assert p0.dim() == 2 and p0.shape == p1.shape and p0.dtype == p1.dtype
    assert p0.stride(1) == 1 and p1.stride(1) == 1
    dev = _dev(p0)
    rows, n = p0.shape
    if out is None:
        out = torch.empty((rows, n), dtype=p0.dtype, device=p0.device)
    assert out.shape == p0.shape and out.stride(1) == 1 and out.dtype == p0.dtype
    lib = _cabi.load()
    ws = _workspace(dev, lib.lb_slerp_workspace_bytes(rows, n))
    check(lib.lb_slerp_rows(ctx(dev), ptr(p0), ptr(p1), ptr(out), rows, n,
                            p0.stride(0) if rows > 1 else n, p1.stride(0) if rows > 1 else n,
                            out.stride(0) if rows > 1 else n, _DT[p0.dtype], float(fract),
                            ptr(fract_rows), ptr(ws), stream_ptr()), "lb_slerp_rows")
    return out


def lerp(p0, p1, fract):
    assert p0.shape == p1.shape and p0.dtype == p1.dtype and p0.is_contiguous() and p1.is_contiguous()
    dev = _dev(p0)
    out = torch.empty_like(p0)
    check(_cabi.load().lb_lerp(ctx(dev), ptr(p0), ptr(p1), ptr(out), p0.numel(), _DT[p0.dtype], float(fract),
                               stream_ptr()), "lb_lerp")
    return out


def scale_model_input(latents, batch, divisor, out=None):
    assert latents.dtype == torch.float16 and latents.is_contiguous()
    dev = _dev(latents)
    n = latents.numel()
    if out is None:
        out = torch.empty((batch,) + tuple(latents.shape[1:]), dtype=torch.float16, device=latents.device)
    check(_cabi.load().lb_scale_model_input(ctx(dev), ptr(latents), ptr(out), n, int(batch), float(divisor),
                                            stream_ptr()), "lb_scale_model_input")
    return out


def cfg_euler_step(latents, eps, guidance, sigma, dt, sigma_up=0.0, noise=None, out=None, traj=None):
    """eps: [2,...] (uncond, text) when CFG is on else [1,...]."""
    assert latents.dtype == torch.float16 and eps.dtype == torch.float16
    assert latents.is_contiguous() and eps.is_contiguous()
    dev = _dev(latents)
    n = latents.numel()
    use_cfg = eps.numel() == 2 * n
    assert use_cfg or eps.numel() == n
    if out is None:
        out = torch.empty_like(latents)
    check(_cabi.load().lb_cfg_euler_step(ctx(dev), ptr(latents), ptr(eps), ptr(noise), ptr(out), ptr(traj), n,
                                         int(use_cfg), float(np.float32(guidance)), float(sigma), float(dt),
                                         float(sigma_up), stream_ptr()), "lb_cfg_euler_step")
    return out


def _p(t):
    return None if t is None else t.data_ptr()


def gemm(a0, w, N, B, H, W, taps=1, a0_c=None, a1=None, a1_c=None, bias=None, bias2=None, res=None, out=None,
         mode=0, out_cols=None):
    """Tensor-core GEMM / implicit-GEMM conv (lb_gemm).  a0: NHWC activation viewed as
    [B*H*W, >=a0_c] (row stride = a0.stride(0)); w: [N, K] packed weights."""
    dev = _dev(a0)
    M = B * H * W
    a0_c = a0.shape[-1] if a0_c is None else a0_c
    n_out = (N // 2 if mode == 1 else N) if out_cols is None else out_cols
    if out is None:
        out = torch.empty((M, n_out), dtype=torch.float16, device=a0.device)
    d = _cabi.GemmDesc()
    d.a0, d.a0_ld, d.a0_c = _p(a0), a0.stride(-2), a0_c
    if a1 is not None:
        d.a1, d.a1_ld, d.a1_c = _p(a1), a1.stride(-2), (a1.shape[-1] if a1_c is None else a1_c)
    d.B, d.H, d.W, d.taps = B, H, W, taps
    d.w, d.w_ld, d.N = _p(w), w.stride(0), N
    d.bias = _p(bias)
    if bias2 is not None:
        d.bias2, d.bias2_ld = _p(bias2), bias2.stride(0)
    if res is not None:
        d.res, d.res_ld = _p(res), res.stride(-2)
    d.out, d.out_ld, d.mode = _p(out), out.stride(-2), mode
    check(_cabi.load().lb_gemm(ctx(dev), d, stream_ptr()), "lb_gemm")
    return out


def error_flag(dev=0):
    import ctypes
    code = ctypes.c_int(0)
    check(_cabi.load().lb_ctx_error_flag(ctx(dev), ctypes.byref(code)), "lb_ctx_error_flag")
    return code.value
