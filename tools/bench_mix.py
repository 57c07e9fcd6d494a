"""Micro-benchmark of K1 (batched parental mix) and K9 (CFG+Euler step):
achieved algorithmic GB/s vs the measured HBM peak.  CUDA-event timed, inputs
larger than L2 for the headline row count."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from latentblending_b200 import ops  # noqa: E402


def time_it(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e-3


def main():
    peak = 6573.5
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk))["hbm_gbs"]
    n = 4 * 128 * 128
    res = []
    for rows in (1, 30, 210, 840, 2048):
        p0 = torch.randn(rows, n, device="cuda").half()
        p1 = torch.randn(rows, n, device="cuda").half()
        out = torch.empty_like(p0)
        t = time_it(lambda: ops.slerp_rows(p0, p1, 0.4, out=out))
        gbs = rows * n * 6 / t / 1e9
        res.append(dict(kernel="slerp_rows", rows=rows, n=n, us=t * 1e6, GBs=gbs, frac=gbs / peak))
    for n_ in (4 * 64 * 64, 4 * 128 * 128):
        p0 = torch.randn(840, n_, device="cuda").half()
        p1 = torch.randn(840, n_, device="cuda").half()
        out = torch.empty_like(p0)
        t = time_it(lambda: ops.slerp_rows(p0, p1, 0.4, out=out))
        res.append(dict(kernel="slerp_rows", rows=840, n=n_, us=t * 1e6, GBs=840 * n_ * 6 / t / 1e9))
    x = torch.randn(1, 4, 128, 128, device="cuda").half()
    eps = torch.randn(2, 4, 128, 128, device="cuda").half()
    o = torch.empty_like(x)
    t = time_it(lambda: ops.cfg_euler_step(x, eps, 4.0, 1.5, -0.2, out=o, traj=o))
    res.append(dict(kernel="cfg_euler_step", n=x.numel(), us=t * 1e6))
    for r in res:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
