"""Run one SDXL UNet forward (CFG batch 2, 128x128 latents) for ncu.  The profiled region is
bracketed with cudaProfilerStart/Stop: use `ncu --profile-from-start off ...`."""
import os
import sys

os.environ.setdefault("LB_NO_GRAPH", "1")     # profile the direct launches (one ncu record per kernel either way)

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from latentblending_b200.pipe import SDXL_BASE, random_state_dict, unet_param_shapes  # noqa: E402
from latentblending_b200.unet import UNetB200  # noqa: E402


def main():
    B, h, w = 2, 128, 128
    if len(sys.argv) > 1:
        h = w = int(sys.argv[1])
    dev = "cuda:0"
    net = UNetB200(SDXL_BASE, random_state_dict(unet_param_shapes(SDXL_BASE), 0, dev), dev)
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(B, 4, h, w, generator=g, device=dev).half()
    ctx = (torch.randn(B, 77, 2048, generator=g, device=dev) * 0.5).half()
    pooled = torch.randn(B, 1280, generator=g, device=dev).half()
    tids = torch.tensor([[8. * h, 8. * w, 0, 0, 8. * h, 8. * w]] * B, device=dev).half()
    for _ in range(2):
        net.forward(x, 500.0, ctx, pooled, tids)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    net.forward(x, 500.0, ctx, pooled, tids, ctx_changed=False)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("done", net.launches_per_forward(B, h, w))


if __name__ == "__main__":
    main()
