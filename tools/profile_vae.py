"""Run one SDXL-width VAE decode (128x128 latents -> 1024^2 frame) for ncu / CUDA-event timing.
The profiled region is bracketed with cudaProfilerStart/Stop: use `ncu --profile-from-start off ...`."""
import os
import sys

os.environ.setdefault("LB_NO_GRAPH", "1")     # one ncu record per kernel; time the direct launches

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from latentblending_b200.pipe import VAE_CHANNELS, random_state_dict, vae_param_shapes  # noqa: E402
from latentblending_b200.vae import VAEDecoderB200  # noqa: E402


def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    dev = "cuda:0"
    vae = VAEDecoderB200(random_state_dict(vae_param_shapes(VAE_CHANNELS), 1, dev, damp=0.3), VAE_CHANNELS, 0.13025, dev)
    lat = (torch.randn(1, 4, L, L, device=dev) * 0.8).half()
    for _ in range(3):
        vae.decode_to_u8(lat)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        vae.decode_to_u8(lat)
    e1.record()
    torch.cuda.synchronize()
    print(f'{{"vae_decode_ms": {e0.elapsed_time(e1) / 5:.3f}, "latent": {L}, "launches": {vae.plan(L, L).prog.num_launches}}}', flush=True)
    torch.cuda.profiler.start()
    vae.decode_to_u8(lat)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
