"""Where does a transition spend its time outside the UNet?  (CUDA-event timings of the pieces)"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from latentblending_b200 import BlendingEngine, SyntheticSDXLPipe, ops

def ev(fn, n=5):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n

pipe = SyntheticSDXLPipe("stabilityai/stable-diffusion-xl-base-1.0", "cuda:0")
be = BlendingEngine(pipe)
be.set_prompt1("a"); be.set_prompt2("b"); be.set_branching(depth_strength=0.5, nmb_max_branches=15)
lat = be.dh.get_noise(1) * 0.1
print("vae decode ms", ev(lambda: be.dh.decode_to_device(lat)))
f = be.dh.decode_to_device(lat); g = be.dh.decode_to_device(lat * 0.5)
print("lpips pair ms", ev(lambda: be.lpips.distance(f, g)))
emb = be.get_mixed_conditioning(0.3)[0]
plan = be.dh.unet.plan(2, 128, 128)
print("prog_ctx (70 kv gemms) ms", ev(lambda: plan.prog_ctx.run()))
print("unet step ms", ev(lambda: plan.prog_step.run(500.0)))
print("mixed conditioning ms", ev(lambda: be.get_mixed_conditioning(0.3)))
t1 = be.compute_latents1(); t2 = be.compute_latents2()
print("parental mix ms", ev(lambda: be._parental_mix(t1, t2, 0.4)))
torch.cuda.synchronize(); t0 = time.time()
be.output_device_frames = True
be.run_transition(fixed_seeds=[420, 421]); torch.cuda.synchronize()
print("transition s", time.time() - t0, "unet calls", be.dh.n_unet_calls)
# host-side cost of the python loop around one denoise step
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
be.run_transition(fixed_seeds=[420, 421]); torch.cuda.synchronize()
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
