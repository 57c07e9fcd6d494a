#!/usr/bin/env python
"""bench.py -- transition frames/sec of the branch-tree denoising hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one ``BlendingEngine.run_transition()`` of BASELINE.json configs[1]:
SDXL-base-shaped UNet (2.57 B random-init parameters), 1024x1024 (128x128 latents),
30 Euler steps, depth_strength 0.5, nmb_max_branches 15 -> 15 frames, 198 CFG-batch-2
UNet forwards, 13 parental mixes, 15 VAE decodes, 26 LPIPS evaluations.  Synthetic
data (no network): seeded random weights / embeddings, fixed seeds [420, 421].

Prints ONE JSON line (rank 0).  ``value`` = frames/s with the conditioning already
on the device and frames left on the device; ``e2e`` = the same through the public
API (set_prompt1/2 -> run_transition -> PIL frames), host<->device copies timed.
``--impl reference`` times the CPU oracle (a port of the reference path: diffusers /
lpips are not installable here) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "transition frames/sec (1024^2 SDXL, 30 steps, 15 branches)"
WORKLOAD = dict(workload="single_trans SDXL 1024x1024, 30 steps, depth_strength=0.5, nmb_max_branches=15",
                frames=15, unet_forwards=198, cfg_batch=2, latent="128x128", weights="random-init SDXL-base shape",
                seeds=[420, 421], l2="working set (5.1 GB fp16 weights per UNet forward) exceeds L2; no flush needed")
PROMPTS = ("photo of underwater landscape, fish, und the sea, incredible detail, high resolution",
           "rendering of an alien planet, strange plants, strange creatures, surreal")
NEG = "blurry, ugly, pale"


def peaks():
    p = dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")
    fp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(fp):
        with open(fp) as f:
            p.update(json.load(f))
        p["source"] = "measured"
    return p


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        self.th.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ---------------------------------------------------------------------------------------------
# Analytic FLOPs of one SDXL UNet sample: everything but self-attention scales with the
# number of latent pixels, self-attention QK^T / PV with its square.
_UNET_TFLOP_1024 = dict(linear_conv=5.977, self_attn=0.752, cross_attn=0.032)


def unet_tflop(latent):
    r = (latent / 128.0) ** 2
    u = _UNET_TFLOP_1024
    return (u["linear_conv"] + u["cross_attn"]) * r + u["self_attn"] * r * r


CPU_SAMPLE_LATENT = 64            # the bounded CPU sample runs the oracle at 512x512 px (64x64 latents)


def cpu_port_sample(threads=None):
    """One bounded sample of the CPU oracle: a full fp32 CFG-batch-2 SDXL UNet forward, a VAE decode and an LPIPS pair
    at 64x64 latents (512 px) plus a full-size 30-row parental mix; UNet / VAE / LPIPS times are scaled to the bench
    shape (128x128 latents) by the analytic FLOP ratio and the transition time is extrapolated with the exact call
    counts of the workload.  A full-size CPU transition would take ~12 h."""
    import torch
    from oracle import mixing
    from oracle.lpips_alex import LPIPSAlex, lpips_distance
    from oracle.sdxl_unet import SDXL_BASE, SDXLUNet
    from oracle.vae import SDXL_VAE, VAEDecoder, latent2image_np
    if threads:
        torch.set_num_threads(threads)
    cores = torch.get_num_threads()
    L = CPU_SAMPLE_LATENT
    state = cpu_port_sample.__dict__.setdefault("state", {})
    if "unet" not in state:
        with torch.no_grad():
            state["unet"] = SDXLUNet(SDXL_BASE).eval()
            state["vae"] = VAEDecoder(SDXL_VAE).eval()
            state["lpips"] = LPIPSAlex()
        g = torch.Generator().manual_seed(0)
        state["x"] = torch.randn(2, 4, L, L, generator=g)
        state["ctx"] = torch.randn(2, 77, 2048, generator=g) * 0.5
        state["pool"] = torch.randn(2, 1280, generator=g)
        state["tid"] = torch.tensor([[8. * L, 8. * L, 0, 0, 8. * L, 8. * L]] * 2)
        state["lat"] = torch.randn(1, 4, L, L, generator=g).half()
        state["traj"] = [torch.randn(1, 4, 128, 128, generator=g).half() for _ in range(60)]
    s = state
    out = {}
    with torch.no_grad():
        t0 = time.time()
        s["unet"](s["x"], 500.0, s["ctx"], s["pool"], s["tid"])
        out["t_unet_fwd_b2_sample"] = time.time() - t0
        if "t_vae" not in s:
            t0 = time.time()
            img = latent2image_np(s["vae"], s["lat"])
            s["t_vae"] = time.time() - t0
            t0 = time.time()
            mixing.parental_mix(s["traj"][:30], s["traj"][30:], 0.4)
            s["t_mix"] = time.time() - t0
            t0 = time.time()
            lpips_distance(s["lpips"], img, img[::-1].copy())
            s["t_lpips"] = time.time() - t0
    px = (128.0 / L) ** 2
    out["t_unet_fwd_b2"] = out["t_unet_fwd_b2_sample"] * unet_tflop(128) / unet_tflop(L)
    out.update(t_vae=s["t_vae"] * px, t_mix=s["t_mix"], t_lpips=s["t_lpips"] * px,
               t_vae_sample=s["t_vae"], t_lpips_sample=s["t_lpips"])
    total = 198 * out["t_unet_fwd_b2"] + 15 * out["t_vae"] + 13 * out["t_mix"] + 26 * out["t_lpips"]
    out["transition_s_extrapolated"] = total
    out["frames_per_s"] = 15.0 / total
    out["cores"] = cores
    return out


CPU_SAMPLE_TEXT = ("bounded sample: 1 fp32 CFG-batch-2 SDXL UNet forward (2.57 B params), 1 VAE decode and 1 LPIPS pair of the "
                   "CPU oracle at 64x64 latents (512 px) + one full-size 30-row parental mix; UNet time scaled by the "
                   "analytic FLOP ratio 128^2 vs 64^2 latents (x%.2f), VAE/LPIPS by the pixel ratio (x4); transition "
                   "time extrapolated with the exact call counts (198 UNet, 15 VAE, 13 mixes, 26 LPIPS)"
                   % (unet_tflop(128) / unet_tflop(CPU_SAMPLE_LATENT)))


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    import torch
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    for _ in range(args.warmup):
        cpu_port_sample(threads)
    vals = [cpu_port_sample(threads) for _ in range(max(1, args.steps))]
    fps = sum(v["frames_per_s"] for v in vals) / len(vals)
    tt = sum(v["transition_s_extrapolated"] for v in vals) / len(vals)
    sample = "per step: " + CPU_SAMPLE_TEXT
    line = dict(metric=METRIC, value=fps, unit="frames/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=tt * 1e3, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f32",
                data="synthetic", config=dict(WORKLOAD), impl="reference",
                cpu_baseline=dict(value=fps, unit="frames/s", cores=vals[-1]["cores"], kind="port", sample=sample,
                                  detail={k: round(v, 4) for k, v in vals[-1].items() if k.startswith("t_")}),
                e2e=dict(value=fps, unit="frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    rank, local_rank, world = dist_env()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device for --impl ours"
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(dev))
    from latentblending_b200 import BlendingEngine, SyntheticSDXLPipe, ops
    from latentblending_b200._cabi import OP_ATTENTION, OP_GEMM, OP_GROUPNORM, OP_LAYERNORM
    pk = peaks()
    pipe = SyntheticSDXLPipe("stabilityai/stable-diffusion-xl-base-1.0", dev, seed=0)
    be = BlendingEngine(pipe)
    be.set_negative_prompt(NEG)
    be.set_prompt1(PROMPTS[0])
    be.set_prompt2(PROMPTS[1])
    be.set_branching(depth_strength=0.5, nmb_max_branches=15)
    assert [int(v) for v in be.list_nmb_stems] == [4, 3, 3, 2, 1], be.list_nmb_stems

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 0
        for _ in range(steps):
            n += fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms / 1e3, n

    def step_device():
        be.output_device_frames = True
        return len(be.run_transition(fixed_seeds=[420, 421]))

    def step_api():
        be.output_device_frames = False
        be.set_prompt1(PROMPTS[0])
        be.set_prompt2(PROMPTS[1])
        return len(be.run_transition(fixed_seeds=[420, 421]))

    for _ in range(max(args.warmup, 0)):
        step_device()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ops.LAUNCHES[0] = 0
    sec, frames = timed(step_device, args.steps)
    launches = ops.LAUNCHES[0]
    clocks = sampler.stop()
    fps = frames / sec
    # e2e through the public API with host buffers
    step_api()
    pipe.h2d_bytes = 0
    be.d2h_bytes = 0
    sec_e, frames_e = timed(step_api, args.steps)
    e2e = dict(value=frames_e / sec_e, unit="frames/s", h2d_bytes_per_step=pipe.h2d_bytes // max(1, args.steps),
               d2h_bytes_per_step=be.d2h_bytes // max(1, args.steps))

    # roofline of the dominant kernel (gemm_tc_kernel): all GEMM launches of one UNet forward, replayed
    # back to back on the launching stream between CUDA events
    plan = be.dh.unet.plan(2, 128, 128)
    work = plan.prog_step.work()
    breakdown = {}
    for name, kinds in (("gemm", [OP_GEMM]), ("attention", [OP_ATTENTION]), ("norms", [OP_GROUPNORM, OP_LAYERNORM]),
                        ("all", list(range(1, 11)))):
        for _ in range(2):
            plan.prog_step.run_kinds(kinds, 500.0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            n_l = plan.prog_step.run_kinds(kinds, 500.0)
        e1.record()
        torch.cuda.synchronize()
        breakdown[name] = dict(ms=e0.elapsed_time(e1) / reps, launches=n_l)
    # K1 (crossfeed / parental mix) in its batched form: 2048 rows x 65536 fp16 (805 MB through the kernel, > L2)
    mp0 = torch.randn(2048, 4 * 128 * 128, device=dev).half()
    mp1 = torch.randn(2048, 4 * 128 * 128, device=dev).half()
    mout = torch.empty_like(mp0)
    for _ in range(3):
        ops.slerp_rows(mp0, mp1, 0.4, out=mout)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.slerp_rows(mp0, mp1, 0.4, out=mout)
    e1.record()
    torch.cuda.synchronize()
    mix_s = e0.elapsed_time(e1) / 10 * 1e-3
    mix_gbs = mp0.numel() * 6 / mix_s / 1e9
    del mp0, mp1, mout
    gemm_tf = work["gemm_flops"] / (breakdown["gemm"]["ms"] * 1e-3) / 1e12
    attn_tf = work["attn_flops"] / (breakdown["attention"]["ms"] * 1e-3) / 1e12
    peak_tf = pk["bf16_tflops_sustained"]
    roofline = dict(kernel="gemm_tc_kernel (tcgen05 GEMM / implicit-GEMM conv)", bound="tensor", achieved=gemm_tf,
                    peak=peak_tf, unit="TFLOP/s", frac=gemm_tf / peak_tf, traffic=None,
                    peak_source=f"{pk['source']} bf16_tflops_sustained (kernel timed inside a long step)",
                    algorithmic_flops_per_unet_forward=work["gemm_flops"],
                    avg_launch_us=breakdown["gemm"]["ms"] * 1e3 / breakdown["gemm"]["launches"],
                    launches_per_unet_forward=breakdown["gemm"]["launches"],
                    unet_forward_breakdown_ms={k: round(v["ms"], 3) for k, v in breakdown.items()},
                    attention=dict(kernel="attn_tc_kernel (tcgen05 QK^T / PV, head dim 64)", achieved=attn_tf,
                                   frac=attn_tf / peak_tf, flops=work["attn_flops"],
                                   note="all 140 attention launches of one UNet forward incl. 70 cross-attention "
                                        "(77 keys); at d=64 MUFU.EX2 alone caps the tensor pipe at 50 %"),
                    mix=dict(kernel="slerp_l2_kernel (K1 parental / crossfeed mix, batched 2048 x 65536 fp16)",
                             bound="hbm", achieved=mix_gbs, peak=pk["hbm_gbs"], unit="GB/s", frac=mix_gbs / pk["hbm_gbs"],
                             algorithmic_bytes_per_element=6, launch_us=mix_s * 1e6,
                             traffic=790.4e6, traffic_source="ncu dram__bytes_read+write per launch, "
                             "profiles/r01d_ncu_full_summary.txt (algorithmic: 805.3e6)"))

    line = dict(metric=METRIC, value=fps, unit="frames/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=sec / max(1, args.steps) * 1e3, higher_is_better=True, scaling="strong",
                vs_baseline=None, dtype="f16", data="synthetic",
                config=dict(WORKLOAD, parallelism=f"branch-sharded x{world}" if world > 1 else "single GPU"),
                clocks=clocks, e2e=e2e, gpu_launches=launches, roofline=roofline, impl="ours")
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            c = cpu_port_sample(os.cpu_count())
            line["cpu_baseline"] = dict(
                value=c["frames_per_s"], unit="frames/s", cores=c["cores"], kind="port",
                sample=CPU_SAMPLE_TEXT,
                detail={k: round(v, 4) for k, v in c.items() if k.startswith("t_")})
        except Exception as ex:   # the baseline is a reported number, never a reason to lose the bench line
            line["cpu_baseline"] = dict(value=None, unit="frames/s", cores=os.cpu_count(), kind="port",
                                        sample=f"failed: {ex!r}")
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
