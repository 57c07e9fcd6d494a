"""Experiment: the two CFG halves of a batch-2 UNet forward as two INDEPENDENT batch-1 programs on two CUDA streams
(space-sharing the SMs, one kernel's ramp / drain overlapping the other stream's main loop) vs the single batch-2
program.  Kernels are batch-invariant, so both give identical eps."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from latentblending_b200.pipe import SDXL_BASE, random_state_dict, unet_param_shapes  # noqa: E402
from latentblending_b200.unet import UNetB200, _Lowering  # noqa: E402

dev = "cuda:0"
net = UNetB200(SDXL_BASE, random_state_dict(unet_param_shapes(SDXL_BASE), 0, dev), dev)
L = int(sys.argv[1]) if len(sys.argv) > 1 else 128


def timed(fn, reps=8, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(2, 4, L, L, generator=g, device=dev).half()
ctx = (torch.randn(2 * 77, 2048, generator=g, device=dev) * 0.5).half()
text = torch.randn(2, 1280, generator=g, device=dev).half()
tids = torch.tensor([[8. * L, 8. * L, 0, 0, 8. * L, 8. * L]] * 2, device=dev).half()

p2 = net.plan(2, L, L)
p2.x_in.copy_(x); p2.ctx.copy_(ctx); p2.text.copy_(text); p2.tids.copy_(tids)
p2.prog_ctx.run()
ms2 = timed(lambda: p2.prog_step.run(500.0))
eps2 = p2.eps.clone()
print(f'{{"variant": "one batch-2 program, one stream", "unet_step_ms": {ms2:.3f}}}', flush=True)

halves = [_Lowering(net, 1, L, L) for _ in range(2)]
for b, pl in enumerate(halves):
    pl.x_in.copy_(x[b:b + 1]); pl.ctx.copy_(ctx[b * 77:(b + 1) * 77]); pl.text.copy_(text[b:b + 1]); pl.tids.copy_(tids[b:b + 1])
    pl.prog_ctx.run()
ms1 = timed(lambda: halves[0].prog_step.run(500.0))
print(f'{{"variant": "one batch-1 program alone", "unet_step_ms": {ms1:.3f}}}', flush=True)
streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
torch.cuda.synchronize()


def dual():
    cur = torch.cuda.current_stream()
    ev = torch.cuda.Event()
    ev.record(cur)
    done = []
    for pl, st in zip(halves, streams):
        st.wait_event(ev)
        with torch.cuda.stream(st):
            pl.prog_step.run(500.0)
            d = torch.cuda.Event()
            d.record(st)
            done.append(d)
    for d in done:
        cur.wait_event(d)


msd = timed(dual)
same = all(torch.equal(halves[b].eps[0], eps2[b]) for b in range(2))
print(f'{{"variant": "two batch-1 programs on two streams", "unet_step_ms": {msd:.3f}, "bit_identical_to_batch2": {str(same).lower()}}}',
      flush=True)
