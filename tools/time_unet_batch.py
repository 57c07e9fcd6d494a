"""UNet step time vs batch (CFG pairs): is batching two branches into one B=4 forward worth it?"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from latentblending_b200.pipe import SDXL_BASE, random_state_dict, unet_param_shapes  # noqa: E402
from latentblending_b200.unet import UNetB200  # noqa: E402

dev = "cuda:0"
net = UNetB200(SDXL_BASE, random_state_dict(unet_param_shapes(SDXL_BASE), 0, dev), dev)
import json
BATCHES = [int(a) for a in sys.argv[1:]] or [2, 4, 6, 8]
for B in BATCHES:
    plan = net.plan(B, 128, 128)
    plan.prog_ctx.run()
    for _ in range(3):
        plan.prog_step.run(500.0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(8):
        plan.prog_step.run(500.0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 8
    from latentblending_b200 import _cabi
    print(json.dumps(dict(B=B, unet_step_ms=round(ms, 3), ms_per_cfg_pair=round(ms / max(1, B // 2), 3),
                          graph=int(_cabi.load().lb_program_is_graph(plan.prog_step.handle)),
                          env={k: v for k, v in os.environ.items() if k.startswith("LB_")})), flush=True)
    del plan
    net._plans.clear()
    torch.cuda.empty_cache()
