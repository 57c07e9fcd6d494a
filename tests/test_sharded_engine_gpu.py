"""2-GPU test (skipped unless >= 2 CUDA devices): the product engine's branch-sharded transition over NCCL
builds exactly the tree of the single-GPU sequential engine (tiny SDXL-shaped pipeline, real kernels)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build(dev):
    from latentblending_b200 import BlendingEngine, SyntheticSDXLPipe
    from latentblending_b200.unet import UNetConfig
    cfg = UNetConfig(block_out_channels=(64, 128, 256), transformer_layers=(0, 1, 2), cross_attention_dim=128,
                     addition_time_embed_dim=32, pooled_dim=64, sample_size=16)
    pipe = SyntheticSDXLPipe("synthetic/sdxl-base-tiny", dev, unet_cfg=cfg, seed=5, vae_channels=(64, 64, 128, 128))
    be = BlendingEngine(pipe, run_benchmark=False)
    be.set_dimensions((128, 128))
    be.set_num_inference_steps(10)
    be.set_prompt1("one")
    be.set_prompt2("two")
    be.set_branching(depth_strength=0.5, nmb_max_branches=9)
    be.output_device_frames = True
    return be


def _summary(be):
    return (list(be.tree_fracts), [int(v) for v in be.tree_idx_injection],
            [float(t[-1].float().sum()) for t in be.tree_latents], [float(s) for s in be.tree_similarities])


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    be = _build(f"cuda:{rank}")
    be.run_transition(fixed_seeds=[420, 421])
    res = _summary(be) + (dict(be.shard_stats),)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        be2 = _build("cuda:0")
        be2.run_transition(fixed_seeds=[420, 421])          # no process group any more: sequential path
        q.put(("seq",) + _summary(be2))
    q.put((rank,) + res)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_engine_matches_sequential_engine():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=600) for _ in range(3)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    seq = next(g for g in got if g[0] == "seq")
    for g in got:
        if g[0] == "seq":
            continue
        assert g[1] == seq[1] and g[2] == seq[2], (g[0], g[1], seq[1])
        # same kernels, same inputs, same order -> identical latents and similarities on every rank
        assert g[3] == seq[3]
        np.testing.assert_allclose(g[4], seq[4], rtol=1e-5)
    print("shard stats", got[0][-1] if got[0][0] != "seq" else got[1][-1])
