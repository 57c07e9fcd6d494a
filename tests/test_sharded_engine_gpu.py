"""Multi-GPU tests (skipped unless enough CUDA devices): the product engine's branch-sharded transitions over NCCL
-- single GPUs per candidate on 2 ranks (a CFG pair for the last stem of a level), CFG pairs on 4 ranks -- build
exactly the tree of the single-GPU sequential engine (tiny SDXL-shaped pipeline, real kernels), also over a CHAIN of
two transitions (swap_forward + recycle_img1), after which every rank must hold the sequential path's guidance
state, and with branch-1 crossfeed (the outer pair in lockstep, CFG halves split over the two ranks)."""
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


def _build(dev, crossfeed=False):
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
    if crossfeed:
        be.set_branch1_crossfeed(0.8, 0.6, 0.4)
    be.output_device_frames = True
    return be


def _chain(be):
    """Two chained transitions like example_multi_trans.py:38-62; returns the summaries of both."""
    be.run_transition(fixed_seeds=[420, 421])
    first = _summary(be)
    be.swap_forward()
    be.set_prompt2("three")
    be.run_transition(recycle_img1=True, fixed_seeds=[421, 422])
    return first, _summary(be), float(be.guidance_scale), float(be.dh.guidance_scale)


def _summary(be):
    return (list(be.tree_fracts), [int(v) for v in be.tree_idx_injection],
            [float(t[-1].float().sum()) for t in be.tree_latents], [float(s) for s in be.tree_similarities])


def _worker(rank, world, port, q, crossfeed):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    be = _build(f"cuda:{rank}", crossfeed)
    res = _chain(be) + (dict(be.shard_stats),)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        be2 = _build("cuda:0", crossfeed)
        q.put(("seq",) + _chain(be2))                        # no process group any more: sequential path
    q.put((rank,) + res)


def _run(world, crossfeed):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, crossfeed)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=900) for _ in range(world + 1)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    seq = next(g for g in got if g[0] == "seq")
    stats = None
    for g in got:
        if g[0] == "seq":
            continue
        for t in (1, 2):                                     # first and second (chained) transition
            assert g[t][0] == seq[t][0] and g[t][1] == seq[t][1], (g[0], t, g[t][0], seq[t][0])
            # same kernels, same inputs, same order -> identical latents and similarities on every rank
            assert g[t][2] == seq[t][2], (g[0], t)
            assert g[t][3] == seq[t][3], (g[0], t)
        # the guidance state every rank is left with equals the sequential engine's (it steers the next transition)
        assert g[3] == seq[3] and g[4] == seq[4], (g[0], g[3], seq[3])
        stats = g[5]
    print(f"world {world} crossfeed {crossfeed}: shard stats {stats}")
    return stats


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("crossfeed", [False, True])
def test_sharded_engine_matches_sequential_engine(crossfeed):
    stats = _run(2, crossfeed)
    assert stats["paired_rounds"] >= 1          # 2 ranks: the last stem of each level runs as a CFG pair


@pytest.mark.skipif(torch.cuda.device_count() < 4, reason="needs 4 GPUs")
def test_sharded_engine_cfg_pairs_on_four_gpus():
    stats = _run(4, False)
    assert 1 <= stats["paired_rounds"] <= stats["rounds"]
