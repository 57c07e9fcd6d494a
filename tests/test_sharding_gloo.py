"""world_size-2 (and 3) gloo tests of the branch-level sharding logic on CPU.

The arithmetic is injected (closed-form fake "denoise"), exactly like tests/golden/fakes.py, so
no CUDA kernel is needed: what is checked is that every rank ends with the SAME tree as a
sequential single-process run of the reference's greedy loop (same tree_fracts order, same
parents, same similarities, same latents)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from latentblending_b200.sharding import LevelSharder, older_parents, plan_candidates, run_level_local


class Tree:
    def __init__(self, N, seed):
        g = torch.Generator().manual_seed(seed)
        a = [torch.randn(1, 4, 8, 8, generator=g).half() for _ in range(N)]
        b = [torch.randn(1, 4, 8, 8, generator=g).half() for _ in range(N)]
        self.tree_latents = [a, b]
        self.tree_fracts = [0.0, 1.0]
        self.tree_idx_injection = [0, 0]
        self.tree_similarities = [None]
        self.frames = [frame_of(a[-1]), frame_of(b[-1])]


def frame_of(lat):
    return ((torch.tanh(lat.float()[0, :3]) * 0.5 + 0.5) * 255).round().to(torch.uint8)


def sim(a, b):
    return float((a.float() - b.float()).abs().mean() / 255.0)


def make_compute(tree, N, log=None):
    def compute(fract, p1, p2, idx, cfg_split=None):
        """cfg_split mimics the CFG pair: each member computes HALF of the per-step update (its "eps half") and the
        halves are exchanged inside the 2-rank team group every step, like DiffusersHolder.cfg_split."""
        f = (fract - tree.tree_fracts[p1]) / (tree.tree_fracts[p2] - tree.tree_fracts[p1])
        traj = [None] * N
        lat = ((1 - f) * tree.tree_latents[p1][idx - 1].float() + f * tree.tree_latents[p2][idx - 1].float()).half()
        if log is not None:
            log.append((fract, None if cfg_split is None else cfg_split["half"]))
        for i in range(idx, N):
            upd = 0.1 * torch.sin(3 * lat.float() + i + 10 * fract)
            if cfg_split is not None:
                half = upd.clone()
                flat = half.view(-1)
                n2 = flat.numel() // 2
                mine = flat[:n2].clone() if cfg_split["half"] == 0 else flat[n2:].clone()
                parts = [torch.empty_like(mine), torch.empty_like(mine)]
                dist.all_gather(parts, mine, group=cfg_split["group"])
                upd = torch.cat(parts).view(upd.shape)
            lat = (lat.float() * 0.9 + upd).half()
            traj[i] = lat.clone()
        return traj, frame_of(traj[-1])
    return compute


def sequential(N, levels, seed):
    """The reference's loop (blending_engine.py:354-362, :531-588) in one process."""
    tree = Tree(N, seed)
    tree.insert_order = []
    compute = make_compute(tree, N)
    for idx, stems in levels:
        for _ in range(stems):
            s = tree.tree_similarities
            c1 = 0 if len(s) == 1 else int(np.argmax(s))
            mid = (tree.tree_fracts[c1] + tree.tree_fracts[c1 + 1]) / 2
            p1, p2 = c1, c1 + 1
            while tree.tree_idx_injection[p1] >= idx:
                p1 -= 1
            while tree.tree_idx_injection[p2] >= idx:
                p2 += 1
            assert (p1, p2) == older_parents(tree.tree_fracts, tree.tree_idx_injection, mid, idx)
            traj, frm = compute(mid, p1, p2, idx)
            tree.insert_order.append(mid)
            left, right = sim(frm, tree.frames[c1]), sim(frm, tree.frames[c1 + 1])
            k = c1 + 1
            tree.tree_latents.insert(k, traj)
            tree.frames.insert(k, frm)
            tree.tree_fracts.insert(k, mid)
            tree.tree_idx_injection.insert(k, idx)
            tree.tree_similarities[c1] = left
            tree.tree_similarities.insert(k, right)
    return tree


def _worker(rank, world, port, N, levels, seed, q, pairs=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tree = Tree(N, seed)
    sh = LevelSharder(rank, world, cfg_pairs=pairs)
    log, inserted = [], []
    compute = make_compute(tree, N, log)
    for idx, stems in levels:
        # every rank evaluates the similarities itself (deterministic) -- no broadcast from rank 0
        sh.run_level(tree, idx, stems, compute, sim, N, on_insert=inserted.append)
    q.put((rank, tree.tree_fracts, tree.tree_idx_injection, tree.tree_similarities,
           [float(t[-1].float().sum()) for t in tree.tree_latents], sh.stats, inserted, log))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,levels,pairs", [(2, [(15, 4), (18, 3), (21, 3), (24, 2), (27, 1)], False),
                                                (3, [(2, 7)], False), (2, [(10, 1), (20, 6)], False),
                                                (2, [(15, 4), (18, 3), (27, 1)], True),
                                                (4, [(15, 4), (18, 3), (21, 3), (24, 2), (27, 1)], True)])
def test_sharded_tree_equals_sequential(world, levels, pairs):
    N, seed = 30, 7
    ref = sequential(N, levels, seed)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, levels, seed, q, pairs)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, fracts, idxs, sims, sums, stats, inserted, log in results:
        assert fracts == ref.tree_fracts, rank
        assert idxs == ref.tree_idx_injection
        np.testing.assert_allclose(sims, ref.tree_similarities, rtol=0, atol=0)
        assert sums == [float(t[-1].float().sum()) for t in ref.tree_latents]
        assert stats["used"] == sum(s for _, s in levels)
        # on_insert ran on EVERY rank once per inserted branch, in the sequential insertion order
        assert inserted == ref.insert_order, rank
        if pairs:
            assert stats["paired_rounds"] >= 1
            assert any(h == rank % 2 for _, h in log if h is not None) or world == 3
        else:
            assert all(h is None for _, h in log)
    # speculation can only save rounds, never add any
    assert results[0][5]["rounds"] <= sum(s for _, s in levels)
    if pairs and world == 4:
        assert 1 <= results[0][5]["paired_rounds"] <= results[0][5]["rounds"]   # pairs for the last stem(s) of a level
    print("rounds", results[0][5]["rounds"], "of", sum(s for _, s in levels), "branches; computed",
          results[0][5]["computed"])


def test_plan_candidates_first_is_reference_choice():
    fr = [0.0, 0.25, 0.5, 1.0]
    sims = [0.1, 0.4, 0.3]
    c = plan_candidates(fr, sims, 4, {})
    assert c[0] == (0.375, 0.25, 0.5)
    assert len(c) == 4 and len({x[0] for x in c}) == 4
    # cached midpoints are skipped but their halves are still explored
    c2 = plan_candidates(fr, sims, 2, {0.375: None})
    assert all(x[0] != 0.375 for x in c2)
    # the very first insertion: a 1-list holding a non-number (blending_engine.py:349)
    assert plan_candidates([0.0, 1.0], [None], 3, {})[0] == (0.5, 0.0, 1.0)


@pytest.mark.parametrize("width", [2, 3, 4])
@pytest.mark.parametrize("seed", [7, 8, 9])
def test_local_lockstep_speculation_equals_sequential(width, seed):
    """run_level_local (the single-GPU form of the same plan / replay: candidates share one batched forward) builds the
    reference's sequential tree whatever the width and however many candidates miss; on_insert sees the insertions in
    the sequential order; the adaptive split ratio is carried from level to level."""
    N, levels = 30, [(15, 4), (18, 3), (21, 3), (24, 2), (27, 1)]
    ref = sequential(N, levels, seed)
    tree = Tree(N, seed)
    comp = make_compute(tree, N)
    batches = []

    def many(cands, idx):
        batches.append(len(cands))
        return [comp(m, p1, p2, idx) for m, p1, p2 in cands]
    stats, inserted, ratio = dict(rounds=0, computed=0, used=0), [], 0.6
    for idx, stems in levels:
        ratio = run_level_local(tree, idx, stems, many, sim, width, on_insert=inserted.append, stats=stats,
                                split_ratio=ratio)
    assert tree.tree_fracts == ref.tree_fracts and tree.tree_idx_injection == ref.tree_idx_injection
    np.testing.assert_allclose(tree.tree_similarities, ref.tree_similarities, rtol=0, atol=0)
    assert [float(t[-1].float().sum()) for t in tree.tree_latents] == [float(t[-1].float().sum()) for t in ref.tree_latents]
    assert inserted == ref.insert_order
    assert stats["used"] == 13 and stats["computed"] >= 13 and stats["rounds"] == len(batches)
    assert max(batches) <= width and 0.0 < ratio <= 1.0
