"""CPU tests of the product BlendingEngine's host logic against the reference-generated
golden file (tests/golden/tree.json): branching planner, gap lookup, guidance dampening,
setter defaults/quirks, API surface.  No CUDA compute is invoked."""
import inspect
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _tree():
    with open(os.path.join(GOLD, "tree.json")) as f:
        return json.load(f)


def _engine(turbo=False):
    from fakes import FakeHolder, fake_similarity
    from latentblending_b200 import BlendingEngine
    return BlendingEngine(None, holder=FakeHolder(turbo=turbo), similarity_fn=fake_similarity, run_benchmark=False)


def test_branching_sweep_matches_reference():
    be = _engine()
    for row in _tree()["branching_sweep"]:
        be.set_num_inference_steps(row["n"])
        idx, stems = be.get_time_based_branching(row["depth_strength"], **row["kw"])
        assert [int(v) for v in idx] == row["idx"], row
        assert [int(v) for v in stems] == row["stems"], row


def test_dampening_and_closest_idx_match_reference():
    be = _engine()
    t = _tree()
    for d in t["damp"]:
        be.set_guidance_mid_dampening(d["f"])
        assert abs(float(be.guidance_scale) - d["g"]) < 1e-12 and be.dh.guidance_scale == be.guidance_scale
    be.tree_fracts = [0.0, 0.25, 0.5, 0.75, 1.0]
    for c in t["closest"]:
        assert list(be.get_closest_idx(c["f"])) == c["idx"]


def test_defaults_and_quirks():
    base, turbo = _engine(False), _engine(True)
    assert base.guidance_scale == 4.0 and turbo.guidance_scale == 0.0
    assert base.num_inference_steps == 30 and turbo.num_inference_steps == 4
    # base model ignores user parental-crossfeed arguments (blending_engine.py:200-203)
    base.set_parental_crossfeed(0.9, 0.1, 0.1)
    assert (base.parental_crossfeed_power, base.parental_crossfeed_range, base.parental_crossfeed_decay) == (0.3, 0.6, 0.9)
    turbo.set_parental_crossfeed(0.5)
    assert (turbo.parental_crossfeed_power, turbo.parental_crossfeed_range, turbo.parental_crossfeed_decay) == (0.5, 1.0, 1.0)
    assert turbo.list_idx_injection == [2] and turbo.list_nmb_stems == [10]
    base.set_prompt1("a_b c")
    assert base.prompt1 == "a b c"
    with pytest.raises(ValueError):
        base.set_branching(t_compute_max_allowed=3.0, nmb_max_branches=5)
    with pytest.raises(AssertionError):
        turbo.set_branching(t_compute_max_allowed=3.0)
    with pytest.raises(AssertionError):
        from latentblending_b200 import BlendingEngine
        from fakes import FakeHolder
        BlendingEngine(None, guidance_scale_mid_damper=0.0, holder=FakeHolder())


def test_api_surface_matches_reference():
    from latentblending_b200 import BlendingEngine, DiffusersHolder
    want = {
        "set_dimensions": ["size_output"], "set_guidance_scale": ["guidance_scale"],
        "set_negative_prompt": ["negative_prompt"], "set_guidance_mid_dampening": ["fract_mixing"],
        "set_branch1_crossfeed": ["crossfeed_power", "crossfeed_range", "crossfeed_decay"],
        "set_parental_crossfeed": ["crossfeed_power", "crossfeed_range", "crossfeed_decay"],
        "set_prompt1": ["prompt"], "set_prompt2": ["prompt"], "set_image1": ["image"], "set_image2": ["image"],
        "set_num_inference_steps": ["num_inference_steps"],
        "set_branching": ["depth_strength", "t_compute_max_allowed", "nmb_max_branches"],
        "run_transition": ["recycle_img1", "recycle_img2", "fixed_seeds"],
        "compute_latents1": ["return_image"], "compute_latents2": ["return_image"],
        "compute_latents_mix": ["fract_mixing", "b_parent1", "b_parent2", "idx_injection"],
        "get_time_based_branching": ["depth_strength", "t_compute_max_allowed", "nmb_max_branches"],
        "get_mixing_parameters": ["idx_injection"], "insert_into_tree": ["fract_mixing", "idx_injection", "list_latents"],
        "write_imgs_transition": ["dp_img"], "write_movie_transition": ["fp_movie", "duration_transition", "fps"],
        "get_state_dict": [], "swap_forward": [], "get_lpips_similarity": ["imgA", "imgB"],
        "get_closest_idx": ["fract_mixing"], "benchmark_speed": [],
    }
    for name, args in want.items():
        fn = getattr(BlendingEngine, name)
        got = [p for p in inspect.signature(fn).parameters if p != "self"]
        assert got == args, (name, got)
    init = [p for p in inspect.signature(BlendingEngine.__init__).parameters][1:5]
    assert init == ["pipe", "do_compile", "guidance_scale_mid_damper", "mid_compression_scaler"]
    for name in ("get_text_embedding", "get_noise", "run_diffusion_sd_xl", "latent2image", "set_dimensions",
                 "set_negative_prompt", "set_num_inference_steps", "prepare_mixing"):
        assert hasattr(DiffusersHolder, name)
    sig = [p for p in inspect.signature(DiffusersHolder.run_diffusion_sd_xl).parameters][1:]
    assert sig == ["text_embeddings", "latents_start", "idx_start", "list_latents_mixing", "mixing_coeffs", "return_image"]


def test_scheduler_tables_known_answers():
    from latentblending_b200.schedulers import EulerTables
    s = EulerTables("euler")
    s.set_timesteps(30)
    assert s.timesteps[:2].tolist() == [958.0, 925.0]
    np.testing.assert_allclose(s.sigmas[[0, 15, 29, 30]].numpy(), [11.4769, 1.4316, 0.0413, 0.0], atol=6e-5)
    assert abs(float(s.init_noise_sigma) - 11.5203) < 1e-4
    t = EulerTables("euler_ancestral")
    t.set_timesteps(4)
    assert t.timesteps.tolist() == [999.0, 749.0, 499.0, 249.0]
    np.testing.assert_allclose([sc["sigma_up"] for sc in t.step_scalars], [3.9193, 1.4816, 0.6259, 0.0], atol=3e-3)
    # tables agree with the oracle's scheduler bit for bit after the fp16 cast
    from oracle.schedulers import EulerAncestralDiscrete, EulerDiscrete
    o = EulerDiscrete(); o.set_timesteps(30)
    assert np.array_equal(o.sigmas.numpy(), s.sigmas.numpy())
    for i, sc in enumerate(s.step_scalars):
        assert sc["sigma"] == float(o.sigmas[i].half()) and sc["dt"] == float((o.sigmas[i + 1] - o.sigmas[i]).half())
        assert sc["divisor"] == float(((o.sigmas[i] ** 2 + 1) ** 0.5).half())
    oa = EulerAncestralDiscrete(); oa.set_timesteps(4)
    for i, sc in enumerate(t.step_scalars):
        up, down = oa.sigma_up_down(i)
        assert sc["sigma_up"] == float(up.half()) and sc["dt"] == float((down - oa.sigmas[i]).half())


def test_batched_outer_pair_wiring_matches_sequential_order():
    """run_transition hands the two outer trajectories to the holder's lockstep multi-branch loop when it has one
    (DiffusersHolder.run_diffusion_sd_xl_multi).  With a fake holder whose multi loop simply runs its jobs one after
    the other (resolving the ("job", j) crossfeed reference), the trajectories and the per-branch call log must equal
    compute_latents1() followed by compute_latents2() -- with and without branch-1 crossfeed.  (The inner levels need
    the CUDA parental mix: the whole transition is compared on the GPU in tests/test_engine_gpu.py.)"""
    import torch
    from fakes import FakeHolder, fake_similarity
    from latentblending_b200 import BlendingEngine

    class MultiFake(FakeHolder):
        def run_diffusion_sd_xl_multi(self, jobs, idx_start=0):
            outs = []
            for job in jobs:
                m = job.get("list_latents_mixing")
                if isinstance(m, tuple) and m[0] == "job":
                    m = outs[m[1]]
                outs.append(self.run_diffusion_sd_xl(job["text_embeddings"], job["latents_start"], idx_start, m,
                                                     job.get("mixing_coeffs", 0.0)))
            return outs

    for crossfeed in (False, True):
        runs = []
        for batched in (False, True):
            holder = MultiFake()
            be = BlendingEngine(None, holder=holder, similarity_fn=fake_similarity, run_benchmark=False)
            be.set_num_inference_steps(10)
            be.set_prompt1("a lake")
            be.set_prompt2("a planet")
            if crossfeed:
                be.set_branch1_crossfeed(0.8, 0.6, 0.4)
            be.seed1, be.seed2 = 420, 421
            if batched:
                l1, l2 = be._compute_latents_pair()
            else:
                l1, l2 = be.compute_latents1(), be.compute_latents2()
            assert be.tree_latents[0] is l1 and be.tree_latents[-1] is l2
            runs.append((l1, l2, holder.calls))
        (a1, a2, ca), (b1, b2, cb) = runs
        assert ca == cb and len(ca) == 2
        if crossfeed:
            assert isinstance(cb[1]["coeffs"], list) and cb[1]["coeffs"][0] == 0.8 and cb[1]["n_mix_none"] == 0
        for x0, x1 in zip(a1 + a2, b1 + b2):
            assert torch.equal(x0, x1)
