"""One small end-to-end invocation of the hot path on cuda:0, checked against the
CPU oracle (called by __graft_entry__.smoke())."""
import torch


def run():
    from latentblending_b200 import ops, utils
    from oracle import mixing                         # checker only
    from oracle.schedulers import EulerDiscrete
    g = torch.Generator().manual_seed(0)
    p0 = torch.randn(1, 4, 64, 64, generator=g).half()
    p1 = torch.randn(1, 4, 64, 64, generator=g).half()
    out = utils.interpolate_spherical(p0.cuda(), p1.cuda(), 0.3).cpu()
    assert torch.equal(out, mixing.interpolate_spherical(p0, p1, 0.3)), "slerp mismatch vs oracle"
    s = EulerDiscrete()
    s.set_timesteps(30)
    eps = torch.randn(2, 4, 64, 64, generator=g).half()
    e = eps[:1] + 4.0 * (eps[1:] - eps[:1])
    ref = s.step(e, 3, p0)
    got = ops.cfg_euler_step(p0.cuda(), eps.cuda(), 4.0, float(s.sigmas[3]), float(s.sigmas[4] - s.sigmas[3])).cpu()
    assert torch.equal(got, ref), "euler step mismatch vs oracle"
    torch.cuda.synchronize()
    print("smoke ok: slerp + cfg/euler step bit-exact vs oracle")
