"""Probe (on the GPU box) how PyTorch CUDA binary ops treat an fp32 0-dim CUDA
tensor operand next to an fp16 tensor: kept at fp32 precision ("fp32-scalar") or
rounded to fp16 first ("fp16-scalar")?  This decides the rounding chain the
reference's scheduler arithmetic has on its real (CUDA) stack, which the oracle
and the fused step kernel must reproduce."""
import torch

torch.manual_seed(0)
x = torch.randn(200000).half().cuda()
ops = {"*": lambda p, q: p * q, "/": lambda p, q: p / q, "-": lambda p, q: p - q, "+": lambda p, q: p + q}
print("torch", torch.__version__)
for sv in (11.4769, 6.7684, 0.4179):
    s = torch.tensor(sv, device="cuda")           # 0-dim fp32 CUDA tensor, like scheduler.sigmas[i]
    for nm in ("x*s", "s*x", "x/s", "s/x", "x-s", "s-x", "x+s", "s+x"):
        f = ops[nm[1]]
        if nm[0] == "x":
            a = f(x, s)
            b = f(x.float(), s).half()
            c = f(x.float(), s.half().float()).half()
        else:
            a = f(s, x)
            b = f(s, x.float()).half()
            c = f(s.half().float(), x.float()).half()
        print(f"sigma={sv} {nm}: dtype={a.dtype} mismatches vs fp32-scalar={int((a != b).sum())} "
              f"vs fp16-scalar={int((a != c).sum())}")
# python-scalar (wrapped number) for comparison
a = 3.37 * x
print("py-scalar*x: vs fp32", int((a != (3.37 * x.float()).half()).sum()),
      "vs fp16", int((a != (float(torch.tensor(3.37).half()) * x.float()).half()).sum()))
