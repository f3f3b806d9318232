"""Perf probe (test tooling): NCELoss step time with the gradient-of-logits stores / all epilogue functors disabled."""
import os, sys
sys.path.insert(0, '/root/repo')
import torch, drs_b200 as drs
dev = torch.device('cuda:0'); n, dim = 4096, 768
g = torch.Generator(device=dev).manual_seed(1337)
q = torch.nn.functional.normalize(torch.randn(n, dim, generator=g, device=dev), dim=1).requires_grad_(True)
k = torch.nn.functional.normalize(torch.randn(n, dim, generator=g, device=dev) * 0.5 + q.detach(), dim=1).requires_grad_(True)
crit = drs.NCELoss({"temperature": 0.05})
def step():
    q.grad = None; k.grad = None
    crit(q, k, None).backward()
for flags in (0, 8, 1):   # 8: gradient-of-logits stores off, 1: every epilogue functor off
    drs.set_option("debug.flags", flags)
    for _ in range(5): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): step()
    e1.record(); torch.cuda.synchronize()
    print(f"debug.flags={flags:2d}: {e0.elapsed_time(e1) / 50 * 1e3:7.1f} us/step", flush=True)
drs.set_option("debug.flags", 0)
