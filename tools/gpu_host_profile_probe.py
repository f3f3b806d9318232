#!/usr/bin/env python
"""Perf probe (test tooling): host time of the NCELoss module at the reference's shapes (N = 128, D = 128, queue 12 544) --
forward and backward separately, then a cProfile of 300 forward + backward steps."""
import cProfile
import io
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import drs_b200 as drs  # noqa: E402
n, dim, klen = 128, 128, 12544
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
unit = lambda x: torch.nn.functional.normalize(x, dim=1)
q = unit(torch.randn(n, dim, generator=g, device=dev)).requires_grad_(True)
k = unit(torch.randn(n, dim, generator=g, device=dev)).requires_grad_(True)
queue = torch.nn.functional.normalize(torch.randn(dim, klen, generator=g, device=dev), dim=0)
crit = drs.NCELoss({"temperature": 0.05, "precision": "bf16"})
def step():
    q.grad = None; k.grad = None
    crit(q, k, queue).backward()
for _ in range(50): step()
torch.cuda.synchronize()
# split: forward only with grad, then backward
t0=time.perf_counter()
for _ in range(300):
    q.grad=None;k.grad=None
    l = crit(q,k,queue)
t1=time.perf_counter()
print("forward with grad (host us):", (t1-t0)/300*1e6)
torch.cuda.synchronize()
ls=[]
for _ in range(100):
    ls.append(crit(q,k,queue))
torch.cuda.synchronize()
t0=time.perf_counter()
for l in ls: l.backward()
t1=time.perf_counter()
print("backward (host us):", (t1-t0)/100*1e6)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(300): step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(25)
print(s.getvalue()[:6000])
