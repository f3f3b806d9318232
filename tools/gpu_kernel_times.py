#!/usr/bin/env python
"""Perf probe (test tooling): per-kernel device times of the NCELoss step at 4096 x 768 from CUPTI activity records
(torch.profiler) -- warm caches and back-to-back launches, unlike an ncu launch list (which flushes the caches and
serialises: the small latency-bound kernels read 2-3 x slower there).
    python tools/gpu_kernel_times.py [n] [dim] [queue_len]      (DRS_OPTIONS=name=value,... sets engine options)"""
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import drs_b200 as drs  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 768
klen = int(sys.argv[3]) if len(sys.argv) > 3 else 0
for name in os.environ.get("DRS_OPTIONS", "").split(","):
    if "=" in name:
        drs.set_option(name.split("=")[0], int(name.split("=")[1]))
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1337)
unit = lambda x: torch.nn.functional.normalize(x, dim=1)
q = unit(torch.randn(n, dim, generator=g, device=dev)).requires_grad_(True)
k = unit(torch.randn(n, dim, generator=g, device=dev) * 0.5 + q.detach()).requires_grad_(True)
queue = torch.nn.functional.normalize(torch.randn(dim, klen, generator=g, device=dev), dim=0) if klen else None
crit = drs.NCELoss({"temperature": 0.05})


def step():
    q.grad = None
    k.grad = None
    crit(q, k, queue).backward()


for _ in range(5):
    step()
torch.cuda.synchronize()
reps = 20
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(reps):
        step()
    torch.cuda.synchronize()
tot = defaultdict(float)
cnt = defaultdict(int)
order = []
for ev in prof.events():
    if ev.device_type.name != "CUDA":
        continue
    name = ev.name
    if name not in tot:
        order.append(name)
    tot[name] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
    cnt[name] += 1
total = 0.0
for name in order:
    per_step = tot[name] / reps
    total += per_step
    print(f"{per_step:8.1f} us/step  x{cnt[name] / reps:4.1f}  {name[:110]}")
print(f"{total:8.1f} us/step  sum of kernel times")
