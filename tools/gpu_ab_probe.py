#!/usr/bin/env python
"""Perf probe (test tooling): NCELoss fwd+bwd at 4096 x 768 with an engine option toggled back and forth in ONE process
(the boxes differ by several per cent, so A/B across runs says little).
    python tools/gpu_ab_probe.py <option> <value A> <value B> [rounds] [n] [dim]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import drs_b200 as drs  # noqa: E402

opt, va, vb = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
rounds = int(sys.argv[4]) if len(sys.argv) > 4 else 4
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1337)
n = int(sys.argv[5]) if len(sys.argv) > 5 else 4096
dim = int(sys.argv[6]) if len(sys.argv) > 6 else 768
unit = lambda x: torch.nn.functional.normalize(x, dim=1)
q = unit(torch.randn(n, dim, generator=g, device=dev)).requires_grad_(True)
k = unit(torch.randn(n, dim, generator=g, device=dev) * 0.5 + q.detach()).requires_grad_(True)
crit = drs.NCELoss({"temperature": 0.05})


def step():
    q.grad = None
    k.grad = None
    crit(q, k, None).backward()


def timed(reps=30):
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


res = {va: [], vb: []}
for _ in range(rounds):
    for v in (va, vb):
        drs.set_option(opt, v)
        res[v].append(timed())
for v in (va, vb):
    print(f"n={n} dim={dim} {opt}={v}: " + " ".join(f"{t * 1e3:.1f}" for t in res[v]) + f" us/step (min {min(res[v]) * 1e3:.1f})")
