#!/usr/bin/env python
"""Perf probe (test tooling): NCELoss forward + backward at the reference's OWN shapes (config.yaml: batch 128, 128-d,
MoCo queue 12 544, T = 0.05), eager and replayed from a CUDA graph, next to the closed form in torch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import drs_b200 as drs  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1337)
unit = lambda x: torch.nn.functional.normalize(x, dim=1)
n, dim, klen, temp = 128, 128, 12544, 0.05
q = unit(torch.randn(n, dim, generator=g, device=dev)).requires_grad_(True)
k = unit(torch.randn(n, dim, generator=g, device=dev) * 0.5 + q.detach()).requires_grad_(True)
queue = torch.nn.functional.normalize(torch.randn(dim, klen, generator=g, device=dev), dim=0)
crit = drs.NCELoss({"temperature": temp})


def step():
    q.grad = None
    k.grad = None
    crit(q, k, queue).backward()


def timed(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


eager = timed(step)
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    step()
torch.cuda.current_stream().wait_stream(side)
q.grad = k.grad = None
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    crit(q, k, queue).backward()
replay = timed(graph.replay)
f = torch.cat([q.detach(), k.detach()]).requires_grad_(True)


def torch_step():
    f.grad = None
    s = (f @ f.T) / temp
    s = s.masked_fill(torch.eye(2 * n, dtype=torch.bool, device=dev), float("-inf"))
    lq = (f[:n] @ queue) / temp
    logits = torch.cat([s, torch.cat([lq, lq])], 1)
    tgt = (torch.arange(2 * n, device=dev) + n) % (2 * n)
    (torch.nn.functional.cross_entropy(logits, tgt, reduction="sum") / 2).backward()


crit32 = drs.NCELoss({"temperature": temp, "precision": "fp32"})


def step32():
    q.grad = None
    k.grad = None
    crit32(q, k, queue).backward()


if os.environ.get("REFSHAPE_BRIEF"):
    print(f"NCELoss fwd+bwd at the reference shapes (N={n}, D={dim}, queue {klen}): eager {eager * 1e3:.1f} us, CUDA-graph replay {replay * 1e3:.1f} us")
    sys.exit(0)
fp32_split = timed(step32)
drs.set_option("tune.k_split", 1)
fp32_plain = timed(step32, 10)
bf16_plain = timed(step)
drs.set_option("tune.k_split", 0)
print(f"  precision='fp32' (FFMA kernels): eager {fp32_split * 1e3:.1f} us; without the split-K of dq = Hq x queue: fp32 {fp32_plain * 1e3:.1f} us, "
      f"bf16 eager {bf16_plain * 1e3:.1f} us")
print(f"NCELoss fwd+bwd at the reference shapes (N={n}, D={dim}, queue {klen}): eager {eager * 1e3:.1f} us, CUDA-graph replay {replay * 1e3:.1f} us, "
      f"torch closed form {timed(torch_step) * 1e3:.1f} us")
