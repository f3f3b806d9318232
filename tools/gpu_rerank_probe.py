#!/usr/bin/env python
"""Perf probe (test tooling): the HBM-bound gathers -- candidate re-rank (report 3.2: sparse top-100 -> dense top-15) and
paired scores -- as achieved GB/s of the rows they must read."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import drs_b200 as drs  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
nc, dim = 5_400_000, 768
c = torch.empty(nc, dim, dtype=torch.bfloat16, device=dev)
for r0 in range(0, nc, 1 << 20):
    r1 = min(nc, r0 + (1 << 20))
    c[r0:r1] = torch.nn.functional.normalize(torch.randn(r1 - r0, dim, generator=g, device=dev), dim=1)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for nq, m, k in ((10000, 100, 15), (65536, 100, 15), (1000, 100, 15)):
    q = torch.nn.functional.normalize(torch.randn(nq, dim, generator=g, device=dev), dim=1).bfloat16()
    cand = torch.randint(0, nc, (nq, m), generator=g, device=dev)
    ms = timed(lambda: drs.rerank(q, c, cand, k))
    byt = nq * m * dim * 2 + nq * m * 8 + nq * dim * 2
    print(f"rerank {nq} claims x {m} candidates x {dim} bf16 -> top-{k}: {ms:.3f} ms, {byt / ms / 1e6:.0f} GB/s of gathered rows "
          f"({byt / ms / 1e6 / 6551 * 100:.0f} % of the measured copy bandwidth)", flush=True)
a, b = c[:4_000_000], c[1_000_000:5_000_000]
ms = timed(lambda: drs.paired_scores(a, b))
byt = 2 * a.numel() * 2
print(f"paired_scores 4M x {dim} bf16: {ms:.3f} ms, {byt / ms / 1e6:.0f} GB/s ({byt / ms / 1e6 / 6551 * 100:.0f} %)", flush=True)
