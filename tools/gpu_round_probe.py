#!/usr/bin/env python
"""Perf probe (test tooling): the round barrier -- 10 000 claims over an 8-GPU-sized shard (3.125 M rows) and over 12 M
rows, with and without it.  (Producers allowed to run 1 or 2 rounds ahead of the slowest were also measured here:
37.0 ms against 33.4 ms in strict lockstep and 36.0 ms with no barrier, so that knob was removed.)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import drs_b200 as drs  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
nc = 12_000_000
c = torch.empty(nc, 768, dtype=torch.bfloat16, device=dev)
for r0 in range(0, nc, 1 << 20):
    r1 = min(nc, r0 + (1 << 20))
    c[r0:r1] = torch.nn.functional.normalize(torch.randn(r1 - r0, 768, generator=g, device=dev), dim=1)
q = torch.nn.functional.normalize(torch.randn(10000, 768, generator=g, device=dev), dim=1).bfloat16()


def scan_ms(corpus, iters=6):
    prof = []
    for _ in range(2):
        drs.search(q, corpus, 10, profile=prof)
    prof.clear()
    for _ in range(iters):
        drs.search(q, corpus, 10, profile=prof)
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in prof) / len(prof)


for rows in (3_125_000, 12_000_000):
    cc = c[:rows]
    line = f"rows={rows}:"
    for name, opts in (("barrier", {"tune.round_barrier": 1}), ("nobarrier", {"tune.round_barrier": 0}), ("barrier again", {"tune.round_barrier": 1})):
        for k, v in opts.items():
            drs.set_option(k, v)
        ms = scan_ms(cc)
        fl = 2.0 * 10000 * rows * 768
        line += f"  {name} {ms:.3f} ms ({fl / ms / 1e9:.0f} TF)"
        drs.set_option("tune.round_barrier", 1)
    print(line, flush=True)
