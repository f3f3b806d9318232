#!/usr/bin/env python
"""Perf probe (test tooling): where the fixed cost of a small-batch scan over a 1M-row shard goes."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import drs_b200 as drs  # noqa: E402


def scan_us(q, c, k, iters=50):
    prof = []
    for _ in range(5):
        drs.search(q, c, k, profile=prof)
    prof.clear()
    for _ in range(iters):
        drs.search(q, c, k, profile=prof)
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in prof) / len(prof) * 1e3


def main():
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(1)
    nc = 1_000_000
    c = torch.nn.functional.normalize(torch.randn(nc, 768, generator=g, device=dev), dim=1).bfloat16()
    print(f"corpus stream alone: {nc * 1536 / 6.551e12 * 1e6:.0f} us")
    for nq in (1, 16, 128):
        q = torch.nn.functional.normalize(torch.randn(nq, 768, generator=g, device=dev), dim=1).bfloat16()
        line = f"nq={nq:4d}:"
        for flags, label in ((0, "full"), (1, "no_functor"), (3, "no_tmem_ld"), (7, "tma_only")):
            drs.set_option("debug.flags", flags)
            line += f"  {label} {scan_us(q, c, 10):6.1f} us"
        drs.set_option("debug.flags", 0)
        for opt, val, label in (("tune.seed_thresholds", 0, "no_seeds"), ("tune.round_barrier", 0, "no_barrier")):
            drs.set_option(opt, val)
            line += f"  {label} {scan_us(q, c, 10):6.1f} us"
            drs.set_option(opt, 1)
        print(line, flush=True)


if __name__ == "__main__":
    main()
