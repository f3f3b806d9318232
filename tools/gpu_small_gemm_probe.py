#!/usr/bin/env python
"""Perf probe (test tooling): the 8192 x 8192 x 768 shape of the InfoNCE GEMMs through the search scan, with
parts of the pipeline disabled -- separates fixed launch/pipeline cost from epilogue cost at ~100 us scale."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import drs_b200 as drs  # noqa: E402


def scan_ms(q, c, k, iters=20):
    prof = []
    for _ in range(3):
        drs.search(q, c, k, profile=prof)
    prof.clear()
    for _ in range(iters):
        drs.search(q, c, k, profile=prof)
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in prof) / len(prof)


def main():
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(1)
    for n in (8192, 16384):
        f = torch.nn.functional.normalize(torch.randn(n, 768, generator=g, device=dev), dim=1).bfloat16()
        fl = 2.0 * n * n * 768
        for splits in (0, 16, 8):
            drs.set_option("search.splits", splits)
            for flags, label in ((0, "full"), (1, "no_functor"), (7, "tma_only")):
                drs.set_option("debug.flags", flags)
                ms = scan_ms(f, f, 10)
                print(f"{n}x{n}x768 splits={splits:2d} {label:12s} {ms * 1e3:8.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s", flush=True)
        drs.set_option("debug.flags", 0)
        drs.set_option("search.splits", 0)
        for rb in (1, 0):
            drs.set_option("tune.round_barrier", rb)
            ms = scan_ms(f, f, 10)
            print(f"{n}x{n}x768 round_barrier={rb} full {ms * 1e3:8.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s", flush=True)
        drs.set_option("tune.round_barrier", 1)
        out = torch.empty(n, n, dtype=torch.bfloat16, device=dev)
        torch.matmul(f, f.T, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            torch.matmul(f, f.T, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"{n}x{n}x768 cuBLAS bf16   {ms * 1e3:8.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
