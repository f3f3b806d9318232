#!/usr/bin/env python
"""Perf probe (test tooling): fixed per-search latency -- small batches over small corpora, where the corpus
stream takes tens of microseconds and launch/staging overhead is what is left."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import drs_b200 as drs  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(1)
    for nc in (100_000, 1_000_000):
        c = torch.nn.functional.normalize(torch.randn(nc, 768, generator=g, device=dev), dim=1).bfloat16()
        for nq in (1, 16, 128):
            q = torch.nn.functional.normalize(torch.randn(nq, 768, generator=g, device=dev), dim=1).bfloat16()
            for _ in range(5):
                drs.search(q, c, 10)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            for _ in range(200):
                drs.search(q, c, 10)
            e1.record()
            t_host = (time.perf_counter() - t0) / 200
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 200
            # one search with a sync on both sides: what a caller waiting for the answer sees
            lat = []
            for _ in range(20):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                s, i = drs.search(q, c, 10)
                torch.cuda.synchronize()
                lat.append(time.perf_counter() - t0)
            # the same search captured once in a CUDA graph and replayed
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                gs, gi = drs.search(q, c, 10)
            glat = []
            for _ in range(20):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                graph.replay()
                torch.cuda.synchronize()
                glat.append(time.perf_counter() - t0)
            e0.record()
            for _ in range(200):
                graph.replay()
            e1.record()
            torch.cuda.synchronize()
            gms = e0.elapsed_time(e1) / 200
            ideal = nc * 768 * 2 / 6.55e12
            print(f"nq={nq:4d} nc={nc:8d}: {ms * 1e3:7.1f} us/search back-to-back (host enqueue {t_host * 1e6:6.1f} us), "
                  f"sync latency {sorted(lat)[len(lat) // 2] * 1e6:7.1f} us | CUDA graph: {gms * 1e3:7.1f} us back-to-back, "
                  f"sync latency {sorted(glat)[len(glat) // 2] * 1e6:7.1f} us | corpus stream alone {ideal * 1e6:6.1f} us", flush=True)


if __name__ == "__main__":
    main()
