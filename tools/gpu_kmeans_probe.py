#!/usr/bin/env python
"""Perf probe (test tooling): run_kmeans at the reference's clustering config (config.yaml: num_cluster
[4096, 6144, 8192], niter 20, nredo 5, max_points_per_centroid 1000) over 200 000 x 128 embeddings, and the cost of the
pieces of one Lloyd iteration at 8192 centroids."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import drs_b200 as drs  # noqa: E402
from drs_b200 import clustering  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1337)
n, d = 200_000, 128
centers = torch.nn.functional.normalize(torch.randn(5000, d, generator=g, device=dev), dim=1)
x = torch.nn.functional.normalize(centers[torch.randint(0, 5000, (n,), generator=g, device=dev)] + 0.08 * torch.randn(n, d, generator=g, device=dev), dim=1)
cfg = {"temperature": 0.05, "cluster": {"num_cluster": [4096, 6144, 8192], "num_neg_proto": 3072, "verbose": False, "niter": 20, "nredo": 5,
                                        "max_points_per_centroid": 1000, "min_points_per_centroid": 1}}
drs.run_kmeans({"temperature": 0.05, "cluster": dict(cfg["cluster"], num_cluster=[64], niter=2, nredo=1)}, x, device=dev)   # warm-up
torch.cuda.synchronize()
t0 = time.perf_counter()
res = drs.run_kmeans(cfg, x, device=dev)
torch.cuda.synchronize()
t = time.perf_counter() - t0
print(f"run_kmeans {n} x {d}, clusters {cfg['cluster']['num_cluster']}, niter 20 x nredo 5 = 300 Lloyd iterations: {t:.2f} s "
      f"({t / 300 * 1e3:.2f} ms per iteration); cluster sizes at 8192: min {int(torch.bincount(res['emb2cluster'][2], minlength=8192).min())}, "
      f"max {int(torch.bincount(res['emb2cluster'][2], minlength=8192).max())}")


def timed(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


cent = x[torch.randperm(n, device=dev)[:8192]].clone()
dist, assign = drs.flat_l2_search(x, cent, 1)
a = assign[:, 0].contiguous()
order, offsets, counts = clustering._runs(a, 8192)
c2 = cent.clone()
print(f"one iteration at 8192 centroids: assignment {timed(lambda: drs.flat_l2_search(x, cent, 1)):.3f} ms, "
      f"sort by cluster {timed(lambda: clustering._runs(a, 8192)):.3f} ms, "
      f"centroid update kernel {timed(lambda: clustering._cluster_update(x, order, offsets, 8192, centroids=c2)):.3f} ms "
      f"({n * d * 4 / 1e6:.0f} MB of samples)")
