"""CPU oracle: the prototype clustering around the flat-L2 search.  TEST INFRASTRUCTURE.

``run_kmeans`` (src/contrastor/utils.py:50-105) does three things per cluster count:

1. trains k-means with ``faiss.Clustering`` on a ``faiss.GpuIndexFlatL2`` (:28-36, :39-47, :61-64).  faiss is a
   third-party dependency that is NOT vendored, NOT pinned (requirements.txt says just ``faiss``) and not installed
   here, and its random initialisation and empty-cluster splitting draw from its own RNG: PARITY UNPINNED for the
   training itself.  ``lloyd`` below is the published algorithm (Lloyd iterations from k distinct seeded points:
   exact squared-L2 assignment with ties to the lower index, centroid = mean of its points, an empty cluster takes
   the place next to the currently largest one, both perturbed by +-1/1024 per coordinate like faiss's
   ``split_clusters``) and is what the GPU implementation is compared with from the same starting centroids;
2. assigns every sample to its nearest centroid: ``D, I = index.search(x, 1)`` (:67-68);
3. turns the per-cluster distances into the "concentration" temperatures of ProtoNCE (:73-94) -- the reference's OWN
   numpy code, restated literally in ``density`` and pinned by tests/golden/kmeans_density.npz, which was produced
   by running the reference's ``run_kmeans`` itself over a numpy stand-in for the faiss objects
   (tests/golden/make_golden.py::gen_kmeans).
"""
from __future__ import annotations

import numpy as np

SPLIT_EPS = 1.0 / 1024.0


def assign(x: np.ndarray, centroids: np.ndarray):
    """Exact squared L2 to every centroid, float64; nearest with ties -> lower index.  Returns (D float32, I int64)."""
    x64, c64 = x.astype(np.float64), centroids.astype(np.float64)
    d = (x64 * x64).sum(1)[:, None] + (c64 * c64).sum(1)[None, :] - 2.0 * x64 @ c64.T
    idx = d.argmin(1)                                  # first minimum = lower index
    return np.maximum(d[np.arange(len(x)), idx], 0.0).astype(np.float32), idx.astype(np.int64)


def init_centroids(x: np.ndarray, k: int, seed: int) -> np.ndarray:
    """k distinct sample rows, chosen by a seeded permutation (faiss: ``rand_perm(n, seed)[:k]``; the generator differs)."""
    perm = np.random.RandomState(seed).permutation(len(x))[:k]
    return x[perm].astype(np.float32).copy()


def update(x: np.ndarray, idx: np.ndarray, centroids: np.ndarray):
    """Centroid = mean of its members; an empty cluster is re-seeded next to the largest one (see the module header)."""
    k, _ = centroids.shape
    counts = np.bincount(idx, minlength=k).astype(np.int64)
    new = centroids.astype(np.float64).copy()
    sums = np.zeros_like(new)
    np.add.at(sums, idx, x.astype(np.float64))
    live = counts > 0
    new[live] = sums[live] / counts[live][:, None]
    new = new.astype(np.float32)
    nsplit = 0
    for ci in np.flatnonzero(~live):
        cj = int(counts.argmax())                      # ties -> lower index
        sign = np.where(np.arange(new.shape[1]) % 2 == 0, 1.0, -1.0).astype(np.float32)
        base = new[cj].copy()
        new[ci] = base * (1.0 + sign * np.float32(SPLIT_EPS))
        new[cj] = base * (1.0 - sign * np.float32(SPLIT_EPS))
        counts[ci] = counts[cj] // 2
        counts[cj] -= counts[ci]
        nsplit += 1
    return new, nsplit


def lloyd(x: np.ndarray, centroids: np.ndarray, niter: int):
    """niter Lloyd iterations from `centroids`.  Returns (centroids float32, [objective per iteration])."""
    objective = []
    c = centroids.astype(np.float32).copy()
    for _ in range(niter):
        d, idx = assign(x, c)
        objective.append(float(d.astype(np.float64).sum()))
        c, _ = update(x, idx, c)
    return c, objective


def density(D: np.ndarray, I: np.ndarray, k: int, temperature: float) -> np.ndarray:
    """src/contrastor/utils.py:73-94, line by line.  D, I: [n, 1] as ``index.search(x, 1)`` returns them."""
    emb2cluster = [int(n[0]) for n in I]                                      # :68
    Dcluster = [[] for _ in range(k)]                                         # :74
    for nis, i in enumerate(emb2cluster):                                     # :75-76
        Dcluster[i].append(D[nis][0])
    dens = np.zeros(k)                                                        # :79
    for i, dist in enumerate(Dcluster):                                       # :80-83
        if len(dist) > 1:
            dens[i] = (np.asarray(dist) ** 0.5).mean() / np.log(len(dist) + 10)
    dmax = dens.max()                                                         # :86
    for i, dist in enumerate(Dcluster):                                       # :87-89
        if len(dist) <= 1:
            dens[i] = dmax
    dens = dens.clip(np.percentile(dens, 10), np.percentile(dens, 90))        # :91-92
    return temperature * dens / dens.mean()                                   # :93-94
