"""Prototype clustering for ProtoNCE: the caller side of the flat-L2 search.

Mirrors what ``run_kmeans`` (src/contrastor/utils.py:50-105) builds out of faiss -- ``faiss.Clustering`` (:28-36),
``faiss.GpuIndexFlatL2`` (:39-47, here ``retrieval.FlatL2Index``), ``faiss.vector_to_array`` (:71) -- so that the
function runs with ``import drs_b200 as faiss``-style substitutions, plus ``run_kmeans`` itself with the same
config keys and the same result dictionary (``emb2cluster``, ``centroids``, ``density``: what
``NCELoss._compute_proto_loss`` consumes, contrastive_loss.py:95-135).

Every iteration is two launches of this engine: the assignment (``drs_search_l2``: the scan kernel with a distance
epilogue, fp32 operands on the tensor cores) and ``drs_cluster_update`` (csrc/kmeans.cuh: per-cluster means in
fp64, deterministic); a stable sort by cluster in between is torch plumbing.

PARITY: faiss is not vendored, pinned or installed, and its initial centroids and empty-cluster splits come from its own
random generator, so the TRAINING is "parity unpinned" (compared with oracle/kmeans.py's Lloyd iteration from the same
start).  The concentration estimate (:73-94) is the reference's own code and is pinned by a golden fixture produced by
running the reference (tests/golden/kmeans_density.npz).  CUDA only: there is no CPU path.
"""
from __future__ import annotations

import math
import warnings
from typing import Optional

import numpy as np
import torch

from . import _lib
from .retrieval import FlatL2Index, flat_l2_search

SPLIT_EPS = 1.0 / 1024.0     # faiss Clustering.cpp: EPS of split_clusters


def vector_to_array(v) -> np.ndarray:
    """``faiss.vector_to_array`` as used at src/contrastor/utils.py:71 (``clus.centroids`` -> flat float32 array)."""
    if isinstance(v, torch.Tensor):
        return v.detach().cpu().numpy().reshape(-1)
    return np.asarray(v).reshape(-1)


def _runs(assign: torch.Tensor, k: int):
    """Stable sort of the samples by cluster: (order int64 [n], offsets int64 [k + 1], counts int64 [k])."""
    if assign.numel() and (int(assign.min()) < 0 or int(assign.max()) >= k):
        raise ValueError(f"cluster ids must lie in [0, {k}), got [{int(assign.min())}, {int(assign.max())}]")
    order = torch.sort(assign, stable=True).indices
    counts = torch.bincount(assign, minlength=k)
    offsets = torch.zeros(k + 1, dtype=torch.int64, device=assign.device)
    offsets[1:] = torch.cumsum(counts, 0)
    return order.contiguous(), offsets, counts


def _cluster_update(x: Optional[torch.Tensor], order, offsets, k: int, dist=None, centroids=None, sum_sqrt=None):
    lib = _lib.load()
    dev = order.device
    with torch.cuda.device(dev):
        _lib.check(lib.drs_cluster_update(x.data_ptr() if x is not None else None, order.shape[0],
                                          x.shape[1] if x is not None else 1, order.data_ptr(), offsets.data_ptr(), k,
                                          dist.data_ptr() if dist is not None else None,
                                          centroids.data_ptr() if centroids is not None else None,
                                          sum_sqrt.data_ptr() if sum_sqrt is not None else None,
                                          torch.cuda.current_stream(dev).cuda_stream))


def update_centroids(x: torch.Tensor, assign: torch.Tensor, centroids: torch.Tensor) -> int:
    """One centroid update in place: mean of the members; an empty cluster is re-seeded next to the currently largest
    one, both moved apart by +-1/1024 per coordinate (faiss's ``split_clusters`` picks the donor at random in
    proportion to its size; taking the largest keeps the step deterministic).  Returns the number of splits."""
    if not (x.is_cuda and assign.is_cuda and centroids.is_cuda):
        raise RuntimeError("update_centroids needs CUDA tensors: there is no CPU path")
    if x.dtype != torch.float32 or centroids.dtype != torch.float32 or not x.is_contiguous() or not centroids.is_contiguous():
        raise ValueError("x and centroids must be contiguous float32 matrices")
    if x.dim() != 2 or centroids.dim() != 2 or x.shape[1] != centroids.shape[1] or assign.shape != (x.shape[0],):
        raise ValueError(f"shape mismatch: x {tuple(x.shape)}, assign {tuple(assign.shape)}, centroids {tuple(centroids.shape)}")
    k = centroids.shape[0]
    order, offsets, counts = _runs(assign.to(torch.int64).contiguous(), k)
    _cluster_update(x, order, offsets, k, centroids=centroids)
    empty = (counts == 0).nonzero().flatten()
    if empty.numel() == 0:
        return 0
    counts = counts.clone()
    sign = torch.where(torch.arange(centroids.shape[1], device=centroids.device) % 2 == 0, 1.0, -1.0).to(centroids.dtype)
    for ci in empty.tolist():
        cj = int(torch.argmax(counts))                # first maximum = lower index
        base = centroids[cj].clone()
        centroids[ci] = base * (1.0 + sign * SPLIT_EPS)
        centroids[cj] = base * (1.0 - sign * SPLIT_EPS)
        counts[ci] = counts[cj] // 2
        counts[cj] -= counts[ci]
    return int(empty.numel())


class Clustering:
    """The slice of ``faiss.Clustering`` that ``get_cluster`` / ``run_kmeans`` use (src/contrastor/utils.py:28-36, :64,
    :71): attributes ``niter``, ``nredo``, ``seed``, ``verbose``, ``max_points_per_centroid``,
    ``min_points_per_centroid`` (faiss's defaults), ``train(x, index)``, ``centroids`` (flat float32 array)."""

    def __init__(self, d: int, k: int):
        self.d, self.k = int(d), int(k)
        self.niter = 25
        self.nredo = 1
        self.seed = 1234
        self.verbose = False
        self.max_points_per_centroid = 256
        self.min_points_per_centroid = 39
        self.centroids = None            # flat float32 numpy array [k * d] after train()
        self.centroids_tensor = None     # the same, [k, d] on the device
        self.objective = []              # sum of squared distances before each update of the best run

    def train(self, x, index: FlatL2Index):
        dev = index.device
        x = torch.as_tensor(x).to(device=dev, dtype=torch.float32).contiguous()
        if x.dim() != 2 or x.shape[1] != self.d:
            raise ValueError(f"train expects [n, {self.d}] vectors, got {tuple(x.shape)}")
        n, k = x.shape[0], self.k
        if n < k:
            raise RuntimeError(f"Number of training points ({n}) should be at least as large as number of clusters ({k})")
        gen = torch.Generator().manual_seed(int(self.seed))
        if n > k * self.max_points_per_centroid:                      # faiss subsamples the training set
            keep = torch.randperm(n, generator=gen)[: k * self.max_points_per_centroid].sort().values.to(dev)
            x = x[keep].contiguous()
            n = x.shape[0]
        elif n < k * self.min_points_per_centroid:
            warnings.warn(f"clustering {n} points to {k} centroids: please provide at least {k * self.min_points_per_centroid} training points")
        given = None                                                  # faiss: centroids set before train() are the starting point
        if self.centroids is not None:
            given = torch.as_tensor(self.centroids).reshape(-1)
            if given.numel() != k * self.d:
                raise ValueError(f"initial centroids must hold {k} x {self.d} values, got {given.numel()}")
            given = given.reshape(k, self.d).to(device=dev, dtype=torch.float32)
        best = None
        for redo in range(1 if given is not None else max(1, int(self.nredo))):
            g = torch.Generator().manual_seed(int(self.seed) + redo)
            cent = given.clone() if given is not None else x[torch.randperm(n, generator=g)[:k].to(dev)].clone()
            objective = []
            for it in range(int(self.niter)):
                dist, assign = flat_l2_search(x, cent, 1)
                objective.append(dist.sum(dtype=torch.float64))
                nsplit = update_centroids(x, assign[:, 0].contiguous(), cent)
                if self.verbose:
                    print(f"  redo {redo} iteration {it}: objective {objective[-1].item():.6g}, {nsplit} empty clusters split")
            final = float(objective[-1]) if objective else math.inf
            if best is None or final < best[0]:
                best = (final, cent, [float(o) for o in objective])
        _, cent, self.objective = best
        self.centroids_tensor = cent
        self.centroids = cent.cpu().numpy().reshape(-1)
        index.reset()                                                  # faiss leaves the trained centroids in the index (:67 searches it)
        index.add(cent)


def cluster_density(D, I, k: int, temperature: float) -> torch.Tensor:
    """The concentration estimate of src/contrastor/utils.py:73-94 from ``D, I = index.search(x, 1)``: per cluster
    mean(sqrt(dist)) / log(size + 10); clusters with <= 1 point take the maximum; clamp to the 10th..90th percentile;
    rescale so that the mean is ``temperature``.  Returns float32 [k] on the device."""
    D = torch.as_tensor(D)
    I = torch.as_tensor(I)
    dev = D.device if D.is_cuda else (I.device if I.is_cuda else torch.device("cuda", torch.cuda.current_device()))
    dist = D.reshape(-1).to(device=dev, dtype=torch.float32).contiguous()
    assign = I.reshape(-1).to(device=dev, dtype=torch.int64).contiguous()
    order, offsets, counts = _runs(assign, k)
    sum_sqrt = torch.zeros(k, dtype=torch.float32, device=dev)
    _cluster_update(None, order, offsets, k, dist=dist, sum_sqrt=sum_sqrt)
    cnt = counts.double()
    dens = torch.where(cnt > 1, sum_sqrt.double() / cnt.clamp(min=1) / torch.log(cnt + 10), torch.zeros_like(cnt))   # :80-83
    dens = torch.where(cnt <= 1, dens.max(), dens)                                                                    # :86-89
    lo, hi = torch.quantile(dens, torch.tensor([0.1, 0.9], dtype=torch.float64, device=dev))                          # np.percentile: linear
    dens = dens.clamp(lo, hi)                                                                                         # :91-92
    return (float(temperature) * dens / dens.mean()).float()                                                          # :93-94, :101


def extract_all_emb(loader, model, device):
    """src/contrastor/utils.py:11-25: anchor then positive embeddings of every batch, stacked (kept on the device)."""
    out = []
    with torch.no_grad():
        for _, anchor, positive in loader:
            anchor, positive = model.bert_extract(anchor, positive, device)
            out.append(model.seq2vec(anchor.to(device)).float())
            out.append(model.seq2vec(positive.to(device)).float())
    return torch.cat(out)


def run_kmeans(proto_nce_config, x_or_loader, model=None, device=None):
    """``run_kmeans(proto_nce_config, loader, model, device)`` of src/contrastor/utils.py:50-105; ``x_or_loader`` may
    also be the embedding matrix itself (``model=None``).  Same config keys (``cluster.num_cluster``, ``niter``,
    ``nredo``, ``verbose``, ``max_points_per_centroid``, ``min_points_per_centroid``, ``temperature``), same result:
    per cluster count ``emb2cluster`` int64 [n], ``centroids`` float32 [k, d] L2-normalised, ``density`` float32 [k]."""
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    x = extract_all_emb(x_or_loader, model, dev) if model is not None else torch.as_tensor(x_or_loader)
    x = x.to(device=dev, dtype=torch.float32).contiguous()
    cfg = proto_nce_config["cluster"]
    results = {"emb2cluster": [], "centroids": [], "density": []}
    for seed, num_cluster in enumerate(cfg["num_cluster"]):                       # :57-58: seed = position in the list
        d, k = x.shape[1], int(num_cluster)
        clus = Clustering(d, k)                                                   # get_cluster, :28-36
        clus.verbose = cfg.get("verbose", False)
        clus.niter = cfg.get("niter", clus.niter)
        clus.nredo = cfg.get("nredo", clus.nredo)
        clus.seed = seed
        clus.max_points_per_centroid = cfg.get("max_points_per_centroid", clus.max_points_per_centroid)
        clus.min_points_per_centroid = cfg.get("min_points_per_centroid", clus.min_points_per_centroid)
        index = FlatL2Index(d, device=dev)                                        # get_clus_idx, :39-47
        clus.train(x, index)                                                      # :64
        dist, assign = flat_l2_search(x, clus.centroids_tensor, 1)                # :67, kept on the device
        results["emb2cluster"].append(assign[:, 0].contiguous())                  # :68, :100
        results["density"].append(cluster_density(dist, assign, k, proto_nce_config["temperature"]))   # :73-94, :101
        results["centroids"].append(torch.nn.functional.normalize(clus.centroids_tensor, p=2, dim=1))  # :97-98
    return results
