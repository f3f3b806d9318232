"""Seeded ProtoNCE inputs shared by make_golden.py (which runs the reference on them) and the GPU test (which
regenerates them instead of reading megabytes of centroids from a fixture).  torch's CPU generator is
deterministic across machines for a given torch build; a stored checksum guards the regeneration."""
import numpy as np
import torch


def _unit(x, dim=1):
    return torch.nn.functional.normalize(x, dim=dim)


def make_inputs(n, d, num_cluster, seed=1337):
    """q [n, d] (unit rows), index [n], cluster_result = {'emb2cluster', 'centroids', 'density'} per set -- the
    structure run_kmeans returns (src/contrastor/utils.py:50-105) and _compute_proto_loss consumes
    (src/contrastor/contrastive_loss.py:95-135)."""
    g = torch.Generator().manual_seed(seed)
    q = _unit(torch.randn(n, d, generator=g))
    index = torch.randperm(4 * n, generator=g)[:n]
    cluster_result = {"emb2cluster": [], "centroids": [], "density": []}
    for c in num_cluster:
        e2c = torch.randint(0, c, (4 * n,), generator=g)
        e2c[0] = c - 1                      # make emb2cluster.max() == c-1 as in a real run
        cluster_result["emb2cluster"].append(e2c)
        cluster_result["centroids"].append(_unit(torch.randn(c, d, generator=g)))
        cluster_result["density"].append(torch.rand(c, generator=g) * 0.1 + 0.02)
    return q, index, cluster_result


def checksum(cluster_result):
    """float64 sums of every array, in a fixed order"""
    out = []
    for key in ("emb2cluster", "centroids", "density"):
        for t in cluster_result[key]:
            out.append(float(t.double().sum()))
    return np.array(out, dtype=np.float64)


def selected_from_fixture(z):
    """(protos, temps) per cluster set as numpy arrays: the selection of contrastive_loss.py:101-112,:122-123 with
    make_golden.py's fixed sampler (sorted(set)[:r] in place of random.sample at :109).  Compact fixtures carry a
    seed instead of the cluster sets: those are regenerated and checked against the stored checksum."""
    index, r = z["index"], int(z["num_neg_proto"])
    if "seed" in z.files:
        n, d = z["q"].shape
        _, index_t, cr = make_inputs(n, d, [int(c) for c in z["num_cluster"]], seed=int(z["seed"]))
        if not (np.array_equal(index_t.numpy(), index) and np.array_equal(checksum(cr), z["checksum"])):
            raise AssertionError("regenerated ProtoNCE inputs do not match the fixture's checksum (different torch RNG?)")
        sets = [(cr["emb2cluster"][s].numpy(), cr["centroids"][s].numpy(), cr["density"][s].numpy())
                for s in range(int(z["num_sets"]))]
    else:
        sets = [(z[f"emb2cluster{s}"], z[f"centroids{s}"], z[f"density{s}"]) for s in range(int(z["num_sets"]))]
    protos, temps = [], []
    for e2c, cen, den in sets:
        pos_id = e2c[index]
        neg = sorted(set(range(int(e2c.max()))) - set(pos_id.tolist()))[:r]
        ids = np.concatenate([pos_id, np.array(neg, dtype=np.int64)])
        protos.append(cen[ids])
        temps.append(den[ids])
    return protos, temps
