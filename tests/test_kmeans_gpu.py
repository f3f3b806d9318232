"""Prototype clustering on the GPU (clustering.py, csrc/kmeans.cuh) against the reference-produced fixture and the
oracle's Lloyd iteration.  GPU only."""
import os

import numpy as np
import pytest
import torch

import drs_b200
from oracle import kmeans as okm

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kmeans_density.npz")
DEV = "cuda:0"


def test_assignment_and_density_match_the_reference_run():
    """utils.py:67-101 on the fixture's trained centroids: the assignment (flat-L2 search), the concentration estimate
    and the normalised centroids the reference's run_kmeans returned."""
    z = np.load(GOLDEN)
    x = torch.from_numpy(z["x"]).to(DEV)
    temp = float(z["temperature"])
    for s, k in enumerate(z["num_cluster"]):
        raw = torch.from_numpy(z[f"raw_centroids_{s}"]).to(DEV)
        index = drs_b200.FlatL2Index(x.shape[1], device=DEV)
        index.add(raw)
        D, I = index.search(x, 1)                                             # numpy, like faiss
        np.testing.assert_array_equal(I[:, 0], z[f"emb2cluster_{s}"])
        dens = drs_b200.cluster_density(D, I, int(k), temp)
        assert dens.dtype == torch.float32 and dens.is_cuda
        np.testing.assert_allclose(dens.cpu().numpy(), z[f"density_{s}"], rtol=2e-5, atol=0)
        # and from the oracle's exact distances: isolates the estimate from the search's 4e-6 relative error
        d0, i0 = okm.assign(z["x"], z[f"raw_centroids_{s}"])
        dens0 = drs_b200.cluster_density(d0[:, None], i0[:, None], int(k), temp)
        np.testing.assert_allclose(dens0.cpu().numpy(), z[f"density_{s}"], rtol=2e-6, atol=0)


@pytest.mark.parametrize("n,dim,k", [(600, 32, 16), (5000, 128, 64), (3000, 100, 7)])
def test_centroid_update_is_the_ordered_float64_mean(n, dim, k):
    g = torch.Generator().manual_seed(n)
    x = torch.randn(n, dim, generator=g)
    assign = torch.randint(0, k, (n,), generator=g)
    assign[assign == 3] = 2                                                    # cluster 3 is empty
    cent = torch.randn(k, dim, generator=g)
    ref, nsplit = okm.update(x.numpy(), assign.numpy(), cent.numpy())
    got = cent.clone().to(DEV)
    assert drs_b200.update_centroids(x.to(DEV), assign.to(DEV), got) == nsplit == 1
    np.testing.assert_array_equal(got.cpu().numpy(), ref)                      # same sums in the same order: bit-equal


def test_training_follows_the_oracle_lloyd_iteration_from_the_same_start():
    z = np.load(GOLDEN)
    x = z["x"]
    for s, k in enumerate(z["num_cluster"]):
        clus = drs_b200.Clustering(x.shape[1], int(k))
        clus.niter = int(z["niter"])
        clus.centroids = okm.init_centroids(x, int(k), s).reshape(-1)          # faiss: preset centroids are the start
        index = drs_b200.FlatL2Index(x.shape[1], device=DEV)
        clus.train(x, index)
        got = drs_b200.vector_to_array(clus.centroids).reshape(int(k), -1)     # utils.py:71
        np.testing.assert_allclose(got, z[f"raw_centroids_{s}"], rtol=1e-5, atol=1e-6)
        _, obj = okm.lloyd(x, okm.init_centroids(x, int(k), s), clus.niter)
        np.testing.assert_allclose(clus.objective, obj, rtol=1e-5)
        assert index.ntotal == int(k)                                          # the trained centroids are left in the index (:67)


def test_run_kmeans_result_feeds_the_proto_loss():
    g = torch.Generator().manual_seed(5)
    centers = torch.nn.functional.normalize(torch.randn(40, 128, generator=g), dim=1)
    lab = torch.randint(0, 40, (4000,), generator=g)
    x = torch.nn.functional.normalize(centers[lab] + 0.1 * torch.randn(4000, 128, generator=g), dim=1)
    cfg = {"temperature": 0.05, "cluster": {"num_cluster": [64, 96], "num_neg_proto": 8, "verbose": False, "niter": 8, "nredo": 2,
                                            "max_points_per_centroid": 1000, "min_points_per_centroid": 1}}
    res = drs_b200.run_kmeans(cfg, x, device=DEV)
    res2 = drs_b200.run_kmeans(cfg, x, device=DEV)
    for s, k in enumerate(cfg["cluster"]["num_cluster"]):
        e2c, cen, den = res["emb2cluster"][s], res["centroids"][s], res["density"][s]
        assert e2c.shape == (4000,) and e2c.dtype == torch.int64 and int(e2c.min()) >= 0 and int(e2c.max()) < k
        assert cen.shape == (k, 128) and cen.dtype == torch.float32 and den.shape == (k,) and den.dtype == torch.float32
        torch.testing.assert_close(cen.norm(dim=1), torch.ones(k, device=DEV), rtol=1e-5, atol=1e-5)
        assert abs(float(den.double().mean()) - 0.05) < 1e-6 and float(den.min()) > 0
        assert torch.equal(e2c, res2["emb2cluster"][s]) and torch.equal(cen, res2["centroids"][s]) and torch.equal(den, res2["density"][s])
        # every sample sits with its nearest returned prototype direction
        d_ref, i_ref = okm.assign(x.numpy(), drs_b200.vector_to_array(cen).reshape(k, 128))
        assert (i_ref == e2c.cpu().numpy()).mean() > 0.99                       # (normalising the centroids may flip a borderline sample)
    crit = drs_b200.NCELoss({"temperature": 0.05, "cluster": cfg["cluster"]})
    q = x[:128].to(DEV).requires_grad_(True)
    k_ = x[128:256].to(DEV)
    loss = crit(q, k_, None, res, torch.arange(128, device=DEV))
    loss.backward()
    assert torch.isfinite(loss) and torch.isfinite(q.grad).all()


def test_fewer_points_than_clusters_is_an_error():
    clus = drs_b200.Clustering(16, 50)
    with pytest.raises(RuntimeError, match="at least as large as number of clusters"):
        clus.train(torch.randn(20, 16), drs_b200.FlatL2Index(16, device=DEV))


def test_bad_arguments_are_reported():
    x = torch.randn(10, 8, device=DEV)
    cent = torch.randn(3, 8, device=DEV)
    with pytest.raises(ValueError, match="cluster ids must lie"):
        drs_b200.update_centroids(x, torch.full((10,), 3, device=DEV), cent)
    with pytest.raises(ValueError, match="shape mismatch"):
        drs_b200.update_centroids(x, torch.zeros(9, dtype=torch.int64, device=DEV), cent)
    with pytest.raises(RuntimeError, match="no CPU path"):
        drs_b200.update_centroids(x.cpu(), torch.zeros(10, dtype=torch.int64), cent.cpu())
