"""The clustering oracle (oracle/kmeans.py) against the fixture produced by RUNNING the reference's run_kmeans
(tests/golden/make_golden.py::gen_kmeans; src/contrastor/utils.py:50-105).  CPU only."""
import os

import numpy as np

from oracle import kmeans as okm

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kmeans_density.npz")


def test_density_restatement_matches_the_reference_run():
    z = np.load(GOLDEN)
    x, temp = z["x"], float(z["temperature"])
    for s, k in enumerate(z["num_cluster"]):
        raw = z[f"raw_centroids_{s}"]
        d, i = okm.assign(x, raw)                                             # utils.py:67
        np.testing.assert_array_equal(i, z[f"emb2cluster_{s}"])               # :68, :100
        dens = okm.density(d[:, None], i[:, None], int(k), temp)              # :73-94
        np.testing.assert_allclose(dens.astype(np.float32), z[f"density_{s}"], rtol=1e-6, atol=0)
        unit = raw / np.maximum(np.linalg.norm(raw, axis=1, keepdims=True), 1e-12)
        np.testing.assert_allclose(unit, z[f"centroids_{s}"], rtol=1e-6, atol=1e-7)   # :97-98
        assert abs(dens.mean() - temp) < 1e-9                                 # :93-94: rescaled to the temperature


def test_lloyd_objective_never_increases_and_reproduces_the_fixture_centroids():
    z = np.load(GOLDEN)
    x, niter = z["x"], int(z["niter"])
    for s, k in enumerate(z["num_cluster"]):
        c, obj = okm.lloyd(x, okm.init_centroids(x, int(k), s), niter)
        np.testing.assert_array_equal(c, z[f"raw_centroids_{s}"])
        # (an empty-cluster split may raise the objective by a hair; none happens on this fixture)
        assert all(b <= a * (1 + 1e-9) for a, b in zip(obj, obj[1:]))


def test_empty_cluster_is_reseeded_next_to_the_largest():
    x = np.array([[0.0, 0.0], [0.1, 0.0], [0.0, 0.1], [5.0, 5.0]], np.float32)
    cent = np.array([[0.0, 0.0], [5.0, 5.0], [100.0, 100.0]], np.float32)     # the third attracts nothing
    d, i = okm.assign(x, cent)
    new, nsplit = okm.update(x, i, cent)
    assert nsplit == 1
    base = x[:3].astype(np.float64).mean(0).astype(np.float32)
    np.testing.assert_allclose(new[2], base * np.array([1 + okm.SPLIT_EPS, 1 - okm.SPLIT_EPS], np.float32))
    np.testing.assert_allclose(new[0], base * np.array([1 - okm.SPLIT_EPS, 1 + okm.SPLIT_EPS], np.float32))
    np.testing.assert_array_equal(new[1], x[3])
