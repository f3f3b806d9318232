"""The oracle against the golden vectors produced by the reference itself
(tests/golden/make_golden.py).  CPU only."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import dense_topk, infonce, pairs

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _t(a):
    return torch.from_numpy(np.asarray(a))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "infonce_*.npz"))))
def test_nce_info_loss_matches_reference(path):
    z = np.load(path)
    queue = _t(z["queue"]) if "queue" in z.files else None
    loss, dq, dk = infonce.nce_info_loss(_t(z["q"]), _t(z["k"]), queue, float(z["temperature"]))
    # same fp32 arithmetic, different association: 2e-5 relative on the loss, 1e-4 abs on grads / T
    assert abs(float(loss) - float(z["loss"])) <= 2e-5 * abs(float(z["loss"])) + 1e-5
    np.testing.assert_allclose(dq.numpy(), z["dq"], rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(dk.numpy(), z["dk"], rtol=2e-4, atol=2e-4)
    # and tightly in float64, which removes the summation-order noise of the restatement
    loss64, dq64, dk64 = infonce.nce_info_loss(_t(z["q"]), _t(z["k"]), queue, float(z["temperature"]),
                                               dtype=torch.float64)
    assert abs(float(loss64) - float(z["loss"])) <= 2e-5 * abs(float(z["loss"])) + 1e-5
    np.testing.assert_allclose(dq64.numpy(), z["dq"], rtol=2e-4, atol=2e-4)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "moco_*.npz"))))
def test_moco_infonce_matches_reference(path):
    z = np.load(path)
    loss, dq, dk = infonce.moco_infonce(_t(z["q"]), _t(z["k"]), _t(z["queue"]), float(z["temperature"]))
    assert abs(float(loss) - float(z["loss"])) <= 2e-5 * abs(float(z["loss"])) + 1e-6
    np.testing.assert_allclose(dq.numpy(), z["dq"], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(dk.numpy(), z["dk"], rtol=2e-4, atol=2e-5)


def _proto_inputs(z):
    """the selection contrastive_loss.py:101-112,122-123 makes, with the fixed sampler of
    make_golden.py (sorted(set)[:r]); tests/golden/proto_inputs.py"""
    import sys
    sys.path.insert(0, GOLDEN)
    from proto_inputs import selected_from_fixture
    protos, temps = selected_from_fixture(z)
    return [_t(p) for p in protos], [_t(t) for t in temps]


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "proto_*.npz"))))
def test_proto_loss_matches_reference(path):
    z = np.load(path)
    protos, temps = _proto_inputs(z)
    loss, dq = infonce.proto_loss(_t(z["q"]), protos, temps)
    assert abs(float(loss) - float(z["loss"])) <= 2e-5 * abs(float(z["loss"])) + 1e-5
    np.testing.assert_allclose(dq.numpy(), z["dq"], rtol=2e-4, atol=2e-4)


def test_select_matches_closest_docs():
    """tfidf_doc_ranker.py:60-75 run by the reference on a synthetic CSR with many ties:
    the returned SCORES must agree exactly (tie order among equal scores is unspecified
    upstream, so ids are compared as score-consistent sets)."""
    z = np.load(os.path.join(GOLDEN, "closest_docs.npz"))
    doc_mat = z["doc_mat"]
    for qv, k, ids, scs in zip(z["queries"], z["k"], z["ids"], z["scores"]):
        full = qv @ doc_mat                                    # spvec * doc_mat, dense
        nnz = np.flatnonzero(full != 0)
        n_ret = int((ids >= 0).sum())
        assert n_ret == min(int(k), len(nnz)) or n_ret == len(nnz)   # :67-68 fewer-than-k rule
        ref_ids, ref_scores = ids[:n_ret], scs[:n_ret]
        # oracle select over the non-zero entries (closest_docs only ranks res.data)
        v, i = dense_topk.select_topk_desc(torch.from_numpy(full[nnz][None, :]).float(), n_ret)
        got_ids = nnz[i[0].numpy()]
        np.testing.assert_allclose(full[got_ids], ref_scores, rtol=0, atol=1e-12)
        np.testing.assert_allclose(full[ref_ids], ref_scores, rtol=0, atol=1e-12)
        # descending, and ties in index order on our side
        assert np.all(np.diff(full[got_ids]) <= 0)
        same = np.diff(full[got_ids]) == 0
        assert np.all(np.diff(got_ids)[same] > 0)
        # numpy restatement of the select agrees too
        np.testing.assert_array_equal(nnz[dense_topk.closest_docs_select(full[nnz], n_ret)], got_ids)


def test_pairs_match_reference_loop():
    z = np.load(os.path.join(GOLDEN, "pairs.npz"))
    for d in range(int(z["ndocs"])):
        got = pairs.doc_sentence_pairs(z[f"x{d}"])
        ref_pairs, ref_scores = z[f"pairs{d}"], z[f"scores{d}"]
        assert len(got) == len(ref_scores)
        np.testing.assert_allclose([g[1] for g in got], ref_scores, rtol=0, atol=1e-12)
        # where the reference's scores are strictly separated the pair order must be identical
        sep = np.ones(len(ref_scores), dtype=bool)
        if len(ref_scores) > 1:
            gap = np.abs(np.diff(ref_scores)) > 1e-12
            sep[1:] &= gap
            sep[:-1] &= gap
        got_pairs = np.array([[g[0][0], g[0][1]] for g in got], dtype=np.int64).reshape(-1, 2)
        np.testing.assert_array_equal(got_pairs[sep], ref_pairs.reshape(-1, 2)[sep])


def test_topk_tie_rule_and_chunking():
    g = torch.Generator().manual_seed(1337)
    c = torch.randn(5000, 32, generator=g)
    c[100] = c[7]
    c[4100] = c[7]                                             # exact duplicates -> exact ties
    q = torch.randn(17, 32, generator=g)
    q[0] = c[7]
    v_full, i_full = dense_topk.select_topk_desc(dense_topk.scores_fp32(q, c), 10)
    v_sort, i_sort = torch.sort(dense_topk.scores_fp32(q, c), dim=1, descending=True, stable=True)
    assert torch.equal(i_full, i_sort[:, :10]) and torch.equal(v_full, v_sort[:, :10])
    assert i_full[0, :3].tolist() == [7, 100, 4100]
    for chunk in (64, 999, 5000, 100000):
        v, i = dense_topk.search(q, c, 10, chunk_rows=chunk)
        assert torch.equal(i, i_full)
        torch.testing.assert_close(v, v_full, rtol=1e-6, atol=1e-6)
    for ws in (1, 2, 3, 8):
        v, i = dense_topk.sharded_search(q, c, 10, ws)
        assert torch.equal(i, i_full)
    v, i = dense_topk.search(q, c[:4], 10)                     # k > Nc -> clamped
    assert i.shape == (17, 4)
    vf, jf = dense_topk.search_fast(q, c, 10)
    torch.testing.assert_close(vf, v_full, rtol=1e-6, atol=1e-6)


def test_paired_scores_match_the_reference_expression(golden_dir):
    """src/evaluation.py:112 evaluated literally by the generator -> oracle restatement."""
    z = np.load(os.path.join(golden_dir, "paired.npz"))
    got = dense_topk.paired_scores(torch.from_numpy(z["clm"]), torch.from_numpy(z["evdn"]))
    np.testing.assert_allclose(got.numpy(), z["per_pair"], rtol=0, atol=1e-7)
    assert abs(got.mean().item() - float(z["mean"])) < 1e-7
