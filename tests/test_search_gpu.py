"""Parity of the CUDA search path (through the C ABI) against the CPU oracle.  GPU only."""
import os

import numpy as np
import pytest
import torch

import drs_b200
from oracle import dense_topk

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEV = "cuda:0"


def _unit(x):
    return torch.nn.functional.normalize(x, dim=1)


def _data(nq, nc, dim, dtype, planted=True, seed=1337):
    g = torch.Generator(device=DEV).manual_seed(seed)               # main.py:45 default seed
    c = _unit(torch.randn(nc, dim, generator=g, device=DEV))
    if planted:
        j = torch.randint(0, nc, (nq,), generator=g, device=DEV)
        q = _unit(c[j] + 0.1 * torch.randn(nq, dim, generator=g, device=DEV))
    else:
        q = _unit(torch.randn(nq, dim, generator=g, device=DEV))
    return q.to(dtype), c.to(dtype)


def _check(q, c, k, s, i, score_rtol, gap):
    """scores within tolerance; ids bit-exact wherever the oracle's score gap exceeds `gap`
    (north star: 'top-k ids bit-exact wherever the score gap exceeds the tolerance')."""
    rv, ri = dense_topk.search(q.cpu(), c.cpu(), k)
    s, i = s.cpu(), i.cpu()
    assert s.shape == rv.shape and i.shape == ri.shape and i.dtype == torch.int64 and s.dtype == torch.float32
    torch.testing.assert_close(s, rv, rtol=score_rtol, atol=score_rtol * 1e-1)
    assert torch.all(s[:, :-1] >= s[:, 1:]), "scores must be sorted descending"
    kk = rv.shape[1]
    v2, _ = dense_topk.search(q.cpu(), c.cpu(), kk + 1)          # one more, to know the gap below the k-th
    if v2.shape[1] == kk + 1:
        nxt = v2[:, 1:]
    else:
        nxt = torch.cat([rv[:, 1:], torch.full((rv.shape[0], 1), -1e30)], dim=1)
    prev = torch.cat([torch.full((rv.shape[0], 1), 1e30), rv[:, :-1]], dim=1)
    strict = ((rv - nxt) > gap) & ((prev - rv) > gap)
    assert torch.equal(i[strict], ri[strict]), f"{(i[strict] != ri[strict]).sum().item()} ids differ outside ties"
    return ri


@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("nq,nc,dim,k", [(128, 256, 64, 5), (300, 20000, 768, 10), (77, 12345, 200, 10),
                                         (1, 1000, 128, 1), (513, 70000, 128, 16), (64, 9000, 768, 32)])
def test_bf16_tcgen05_parity(cg, nq, nc, dim, k):
    q, c = _data(nq, nc, dim, torch.bfloat16, planted=(nq % 2 == 0))
    drs_b200.set_option("search.cta_group", cg)
    try:
        s, i = drs_b200.search(q, c, k)
    finally:
        drs_b200.set_option("search.cta_group", 0)
    # bf16 path tolerance from the north star: 2e-2 relative.  The oracle consumes the SAME bf16
    # values upcast to fp32, so the observed error is fp32 summation order only.
    _check(q, c, k, s, i, score_rtol=2e-2, gap=1e-4)


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("nq,nc,dim,k", [(100, 5000, 64, 5), (77, 1234, 100, 10), (1000, 100000, 768, 5),
                                         (3, 40, 7, 4), (130, 3000, 33, 32), (700, 60000, 768, 100), (300, 20000, 1024, 16)])
def test_fp32_exact_parity(nq, nc, dim, k, mode):
    """BASELINE config 0 (1k claims x 100k x 768 fp32, top-5) and ragged shapes; 1e-5 relative (north-star fp32 bar).
    mode 0: fp32 operands on tcgen05 as a 3 x TF32 split (dims that are not a multiple of 4 take the FFMA kernel by
    themselves); mode 1: the FFMA kernel, kept as the in-tree checker."""
    q, c = _data(nq, nc, dim, torch.float32, planted=(dim != 7))
    drs_b200.set_option("search.fp32_mode", mode)
    try:
        s, i = drs_b200.search(q, c, k)
    finally:
        drs_b200.set_option("search.fp32_mode", 0)
    _check(q, c, k, s, i, score_rtol=1e-5, gap=5e-6)


@pytest.mark.parametrize("dtype,dim", [(torch.bfloat16, 2048), (torch.bfloat16, 4096), (torch.float32, 2048), (torch.bfloat16, 8),
                                       (torch.float32, 4)])
def test_extreme_embedding_widths(dtype, dim):
    """Very wide (64 K blocks per tile) and very narrow rows on both tensor-core paths."""
    q, c = _data(200, 30000, dim, dtype, planted=dim >= 64)
    s, i = drs_b200.search(q, c, 10)
    tol = (1e-5, 5e-6) if dtype == torch.float32 else (2e-2, 1e-4)
    _check(q, c, 10, s, i, score_rtol=tol[0], gap=tol[1])


def test_fp32_tensor_core_path_agrees_with_the_ffma_checker():
    """The two fp32 arithmetic paths on the same data: scores within 4e-6 relative of each other (measured worst case
    of the 3 x TF32 split against float64: 3.8e-6), ids identical wherever the FFMA path's gap exceeds 1e-5."""
    q, c = _data(512, 200000, 768, torch.float32, planted=True)
    q[:256] = _unit(torch.randn(256, 768, device=DEV, generator=torch.Generator(device=DEV).manual_seed(9)))
    s0, i0 = drs_b200.search(q, c, 10)
    drs_b200.set_option("search.fp32_mode", 1)
    try:
        s1, i1 = drs_b200.search(q, c, 10)
    finally:
        drs_b200.set_option("search.fp32_mode", 0)
    assert ((s0 - s1).abs() <= 4e-6 * s1.abs() + 1e-6).all()
    gap = torch.ones_like(s1, dtype=torch.bool)
    d = s1[:, :-1] - s1[:, 1:]
    gap[:, 1:] &= d > 1e-5
    gap[:, :-1] &= d > 1e-5
    assert torch.equal(i0[gap], i1[gap])


@pytest.mark.parametrize("dtype,nq,nc,dim,k", [(torch.bfloat16, 300, 50000, 768, 100), (torch.bfloat16, 65, 9000, 128, 33),
                                               (torch.float32, 50, 20000, 64, 100), (torch.bfloat16, 40, 700, 64, 256),
                                               (torch.bfloat16, 16, 90, 64, 64)])
def test_large_k_multi_pass_is_exact(dtype, nq, nc, dim, k):
    """BASELINE config 4's top-100: k > 32 is served by adaptive passes (picks are emitted only while
    no split's full 32-entry list can hide a better row; open claims are rescanned below their last
    pick); must equal the oracle like a single pass (incl. k > Nc)."""
    q, c = _data(nq, nc, dim, dtype, planted=True)
    c[nc // 2] = c[5]
    c[nc - 1] = c[5]                                              # ties that straddle pass boundaries
    q[0] = c[5]
    s, i = drs_b200.search(q, c, k)
    tol = (1e-5, 5e-6) if dtype == torch.float32 else (2e-2, 1e-4)
    ri = _check(q, c, k, s, i, score_rtol=tol[0], gap=tol[1])
    assert i.shape[1] == min(k, nc)
    assert i[0, :3].cpu().tolist() == [5, nc // 2, nc - 1]
    assert all(len(set(r)) == i.shape[1] for r in i.cpu().tolist())


def test_large_k_adaptive_passes_one_pass_when_spread_more_when_clustered():
    """The pass count adapts to the data and the result stays exact either way.
    Spread neighbours (random corpus): the first pass closes every claim.  Clustered neighbours (a
    wiki page's sentences are adjacent rows): 90 near-duplicates of the claim inside ONE corpus split
    overflow its 16-entry list, the claim stays open and later passes finish it."""
    from importlib import import_module
    retrieval = import_module(drs_b200.__name__ + ".retrieval")
    nq, nc, dim, k = 130, 60000, 128, 100
    q, c = _data(nq, nc, dim, torch.bfloat16, planted=True)
    s, i = drs_b200.search(q, c, k)
    _check(q, c, k, s, i, score_rtol=2e-2, gap=1e-4)
    assert retrieval.open_claims_per_pass()[:4] == [0, 0, 0, 0]
    g = torch.Generator(device=DEV).manual_seed(7)
    cf = c.float()
    base = 1000                                                     # rows 1000..1089: inside one 256-row tile
    cf[base:base + 90] = _unit(q[3].float()[None, :] + 0.02 * torch.randn(90, dim, generator=g, device=DEV))
    c2 = cf.to(torch.bfloat16)
    s, i = drs_b200.search(q, c2, k)
    ri = _check(q, c2, k, s, i, score_rtol=2e-2, gap=1e-4)
    opened = retrieval.open_claims_per_pass()
    assert opened[0] >= 1 and opened[6] == 0 and opened[7] == 0, opened          # 7 = ceil(100 / 16) passes at most
    assert set(i[3, :90].cpu().tolist()) == set(range(base, base + 90)) == set(ri[3, :90].tolist())
    d, li = drs_b200.flat_l2_search(q.float(), c2.float(), k)        # same machinery behind the L2 epilogue
    rd, rli = dense_topk.flat_l2_search(q.float().cpu(), c2.float().cpu(), k)
    torch.testing.assert_close(d.cpu(), rd, rtol=1e-4, atol=1e-4)
    assert set(li[3, :90].cpu().tolist()) == set(range(base, base + 90))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_ties_break_to_lower_index(dtype):
    """Exact duplicates of a corpus row: equal scores, ids must come out in ascending order,
    across tiles, CTAs and splits."""
    g = torch.Generator(device=DEV).manual_seed(7)
    c = _unit(torch.randn(30000, 64, generator=g, device=DEV)).to(dtype)
    dup = [5, 300, 4097, 12288, 29999]
    for d in dup[1:]:
        c[d] = c[dup[0]]
    q = c[dup[0]].float().repeat(4, 1).to(dtype)
    s, i = drs_b200.search(q, c, 8)
    assert i[:, :5].cpu().tolist() == [dup] * 4
    assert torch.all(s[:, :5] == s[:, :1])
    rv, ri = dense_topk.search(q.cpu(), c.cpu(), 8)
    assert torch.equal(i[:, :5].cpu(), ri[:, :5])


def test_k_larger_than_corpus_and_tiny_corpus():
    q, c = _data(9, 6, 64, torch.bfloat16, planted=False)
    s, i = drs_b200.search(q, c, 10)                              # closest_docs :67-68 -> all 6 rows
    assert s.shape == (9, 6)
    rv, ri = dense_topk.search(q.cpu(), c.cpu(), 10)
    assert torch.equal(i.cpu(), ri)
    assert sorted(i[0].cpu().tolist()) == list(range(6))


def test_id_base_and_sharded_merge_equals_single_search():
    """SURVEY 8(e) on one GPU: shard the corpus, search each shard with its global offset, merge
    with the engine's merge kernel -> identical to the unsharded search and to the oracle."""
    q, c = _data(200, 50001, 128, torch.bfloat16)
    s_full, i_full = drs_b200.search(q, c, 10)
    for world in (2, 3, 8):
        ss, ii = [], []
        for r in range(world):
            lo, hi = drs_b200.shard_bounds(c.shape[0], r, world)
            s, i = drs_b200.search(q, c[lo:hi], 10, id_base=lo)
            ss.append(s)
            ii.append(i)
        ms, mi = drs_b200.merge_shards(torch.stack(ss), torch.stack(ii))
        assert torch.equal(mi, i_full) and torch.equal(ms, s_full)
    rv, ri = dense_topk.sharded_search(q.cpu(), c.cpu(), 10, 4)
    assert torch.equal(i_full.cpu()[:, 0], ri[:, 0])


def test_merge_shards_with_empty_slots_and_cross_shard_ties():
    s = torch.tensor([[[0.9, 0.5, float("-inf")]], [[0.9, 0.9, 0.1]]], device=DEV)      # [2 shards, 1 query, 3]
    i = torch.tensor([[[10, 11, -1]], [[3, 40, 41]]], device=DEV)
    ms, mi = drs_b200.merge_shards(s, i)
    assert mi.cpu().tolist() == [[3, 10, 40]]
    assert ms.cpu().tolist() == [[pytest.approx(0.9)] * 3]
    s = torch.full((2, 1, 3), float("-inf"), device=DEV)
    i = torch.full((2, 1, 3), -1, dtype=torch.int64, device=DEV)
    s[1, 0, 0], i[1, 0, 0] = 0.25, 77
    ms, mi = drs_b200.merge_shards(s, i)
    assert mi.cpu().tolist() == [[77, -1, -1]]


def test_dense_index_mirrors_closest_docs():
    """TfidfDocRanker surface (tfidf_doc_ranker.py:52-84): (list of doc ids, f64 ndarray of scores)."""
    q, c = _data(5, 3000, 64, torch.float32)
    names = [f"doc_{n}" for n in range(c.shape[0])]
    index = drs_b200.DenseIndex(c, names, dtype=torch.float32)
    ids, scores = index.closest_docs(q[0].cpu(), k=5)              # host query vector is copied in
    rv, ri = dense_topk.search(q[:1].cpu(), c.cpu(), 5)
    assert ids == [names[j] for j in ri[0].tolist()]
    assert isinstance(scores, np.ndarray) and scores.dtype == np.float64
    np.testing.assert_allclose(scores, rv[0].numpy(), rtol=1e-5)
    batch = index.batch_closest_docs(q, k=3, num_workers=4)
    assert len(batch) == 5 and batch[0][0] == ids[:3]
    assert index.get_doc_index("doc_17") == 17 and index.get_doc_id(17) == "doc_17"


def test_closest_docs_golden_through_the_engine():
    """The reference's own closest_docs outputs (golden, dyadic values so fp32 is exact):
    dense restatement = query row x doc_mat columns, fp32 exact path."""
    z = np.load(os.path.join(GOLDEN, "closest_docs.npz"))
    corpus = torch.from_numpy(z["doc_mat"].T.copy()).float().to(DEV)          # [docs, hash]
    for qv, k, ids, scs in zip(z["queries"], z["k"], z["ids"], z["scores"]):
        n_ret = int((ids >= 0).sum())
        kk = min(int(k), 32, n_ret)
        s, i = drs_b200.search(torch.from_numpy(qv[None, :]).float().to(DEV), corpus, kk)
        np.testing.assert_array_equal(s[0].cpu().numpy().astype(np.float64), scs[:kk])   # scores identical
        full = qv @ z["doc_mat"]
        np.testing.assert_array_equal(full[i[0].cpu().numpy()], scs[:kk])                 # ids carry those scores
        same = np.diff(scs[:kk]) == 0
        assert np.all(np.diff(i[0].cpu().numpy())[same] > 0)                              # ties -> lower index first


def test_unsupported_arguments_raise_runtime_error():
    q, c = _data(4, 100, 64, torch.bfloat16)
    with pytest.raises(RuntimeError, match="exceeds the engine limit"):
        drs_b200.search(q[:, :64].repeat(1, 1), torch.cat([c] * 4), 300)
    with pytest.raises(RuntimeError, match="multiple of 8"):
        drs_b200.search(q[:, :60].contiguous(), c[:, :60].contiguous(), 3)


@pytest.mark.parametrize("nq", [16, 128])
def test_small_batch_bandwidth_regime_parity(nq):
    """BASELINE rows 2b/3b: few claims per corpus pass (HBM-bound regime), 1M rows."""
    q, c = _data(nq, 1_000_000, 768, torch.bfloat16)
    s, i = drs_b200.search(q, c, 10)
    _check(q, c, 10, s, i, score_rtol=2e-2, gap=1e-4)


@pytest.mark.parametrize("nq,nc,k", [(10000, 5_400_000, 10),      # BASELINE configs[1]
                                      (10000, 25_000_000, 10),     # configs[2], the metric's config (38.4 GB resident)
                                      (65536, 675_000, 100)])      # configs[4]: one GPU's share of 5.4M docs / 8
def test_full_size_properties(nq, nc, k):
    """BASELINE configs at full size (x 768 bf16): size-independent properties instead of an oracle pass --
    planted neighbours are found first, rows are sorted, ids are unique and in range, returned scores match a
    recomputation of those exact pairs, a slice agrees with the oracle, the search is idempotent, and sharding + merge
    gives the single-pass result bit for bit."""
    dim = 768
    g = torch.Generator(device=DEV).manual_seed(1337)
    c = torch.empty(nc, dim, dtype=torch.bfloat16, device=DEV)
    for r0 in range(0, nc, 1 << 20):
        r1 = min(nc, r0 + (1 << 20))
        c[r0:r1] = _unit(torch.randn(r1 - r0, dim, generator=g, device=DEV))
    planted = torch.randint(0, nc, (nq,), generator=g, device=DEV)
    q = _unit(c[planted].float() + 0.05 * torch.randn(nq, dim, generator=g, device=DEV)).bfloat16()
    s, i = drs_b200.search(q, c, k)
    assert torch.equal(i[:, 0], planted)
    assert torch.all(s[:, :-1] >= s[:, 1:])
    assert int(i.min()) >= 0 and int(i.max()) < nc
    assert all(len(set(r)) == k for r in i[:256].cpu().tolist())
    m = 1024                                                        # recompute the returned pairs of the first m claims
    rec = (q[:m].float()[:, None, :] * c[i[:m].reshape(-1)].float().reshape(m, k, dim)).sum(-1)
    torch.testing.assert_close(s[:m], rec, rtol=2e-2, atol=1e-4)
    # a claim slice against the oracle over a 200k-row window that contains each planted row
    win = c[:200000]
    sub = (planted < 200000).nonzero().flatten()[:64]
    if len(sub):
        sv, si = drs_b200.search(q[sub], win, k)
        rv, ri = dense_topk.search(q[sub].cpu(), win.cpu(), k)
        assert torch.equal(si.cpu()[:, 0], ri[:, 0])
        torch.testing.assert_close(sv.cpu(), rv, rtol=2e-2, atol=1e-4)
    s2, i2 = drs_b200.search(q, c, k)                               # idempotent
    assert torch.equal(i2, i) and torch.equal(s2, s)
    ss, ii = [], []
    for r in range(4):
        lo, hi = drs_b200.shard_bounds(nc, r, 4)
        a, b = drs_b200.search(q, c[lo:hi], k, id_base=lo)
        ss.append(a)
        ii.append(b)
    ms, mi = drs_b200.merge_shards(torch.stack(ss), torch.stack(ii))
    assert torch.equal(mi, i) and torch.equal(ms, s)


@pytest.mark.parametrize("dtype,n,nc,dim,k", [(torch.float32, 2000, 4096, 128, 1), (torch.float32, 300, 8192, 128, 5),
                                              (torch.float32, 77, 333, 50, 40), (torch.bfloat16, 1000, 6144, 128, 1)])
def test_flat_l2_search_kmeans_assignment(dtype, n, nc, dim, k):
    """src/contrastor/utils.py:64-67: `D, I = index.search(x, 1)` against 4096/6144/8192 centroids
    (config.yaml num_cluster).  Oracle = exact squared-L2 argmin in fp64 (PARITY UNPINNED vs faiss)."""
    g = torch.Generator(device=DEV).manual_seed(1337)
    cen = torch.randn(nc, dim, generator=g, device=DEV)
    lab = torch.randint(0, nc, (n,), generator=g, device=DEV)
    x = cen[lab] + 0.3 * torch.randn(n, dim, generator=g, device=DEV)
    cen, x = cen.to(dtype), x.to(dtype)
    d, i = drs_b200.flat_l2_search(x, cen, k)
    rd, ri = dense_topk.flat_l2_search(x.cpu(), cen.cpu(), k)
    assert torch.equal(i[:, 0].cpu(), lab.cpu()) and torch.equal(ri[:, 0], lab.cpu())
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    torch.testing.assert_close(d.cpu(), rd, rtol=tol, atol=tol * float(rd.max()))
    assert torch.all(d[:, 1:] >= d[:, :-1])
    gap = (rd[:, 1:] - rd[:, :-1]) if k > 1 else None
    if gap is not None and dtype == torch.float32:
        strict = torch.ones_like(ri, dtype=torch.bool)
        strict[:, 1:] &= gap > 1e-3
        strict[:, :-1] &= gap > 1e-3
        assert torch.equal(i.cpu()[strict], ri[strict])
    index = drs_b200.FlatL2Index(dim, dtype=dtype)
    index.add(cen[: nc // 2].cpu().float().numpy())
    index.add(cen[nc // 2:].cpu().float().numpy())
    assert index.ntotal == nc
    dd, ii = index.search(x[:50].cpu().float().numpy(), 1)
    assert ii.shape == (50, 1) and ii.dtype == np.int64 and [int(v) for v in ii[:, 0]] == lab[:50].cpu().tolist()


@pytest.mark.parametrize("dtype,nq,nc,dim,m,k", [(torch.bfloat16, 200, 30000, 768, 100, 15), (torch.float32, 64, 5000, 128, 100, 15),
                                                 (torch.float32, 9, 300, 50, 7, 7), (torch.bfloat16, 33, 999, 100, 40, 5)])
def test_rerank_sparse_candidates_then_dense_topk(dtype, nq, nc, dim, m, k):
    """report.pdf section 3.2: sparse top-100 candidates -> dense re-rank -> top-15 (site: evaluation.py:105-116)."""
    q, c = _data(nq, nc, dim, dtype, planted=True)
    g = torch.Generator(device=DEV).manual_seed(5)
    cand = torch.randint(0, nc, (nq, m), generator=g, device=DEV)
    cand[0, 1] = cand[0, 0]                                   # duplicate candidate -> reported once
    cand[1, m // 2:] = -1                                     # ragged list (padding)
    if nq > 2:
        cand[2, :] = -1                                       # no candidates at all
    s, i = drs_b200.rerank(q, c, cand, k)
    rs, ri = dense_topk.rerank(q.cpu(), c.cpu(), cand.cpu(), k)
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    fin = torch.isfinite(rs)
    assert torch.equal(torch.isfinite(s.cpu()), fin)
    torch.testing.assert_close(s.cpu()[fin], rs[fin], rtol=tol, atol=tol * 1e-1)
    gap_next = torch.cat([rs[:, :-1] - rs[:, 1:], torch.full((nq, 1), 1e30)], dim=1)
    gap_prev = torch.cat([torch.full((nq, 1), 1e30), rs[:, :-1] - rs[:, 1:]], dim=1)
    strict = fin & (gap_next > 1e-4) & (gap_prev > 1e-4)
    assert torch.equal(i.cpu()[strict], ri[strict])
    assert torch.equal(i.cpu()[~fin], ri[~fin])               # -1 padding in the same places
    # a DenseIndex shard with an id_base returns global ids
    idx = drs_b200.DenseIndex(c, dtype=dtype, id_base=1000)
    s2, i2 = idx.rerank(q, torch.where(cand >= 0, cand + 1000, cand), k)
    assert torch.equal(s2, s) and torch.equal(i2, torch.where(i >= 0, i + 1000, i))


@pytest.mark.parametrize("dtype,n,dim", [(torch.float32, 1000, 128), (torch.bfloat16, 513, 768), (torch.float32, 7, 50)])
def test_paired_scores_matches_reference_expression(dtype, n, dim):
    """src/evaluation.py:112: `(clm_vec * evdn_vec).sum(dim=-1).mean()`."""
    q, c = _data(n, n, dim, dtype, planted=True)
    out = drs_b200.paired_scores(q, c)
    ref = dense_topk.paired_scores(q.cpu(), c.cpu())
    tol = 1e-5 if dtype == torch.float32 else 1e-4
    torch.testing.assert_close(out.cpu(), ref, rtol=tol, atol=tol)
    assert abs(out.mean().item() - ref.mean().item()) < 1e-5


def test_index_file_round_trip_and_text_ranker(tmp_path):
    """save -> load (whole and as the 2 shards of a 2-rank job) -> same search results; DenseDocRanker is
    TfidfDocRanker's surface (path in, text queries in, (doc_ids, scores) out) with a stand-in encoder."""
    q, c = _data(50, 3000, 128, torch.bfloat16, planted=True)
    ids = [f"Page_{i}" for i in range(3000)]
    idx = drs_b200.DenseIndex(c, ids)
    path = str(tmp_path / "corpus.drsidx")
    idx.save(path, {"note": "test"})
    s0, i0 = idx.search(q, 10)
    loaded, meta = drs_b200.load_dense_index(path)
    assert meta["note"] == "test" and torch.equal(loaded.embeddings, idx.embeddings)
    s1, i1 = loaded.search(q, 10)
    assert torch.equal(s0, s1) and torch.equal(i0, i1)
    parts = [drs_b200.load_dense_index(path, rank=r, world_size=2)[0] for r in range(2)]
    assert [p.id_base for p in parts] == [0, 1500]
    ss, ii = zip(*[p.search(q, 10) for p in parts])
    ms, mi = drs_b200.merge_shards(torch.stack(ss), torch.stack(ii))
    assert torch.equal(ms, s0) and torch.equal(mi, i0)
    assert parts[1].closest_docs(q[0], 3)[0][0].startswith("Page_")
    table = {f"claim {n}": q[n] for n in range(50)}
    ranker = drs_b200.DenseDocRanker(path, encoder=lambda texts: torch.stack([table[t] for t in texts]))
    docs, scores = ranker.closest_docs("claim 3", 5)
    assert docs == [ids[j] for j in i0[3, :5].tolist()] and scores.dtype == np.float64
    batch = ranker.batch_closest_docs(["claim 1", "claim 2"], 2)
    assert [b[0][0] for b in batch] == [ids[i0[1, 0].item()], ids[i0[2, 0].item()]]


@pytest.mark.parametrize("dtype,k", [(torch.bfloat16, 10), (torch.bfloat16, 100), (torch.float32, 10)])
def test_threshold_seeding_keeps_ties_and_matches_unseeded(dtype, k):
    """Units that finish early publish their k-th best score and later units of the same claim start
    from it (epilogues.cuh).  With few CTAs the corpus splits run one after another, so the seeds are
    live; exact duplicates of the best row sit in a LATE split (published first to nobody) and in an
    EARLY one: the answer must be the lowest row numbers, identical to the run without seeding."""
    nq, nc, dim = 40, 74 * 1024, 128
    q, c = _data(nq, nc, dim, dtype, planted=True)
    dup = c[70000].clone()
    c[60000:60012] = dup
    c[100:112] = dup
    c[70000] = c[5]
    q[0] = dup
    try:
        drs_b200.set_option("search.num_ctas", 4)
        s1, i1 = drs_b200.search(q, c, k)
        drs_b200.set_option("tune.seed_thresholds", 0)
        s0, i0 = drs_b200.search(q, c, k)
    finally:
        drs_b200.set_option("search.num_ctas", 0)
        drs_b200.set_option("tune.seed_thresholds", 1)
    assert torch.equal(i1, i0) and torch.equal(s1, s0)
    assert i1[0, :10].cpu().tolist() == list(range(100, 110))
    tol = (1e-5, 5e-6) if dtype == torch.float32 else (2e-2, 1e-4)
    _check(q, c, k, s1, i1, score_rtol=tol[0], gap=tol[1])


def test_randomised_shape_sweep_against_the_oracle():
    """Seeded sweep over awkward shapes (tiny and ragged claims / corpus / dim, k from 1 to 256, both dtypes):
    ids wherever the gap allows, scores within tolerance, descending order, -1 padding when k > Nc."""
    rng = np.random.RandomState(1337)
    dims = [8, 16, 24, 72, 128, 200, 768]
    for case in range(24):
        dtype = torch.bfloat16 if case % 3 else torch.float32
        nq = int(rng.choice([1, 2, 3, 17, 127, 129, 257, 700]))
        nc = int(rng.choice([1, 5, 33, 255, 257, 1000, 4097, 30011]))
        dim = int(rng.choice(dims)) if dtype == torch.bfloat16 else int(rng.choice(dims + [7, 50]))
        k = int(rng.choice([1, 2, 5, 10, 16, 17, 32, 33, 100, 256]))
        q, c = _data(nq, nc, dim, dtype, planted=bool(case % 2), seed=1000 + case)
        s, i = drs_b200.search(q, c, k)
        assert s.shape == (nq, min(k, nc)), (case, nq, nc, dim, k)
        tol = (1e-5, 5e-6) if dtype == torch.float32 else (2e-2, 1e-4)
        _check(q, c, k, s, i, score_rtol=tol[0], gap=tol[1])


def test_paired_scores_golden(golden_dir):
    """The reference's own expression (src/evaluation.py:112), run by tests/golden/make_golden.py."""
    import os
    z = np.load(os.path.join(golden_dir, "paired.npz"))
    clm, evdn = torch.from_numpy(z["clm"]).to(DEV), torch.from_numpy(z["evdn"]).to(DEV)
    out = drs_b200.paired_scores(clm, evdn)
    np.testing.assert_allclose(out.cpu().numpy(), z["per_pair"], rtol=1e-5, atol=1e-6)
    assert abs(out.mean().item() - float(z["mean"])) < 1e-6


@pytest.mark.parametrize("nq", [16, 600])
def test_search_is_cuda_graph_capturable(nq):
    """Serving loop: the staging + scan + select launches of a search captured once in a CUDA graph and replayed on new
    claims (no host-side planning or launches per query batch).  16 claims: a single-round scan; 600 claims: a
    multi-round scan, i.e. the round barrier and its COOPERATIVE launch inside a captured graph."""
    q, c = _data(nq, 200_000, 768, torch.bfloat16, planted=True)
    static_q = q.clone()
    drs_b200.search(static_q, c, 10)                                # warm-up: lazy init happens outside the capture
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        s, i = drs_b200.search(static_q, c, 10)
    for seed in (1, 2):
        q2, _ = _data(nq, 200_000, 768, torch.bfloat16, planted=True, seed=seed)
        q2 = torch.nn.functional.normalize(c[seed * 100: seed * 100 + nq].float() + 0.05 * q2.float(), dim=1).to(torch.bfloat16)
        static_q.copy_(q2)
        graph.replay()
        torch.cuda.synchronize()
        es, ei = drs_b200.search(q2, c, 10)
        assert torch.equal(i, ei) and torch.equal(s, es)
        assert ei[:, 0].cpu().tolist() == list(range(seed * 100, seed * 100 + nq))


@pytest.mark.parametrize("nq,nc,dim,k", [(64, 5000, 128, 5), (300, 40000, 768, 10), (1000, 30000, 256, 100)])
def test_fp16_operands_on_the_tensor_core_path(nq, nc, dim, k):
    """IEEE half embeddings (DRS_F16): same tcgen05 kernel with the f16 instruction descriptor; the oracle takes
    the same fp16 values upcast to fp32.  Also through the squared-L2 epilogue, the re-rank and the pair score."""
    q, c = _data(nq, nc, dim, torch.float16, planted=True)
    s, i = drs_b200.search(q, c, k)
    _check(q, c, k, s, i, score_rtol=2e-2, gap=1e-4)
    d, li = drs_b200.flat_l2_search(q[:50], c, 1)
    rd, rli = dense_topk.flat_l2_search(q[:50].cpu(), c.cpu(), 1)
    assert torch.equal(li.cpu(), rli)
    torch.testing.assert_close(d.cpu(), rd, rtol=2e-2, atol=2e-3)
    g = torch.Generator(device=DEV).manual_seed(3)
    cand = torch.randint(0, nc, (nq, 50), generator=g, device=DEV)
    rs, ri = drs_b200.rerank(q, c, cand, 5)
    os_, oi = dense_topk.rerank(q.cpu(), c.cpu(), cand.cpu(), 5)
    torch.testing.assert_close(rs.cpu(), os_, rtol=1e-4, atol=1e-5)
    assert torch.equal(ri.cpu()[:, 0], oi[:, 0])
    torch.testing.assert_close(drs_b200.paired_scores(q, c[:nq]).cpu(), dense_topk.paired_scores(q.cpu(), c[:nq].cpu()),
                               rtol=1e-4, atol=1e-5)
    idx = drs_b200.DenseIndex(c, dtype=torch.float16)
    assert idx.embeddings.dtype == torch.float16 and torch.equal(idx.search(q, k)[1], i)


def test_many_claims_and_rerank_extremes():
    """More claims than one grid of A tiles can hold at once (70 000 -> 274 pair tiles, seeded thresholds, ragged
    last tile) against the oracle on a sample; re-rank with a single candidate and with 3000 candidates."""
    nq, nc, dim, k = 70_000, 20_000, 64, 10
    q, c = _data(nq, nc, dim, torch.bfloat16, planted=True)
    s, i = drs_b200.search(q, c, k)
    pick = torch.randint(0, nq, (200,), generator=torch.Generator().manual_seed(0))
    pick[0], pick[1] = 0, nq - 1
    _check(q[pick.to(DEV)], c, k, s[pick.to(DEV)], i[pick.to(DEV)], score_rtol=2e-2, gap=1e-4)
    g = torch.Generator(device=DEV).manual_seed(9)
    for m, kk in ((1, 1), (3000, 15)):
        cand = torch.randint(0, nc, (64, m), generator=g, device=DEV)
        rs, ri = drs_b200.rerank(q[:64], c, cand, kk)
        os_, oi = dense_topk.rerank(q[:64].cpu(), c.cpu(), cand.cpu(), kk)
        torch.testing.assert_close(rs.cpu(), os_, rtol=1e-4, atol=1e-5)
        assert torch.equal(ri.cpu()[:, 0], oi[:, 0])


def test_claim_batches_beyond_the_per_pass_limit_are_sliced(monkeypatch):
    from importlib import import_module
    retrieval = import_module(drs_b200.__name__ + ".retrieval")
    q, c = _data(1000, 5000, 64, torch.bfloat16, planted=True)
    s0, i0 = drs_b200.search(q, c, 10)
    monkeypatch.setattr(retrieval, "MAX_CLAIMS_PER_PASS", 300)
    s1, i1 = drs_b200.search(q, c, 10)
    assert torch.equal(i0, i1) and torch.equal(s0, s1)
