"""CPU oracle: dense claim x corpus scoring and top-k select.  TEST INFRASTRUCTURE.

The reference has no running dense retrieval code; its intended site is the
commented-out block at src/evaluation.py:105-116 (``ctx2vec`` embeddings,
dot-product scores).  This restates the pieces it would be made of:

* rows are L2-normalised as ``RetrievalModelWrapper.seq2vec`` does
  (src/contrastor/contrastive_module.py:111, ``nn.functional.normalize``);
* scores are the fp32 dot products the reference computes with
  ``torch.matmul(features, features.T)`` (src/contrastor/contrastive_loss.py:62)
  and ``(clm_vec * evdn_vec).sum(dim=-1)`` (src/evaluation.py:112);
* the select keeps the k best, sorted descending, like
  ``TfidfDocRanker.closest_docs``
  (preprocessing/drqa/retriever/tfidf_doc_ranker.py:67-73: ``argpartition`` of
  the negated scores, then ``argsort`` of the k survivors).  ``argpartition``
  leaves the order of equal scores unspecified; BASELINE.json's north star
  fixes it: ties are broken by the LOWER index.

A bf16 corpus is scored from the same bf16 values upcast to fp32
(SURVEY.md section 8c), so the only difference to the GPU path is summation order.
"""
from __future__ import annotations

import numpy as np
import torch


def normalize_rows(x: torch.Tensor) -> torch.Tensor:
    """contrastive_module.py:111 -- ``nn.functional.normalize(emb)`` (p=2, dim=1)."""
    return torch.nn.functional.normalize(x, dim=1)


def scores_fp32(queries: torch.Tensor, corpus: torch.Tensor) -> torch.Tensor:
    """contrastive_loss.py:62 idiom: one fp32 ``torch.matmul`` of row-major operands."""
    return torch.matmul(queries.float(), corpus.float().T)


def select_topk_desc(scores: torch.Tensor, k: int, index_base: int = 0):
    """k best per row, sorted by (score descending, index ascending).

    tfidf_doc_ranker.py:67-73 semantics (fewer than k candidates -> return them
    all, i.e. k is clamped) with the north-star tie rule made explicit.
    Returns (values fp32 [n, k'], ids int64 [n, k']).
    """
    n, m = scores.shape
    k = min(k, m)
    if k == 0:
        return (torch.empty(n, 0, dtype=torch.float32), torch.empty(n, 0, dtype=torch.int64))
    scores = scores.float()
    if m <= 4096:
        # a stable descending sort keeps equal scores in index order
        vals, idx = torch.sort(scores, dim=1, descending=True, stable=True)
        return vals[:, :k].contiguous(), (idx[:, :k] + index_base).contiguous()
    # Wide rows: the same order without a full sort.  Map each fp32 score to an integer
    # that sorts like the float (-0.0 folded into +0.0), append the complemented column
    # index as the low word, and take the k largest of those unique 64-bit keys.
    out_v = torch.empty(n, k, dtype=torch.float32)
    out_i = torch.empty(n, k, dtype=torch.int64)
    col = (0xFFFFFFFF - torch.arange(m, dtype=torch.int64))[None, :]
    rows_per = max(1, (1 << 27) // m)
    for r0 in range(0, n, rows_per):
        s = scores[r0:r0 + rows_per] + 0.0
        b = s.contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
        ordv = torch.where(b >= 0x80000000, 0xFFFFFFFF - b, b + 0x80000000) - 0x80000000
        key = ordv * (1 << 32) + col
        _, idx = torch.topk(key, k, dim=1, largest=True, sorted=True)
        out_v[r0:r0 + rows_per] = torch.gather(s, 1, idx)
        out_i[r0:r0 + rows_per] = idx + index_base
    return out_v, out_i


def _merge_topk(vals_a, ids_a, vals_b, ids_b, k):
    """Merge two (score desc, id asc) lists per row into the k best of their union."""
    vals = torch.cat([vals_a, vals_b], dim=1)
    ids = torch.cat([ids_a, ids_b], dim=1)
    # order by id first (stable), then by score descending (stable) -> (score desc, id asc)
    o = torch.argsort(ids, dim=1, stable=True)
    vals, ids = torch.gather(vals, 1, o), torch.gather(ids, 1, o)
    o = torch.argsort(vals, dim=1, descending=True, stable=True)
    vals, ids = torch.gather(vals, 1, o), torch.gather(ids, 1, o)
    k = min(k, vals.shape[1])
    return vals[:, :k].contiguous(), ids[:, :k].contiguous()


def search(queries: torch.Tensor, corpus: torch.Tensor, k: int, chunk_rows: int = 65536,
           index_base: int = 0):
    """Exact dense top-k of ``queries @ corpus.T`` streamed over corpus row chunks.

    Never materialises more than ``nq x chunk_rows`` scores, so corpora far larger
    than host RAM's nq x Nc matrix can be checked.  Ties -> lower index.
    """
    nq = queries.shape[0]
    best_v = torch.empty(nq, 0, dtype=torch.float32)
    best_i = torch.empty(nq, 0, dtype=torch.int64)
    qf = queries.float()
    for lo in range(0, corpus.shape[0], chunk_rows):
        blk = corpus[lo:lo + chunk_rows]
        v, i = select_topk_desc(scores_fp32(qf, blk), k, index_base + lo)
        best_v, best_i = _merge_topk(best_v, best_i, v, i, k)
    return best_v, best_i


def search_fast(queries: torch.Tensor, corpus: torch.Tensor, k: int, chunk_rows: int = 262144):
    """The reference idiom at speed, for the CPU baseline timing only.

    fp32 ``torch.matmul`` (contrastive_loss.py:62) + ``torch.topk`` (the
    ``argpartition``+``argsort`` of tfidf_doc_ranker.py:70-71 in one call).
    Tie order is whatever ``torch.topk`` yields, so parity tests use ``search``.
    """
    qf = queries.float()
    best_v = best_i = None
    for lo in range(0, corpus.shape[0], chunk_rows):
        s = torch.matmul(qf, corpus[lo:lo + chunk_rows].float().T)
        v, i = torch.topk(s, min(k, s.shape[1]), dim=1)
        i = i + lo
        if best_v is None:
            best_v, best_i = v, i
        else:
            v = torch.cat([best_v, v], 1)
            i = torch.cat([best_i, i], 1)
            best_v, o = torch.topk(v, min(k, v.shape[1]), dim=1)
            best_i = torch.gather(i, 1, o)
    return best_v, best_i


def closest_docs_select(scores_1d: np.ndarray, k: int):
    """The select of tfidf_doc_ranker.py:67-71 on a dense 1-D score vector, tie rule made
    explicit (stable descending argsort).  Used to pin ``select_topk_desc`` against
    ``TfidfDocRanker.closest_docs`` run on a synthetic CSR matrix."""
    order = np.argsort(-scores_1d, kind="stable")
    return order[:k]


def sharded_search(queries: torch.Tensor, corpus: torch.Tensor, k: int, world_size: int):
    """contiguous row shards, per-shard top-k with global ids, then a
    (score desc, id asc) merge.  Must equal ``search`` on the whole corpus."""
    nc = corpus.shape[0]
    per = -(-nc // world_size)
    best_v = torch.empty(queries.shape[0], 0, dtype=torch.float32)
    best_i = torch.empty(queries.shape[0], 0, dtype=torch.int64)
    for r in range(world_size):
        lo, hi = r * per, min(nc, (r + 1) * per)
        if lo >= hi:
            continue
        v, i = search(queries, corpus[lo:hi], k, index_base=lo)
        best_v, best_i = _merge_topk(best_v, best_i, v, i, k)
    return best_v, best_i


def flat_l2_search(x: torch.Tensor, centroids: torch.Tensor, k: int = 1):
    """PARITY UNPINNED.  src/contrastor/utils.py:64-67: ``index.search(x, 1)`` on a faiss
    ``GpuIndexFlatL2`` = exact squared-L2 nearest neighbours, ascending distance.
    faiss (requirements.txt:6, unpinned) is absent; this is its documented result.
    Returns (squared distances [n,k], ids [n,k]); ties -> lower index."""
    xf, cf = x.double(), centroids.double()
    d = (xf * xf).sum(1, keepdim=True) + (cf * cf).sum(1)[None, :] - 2.0 * xf @ cf.T
    vals, idx = torch.sort(d, dim=1, descending=False, stable=True)
    return vals[:, :k].float().contiguous(), idx[:, :k].contiguous()


def rerank(queries: torch.Tensor, corpus: torch.Tensor, cand_ids: torch.Tensor, k: int):
    """Candidate-restricted scoring: report.pdf section 3.2 ("TF-IDF top-100 -> re-rank -> top-15"),
    intended at src/evaluation.py:105-116.  Per claim: gather its candidate rows, dot products in fp32
    (the idiom of evaluation.py:112), select like closest_docs (tfidf_doc_ranker.py:67-73) with the
    lower-row tie rule.  Negative / out-of-range ids are padding; a row listed twice counts once.
    Returns (scores fp32 [nq, k'], ids int64 [nq, k']), k' = min(k, m), padded with (-inf, -1)."""
    q = queries.float()
    c = corpus.float()
    nq, m = cand_ids.shape
    kk = min(k, m)
    out_s = torch.full((nq, kk), float("-inf"), dtype=torch.float32)
    out_i = torch.full((nq, kk), -1, dtype=torch.int64)
    for r in range(nq):
        ids = [int(v) for v in cand_ids[r].tolist() if 0 <= int(v) < c.shape[0]]
        ids = sorted(set(ids))                       # ascending row: a stable descending sort keeps ties low-row first
        if not ids:
            continue
        idt = torch.tensor(ids, dtype=torch.int64)
        sc = (c[idt] * q[r][None, :]).sum(dim=-1)    # (clm_vec * evdn_vec).sum(dim=-1)
        order = torch.argsort(sc, descending=True, stable=True)[:kk]
        out_s[r, : order.numel()] = sc[order]
        out_i[r, : order.numel()] = idt[order]
    return out_s, out_i


def paired_scores(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """src/evaluation.py:112 -- ``(clm_vec * evdn_vec).sum(dim=-1)`` (fp32)."""
    return (a.float() * b.float()).sum(dim=-1)
