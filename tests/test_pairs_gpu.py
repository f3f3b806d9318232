"""GPU parity: per-document sentence-pair similarity (build_docs_sentence_similarity.py:41-68)."""
import os

import numpy as np
import pytest
import torch

import drs_b200
from oracle import pairs as oracle_pairs

pytestmark = pytest.mark.gpu


def _golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "pairs.npz"))
    n = int(z["ndocs"])
    return [z[f"x{d}"] for d in range(n)], [z[f"pairs{d}"] for d in range(n)], [z[f"scores{d}"] for d in range(n)]


def test_golden_pairs_bit_exact(golden_dir):
    """Golden vectors made by running sklearn's cosine_similarity + the reference loop (:50-65) in the
    build container (tests/golden/make_golden.py): same pairs, same order, scores equal to the last bit
    (incl. the one-sentence document, the duplicated sentence -> tied scores, the empty sentence)."""
    import scipy.sparse as sp
    xs, gp, gs = _golden(golden_dir)
    out = drs_b200.docs_sentence_pairs([sp.csr_matrix(x) for x in xs])
    assert len(out) == len(xs)
    for d, doc in enumerate(out):
        assert [list(p[0]) for p in doc] == gp[d].tolist(), f"doc {d}: pair order differs"
        got = np.array([p[1] for p in doc], dtype=np.float64)
        assert got.tobytes() == gs[d].tobytes(), f"doc {d}: scores differ from the reference bit pattern"
    dense = drs_b200.docs_sentence_pairs(xs)                      # dense input takes the same route
    assert dense == out


def test_random_documents_against_sklearn_and_the_oracle():
    """Ragged batch (0, 1, 2 ... 150 sentences; the largest sorts outside shared memory) against
    sklearn's sparse cosine_similarity + the reference loop restated here, bit for bit, and against the
    dense fp64 oracle within 1e-12."""
    import scipy.sparse as sp
    from sklearn.metrics.pairwise import cosine_similarity
    rng = np.random.RandomState(7)
    sizes = [5, 1, 0, 2, 40, 13, 150, 3, 1, 64, 129]
    docs = []
    for n in sizes:
        x = rng.rand(n, 500) * (rng.rand(n, 500) < 0.05)
        if n >= 13:
            x[7] = x[2]                                            # exact ties
            x[5] = 0.0                                             # empty sentence
        docs.append(sp.csr_matrix(x))
    out = drs_b200.docs_sentence_pairs(docs)
    assert [len(o) for o in out] == [1 if n == 1 else n * (n - 1) // 2 for n in sizes]
    for d, (n, x) in enumerate(zip(sizes, docs)):
        if n == 0:
            assert out[d] == []
            continue
        sim = cosine_similarity(x, x)                              # :50
        ref = []
        if n == 1:
            ref.append(((0, 0), sim[0][0]))
        for i in range(n):
            for j in range(i + 1, n):
                ref.append(((i, j), sim[i][j]))
        ref.sort(key=lambda t: t[1], reverse=True)                # :65
        assert [p[0] for p in out[d]] == [p[0] for p in ref], f"doc {d}: pair order differs"
        assert np.array([p[1] for p in out[d]]).tobytes() == np.array([p[1] for p in ref]).tobytes()
        orc = oracle_pairs.doc_sentence_pairs(x.toarray())
        np.testing.assert_allclose([p[1] for p in out[d]], [p[1] for p in orc], rtol=0, atol=1e-12)


def test_get_docs_sents_similarity_mirrors_the_reference_signature():
    """(full_data, small_data) of sentence strings in, list[doc] of [((i, j), score)] out; the vectoriser is
    the caller's (the reference's LemmaTokenizer needs nltk, which this image does not have)."""
    from sklearn.feature_extraction.text import TfidfVectorizer
    from sklearn.metrics.pairwise import cosine_similarity
    full = [["the cat sat on the mat", "a dog barked at the cat", "the mat was red"],
            ["paris is the capital of france", "france is in europe"],
            ["one sentence only"]]
    vec = TfidfVectorizer(ngram_range=(1, 2))
    out = drs_b200.get_docs_sents_similarity(full, full[:2] + [full[2]], vectorizer=vec)
    assert len(out) == 3 and len(out[0]) == 3 and len(out[1]) == 1 and out[2][0][0] == (0, 0)
    sim = cosine_similarity(vec.transform(full[0]), vec.transform(full[0]))
    assert out[0][0][1] == max(sim[0][1], sim[0][2], sim[1][2])
    assert all(a[1] >= b[1] for a, b in zip(out[0], out[0][1:]))
    k = -(-len(out[0]) // 10)                                     # src/dataset.py:96: the top 10 % the trainer samples from
    assert len(out[0][:k]) == 1
