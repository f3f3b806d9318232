"""CPU oracle: per-document sentence-pair similarity.  TEST INFRASTRUCTURE.

Restates the body of ``get_docs_sents_similarity``
(preprocessing/build_docs_sentence_similarity.py:48-66) from a caller-supplied
sentence-vector matrix per document (the TF-IDF vectorisation at :43-45,:49 is out of
scope; the module itself cannot be imported here: it needs nltk and downloads corpora
at import time, :14-21).
"""
from __future__ import annotations

import numpy as np


def cosine_similarity(x: np.ndarray) -> np.ndarray:
    """sklearn.metrics.pairwise.cosine_similarity(X, X) (:50): rows L2-normalised
    (all-zero rows stay zero), then X_n @ X_n.T, in float64."""
    x = np.asarray(x, dtype=np.float64)
    nrm = np.sqrt((x * x).sum(axis=1, keepdims=True))
    nrm[nrm == 0.0] = 1.0
    xn = x / nrm
    return xn @ xn.T


def doc_sentence_pairs(x: np.ndarray):
    """One document: all (i<j) pairs with their cosine, sorted by score descending
    (:59-65).  ``list.sort`` is stable, so equal scores keep (i, j) lexicographic order.
    A single-sentence document yields [((0, 0), s00)] (:54-57)."""
    sim = cosine_similarity(x)
    n = sim.shape[0]
    out = []
    if n == 1:
        out.append(((0, 0), sim[0][0]))
    for i in range(n):
        for j in range(i + 1, n):
            out.append(((i, j), sim[i][j]))
    out.sort(key=lambda t: t[1], reverse=True)
    return out


def docs_sents_similarity(doc_matrices):
    """:47-68 -- the list over documents."""
    return [doc_sentence_pairs(x) for x in doc_matrices]
