"""Per-document sentence-pair similarity on the GPU.

Host-side mirror of ``get_docs_sents_similarity``
(preprocessing/build_docs_sentence_similarity.py:41-68).  The TF-IDF vectorisation (:43-45, :49)
stays with scikit-learn on the CPU (out of scope: text processing); everything after it -- the
cosine similarity of a document's sentences with themselves (:50), the walk over the strict upper
triangle (:59-63), the one-sentence special case (:54-57) and the stable descending sort (:65) --
runs in csrc/pairs.cuh for all documents in one launch, in float64, in the reference's own order
of floating-point operations (results are bit-identical to the reference).

The return value has the reference's structure: ``list[doc] of list[((i, j), score)]``, which is
what ``DocDataset`` consumes (src/dataset.py:95-99: a random pair among the top 10 %).
"""
from __future__ import annotations

import ctypes
from typing import Sequence

import numpy as np
import torch

from . import _lib


def _as_csr(m):
    """One document's sentence matrix -> scipy CSR, float64, sorted indices, no explicit duplicates."""
    import scipy.sparse as sp
    if isinstance(m, torch.Tensor):
        m = m.detach().cpu().numpy()
    x = m.tocsr() if sp.issparse(m) else sp.csr_matrix(np.asarray(m, dtype=np.float64))
    x = x.astype(np.float64, copy=False)
    if not x.has_canonical_format:
        x = x.copy()
        x.sum_duplicates()
    return x


def pair_counts(sentences_per_doc: np.ndarray) -> np.ndarray:
    """n(n-1)/2 pairs per document; a one-sentence document contributes its (0, 0) entry (:54-57)."""
    n = np.asarray(sentences_per_doc, dtype=np.int64)
    return np.where(n == 1, 1, n * (n - 1) // 2)


def doc_sentence_pairs_arrays(doc_matrices: Sequence, device=None):
    """The engine call: a batch of per-document sentence matrices (scipy sparse / numpy / tensor,
    one row per sentence) -> (pair_offsets int64 [ndocs+1], i int32, j int32, score float64), the
    pairs of document d at ``[pair_offsets[d], pair_offsets[d+1])`` in the reference's order."""
    import scipy.sparse as sp
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
    if device is None:
        raise RuntimeError("drs_b200 needs a CUDA device: there is no CPU path")
    dev = torch.device(device)
    mats = [_as_csr(m) for m in doc_matrices]
    ndocs = len(mats)
    nsent = np.array([m.shape[0] for m in mats], dtype=np.int64)
    doc_offsets = np.zeros(ndocs + 1, dtype=np.int64)
    np.cumsum(nsent, out=doc_offsets[1:])
    pair_offsets = np.zeros(ndocs + 1, dtype=np.int64)
    np.cumsum(pair_counts(nsent), out=pair_offsets[1:])
    total_pairs = int(pair_offsets[-1])
    out_i = torch.empty(total_pairs, dtype=torch.int32, device=dev)
    out_j = torch.empty(total_pairs, dtype=torch.int32, device=dev)
    out_s = torch.empty(total_pairs, dtype=torch.float64, device=dev)
    if total_pairs == 0:
        return pair_offsets, out_i.cpu().numpy(), out_j.cpu().numpy(), out_s.cpu().numpy()
    width = max(m.shape[1] for m in mats)
    nonempty = [sp.csr_matrix((m.data, m.indices, m.indptr), shape=(m.shape[0], width)) for m in mats if m.shape[0]]
    stacked = sp.vstack(nonempty, format="csr")
    nnz = int(stacked.nnz)
    indptr = torch.from_numpy(stacked.indptr.astype(np.int64)).to(dev)
    indices = torch.from_numpy(stacked.indices.astype(np.int32)).to(dev)
    data = torch.from_numpy(stacked.data.astype(np.float64)).to(dev)
    d_off = torch.from_numpy(doc_offsets).to(dev)
    p_off = torch.from_numpy(pair_offsets).to(dev)
    lib = _lib.load()
    with torch.cuda.device(dev):
        need = ctypes.c_size_t(0)
        _lib.check(lib.drs_doc_pairs_workspace_bytes(nnz, total_pairs, ctypes.byref(need)))
        ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.drs_doc_sentence_pairs(indptr.data_ptr(), indices.data_ptr() if nnz else None,
                                              data.data_ptr() if nnz else None, int(doc_offsets[-1]), d_off.data_ptr(),
                                              ndocs, p_off.data_ptr(), total_pairs, nnz, out_i.data_ptr(),
                                              out_j.data_ptr(), out_s.data_ptr(), ws.data_ptr(), ws.numel(), stream))
    return pair_offsets, out_i.cpu().numpy(), out_j.cpu().numpy(), out_s.cpu().numpy()


def docs_sentence_pairs(doc_matrices: Sequence, device=None):
    """Lines :47-68 for pre-vectorised documents: ``list[doc] of list[((i, j), score)]`` sorted by score
    descending (ties keep (i, j) order), scores as Python floats of the float64 values."""
    off, pi, pj, ps = doc_sentence_pairs_arrays(doc_matrices, device=device)
    pi, pj, ps = pi.tolist(), pj.tolist(), ps.tolist()
    return [[((pi[t], pj[t]), ps[t]) for t in range(int(off[d]), int(off[d + 1]))] for d in range(len(off) - 1)]


def _reference_vectorizer():
    """The vectoriser the reference builds at build_docs_sentence_similarity.py:43: word-level TF-IDF over uni- and
    bigrams whose tokenizer (:27-38) drops punctuation marks and English stop words (nltk corpus, :17-22) and
    lemmatises what is left with WordNet.  Needs nltk and its data, exactly like the reference."""
    import string
    from sklearn.feature_extraction.text import TfidfVectorizer
    try:
        from nltk import word_tokenize
        from nltk.corpus import stopwords as nltk_stopwords
        from nltk.stem import WordNetLemmatizer
        stop = set(nltk_stopwords.words("english"))
    except (ImportError, LookupError) as e:  # the reference imports nltk and downloads its corpora at module scope
        raise RuntimeError("the reference's LemmaTokenizer needs nltk with the wordnet, stopwords and punkt data; "
                           "pass vectorizer=... instead") from e
    lemmatizer = WordNetLemmatizer()

    def lemma_tokens(sentence):
        kept = [tok for tok in word_tokenize(sentence) if tok not in string.punctuation and tok not in stop]
        return [lemmatizer.lemmatize(tok) for tok in kept]

    return TfidfVectorizer(tokenizer=lemma_tokens, ngram_range=(1, 2))


def get_docs_sents_similarity(full_data, small_data, vectorizer=None, device=None):
    """Drop-in for build_docs_sentence_similarity.py:41-68.

    ``full_data`` / ``small_data``: lists of documents, each a list of sentence strings.  ``vectorizer``:
    a scikit-learn vectoriser; by default the reference's ``TfidfVectorizer(tokenizer=LemmaTokenizer(),
    ngram_range=(1, 2))`` (:43), which needs nltk -- pass your own when nltk is not installed.  It is
    fitted on every sentence of ``full_data`` (:44-45) unless it already has a vocabulary."""
    if vectorizer is None:
        vectorizer = _reference_vectorizer()
    if not hasattr(vectorizer, "vocabulary_"):
        vectorizer.fit([sent for doc in full_data for sent in doc])
    mats = [vectorizer.transform(doc) for doc in small_data]
    return docs_sentence_pairs(mats, device=device)
