"""CPU oracle for the dense-retrieval / InfoNCE hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or as
the timed CPU baseline.  The product package must never import this module.

Each function restates, in plain torch/numpy on the CPU, what the reference
(PM25/Information-Retrieval-with-Contrastive-Learning) computes on the path
named by BASELINE.json and cites the reference file:line it follows.

Pinning status (see tests/golden/make_golden.py, tests/test_oracle_golden.py):

* ``infonce``  -- pinned.  ``NCELoss`` / ``InfoNCE`` are imported from the
  reference itself (src/contrastor/contrastive_loss.py) in the build
  container and their loss values and autograd gradients are committed as
  golden vectors; the restatement here must reproduce them.
* ``dense_topk`` -- the reference has no dense top-k implementation (the call
  site src/evaluation.py:105-116 is commented out).  The restatement follows
  the reference's scoring idiom (torch.matmul, contrastive_loss.py:62) and the
  select semantics of TfidfDocRanker.closest_docs
  (preprocessing/drqa/retriever/tfidf_doc_ranker.py:60-75); the select step is
  pinned against closest_docs itself, run on a synthetic CSR matrix.
* ``pairs`` -- pinned against sklearn's ``cosine_similarity`` + the reference's
  loop (preprocessing/build_docs_sentence_similarity.py:49-66) restated from a
  caller-supplied matrix (the module itself cannot be imported: it needs nltk
  and downloads corpora at import time).
* ``flat_l2`` -- PARITY UNPINNED: the arithmetic lives in faiss
  (requirements.txt:6, unpinned version, not installed, not vendored); the
  restatement is the exact L2 arg-min faiss's IndexFlatL2 is documented to
  return.
"""
