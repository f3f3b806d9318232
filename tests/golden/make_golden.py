#!/usr/bin/env python
"""Generate the golden vectors in this directory by RUNNING THE REFERENCE ITSELF.

Run in the build container (``python tests/golden/make_golden.py``), where the
reference is mounted read-only at /root/reference.  The GPU box has no reference, so
the outputs are committed as small ``.npz`` fixtures and the tests only read those.

What is executed, unmodified, from /root/reference:

* ``src.contrastor.contrastive_loss.NCELoss`` (forward + autograd backward) with and
  without the MoCo queue -> infonce_*.npz
* ``src.contrastor.contrastive_loss.InfoNCE`` (MoCo form)             -> moco_*.npz
* ``NCELoss._compute_proto_loss`` with the prototype sampling replaced by a fixed
  selection (``random.sample(set, r)`` at :109 raises on Python >= 3.11, so the
  reference's own line cannot run here; everything after :110 is the reference's code,
  executed through a subclass that only overrides the sampling)         -> proto_*.npz
* ``preprocessing.drqa.retriever.TfidfDocRanker.closest_docs`` select logic on a
  synthetic CSR matrix (the class is built with ``__new__`` and given a hand-made
  ``doc_mat`` because its constructor wants a saved index and a tokenizer; ``pexpect``
  is stubbed so the package imports)                                    -> closest_docs.npz
* sklearn's ``cosine_similarity`` + the loop at
  preprocessing/build_docs_sentence_similarity.py:52-65, restated verbatim in a local
  function because the module downloads nltk corpora at import time     -> pairs.npz
* the commented dense diagnostic of src/evaluation.py:112, evaluated literally           -> paired.npz
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("DRS_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


def _unit(x, dim=1):
    return torch.nn.functional.normalize(x, dim=dim)


def gen_infonce():
    sys.path.insert(0, REF)
    from src.contrastor.contrastive_loss import NCELoss, InfoNCE  # the reference

    cases = [  # name, N, D, K (queue, 0 = none), T
        ("n8_d16", 8, 16, 0, 0.05),
        ("n8_d16_q32", 8, 16, 32, 0.05),
        ("n32_d64", 32, 64, 0, 0.05),
        ("n32_d64_q96", 32, 64, 96, 0.07),
        ("n128_d128_q512", 128, 128, 512, 0.05),      # the reference's own N, D (config.yaml:7,87)
        ("n96_d768", 96, 768, 0, 0.05),               # BASELINE D
    ]
    for name, n, d, kq, temp in cases:
        g = torch.Generator().manual_seed(1337)
        q = _unit(torch.randn(n, d, generator=g)).requires_grad_(True)
        k = _unit(torch.randn(n, d, generator=g) * 0.5 + q.detach()).requires_grad_(True)
        queue = _unit(torch.randn(d, kq, generator=g), dim=0) if kq else None
        crit = NCELoss({"temperature": temp})
        loss = crit(q, k, queue)
        loss.backward()
        out = dict(q=q.detach().numpy(), k=k.detach().numpy(), temperature=np.float64(temp),
                   loss=loss.detach().numpy(), dq=q.grad.numpy(), dk=k.grad.numpy())
        if queue is not None:
            out["queue"] = queue.numpy()
        np.savez_compressed(os.path.join(HERE, f"infonce_{name}.npz"), **out)

        if queue is not None:
            q2 = q.detach().clone().requires_grad_(True)
            k2 = k.detach().clone().requires_grad_(True)
            loss2 = InfoNCE({"temperature": temp})(q2, k2, queue)
            loss2.backward()
            np.savez_compressed(
                os.path.join(HERE, f"moco_{name}.npz"), q=q2.detach().numpy(), k=k2.detach().numpy(),
                queue=queue.numpy(), temperature=np.float64(temp), loss=loss2.detach().numpy(),
                dq=q2.grad.numpy(), dk=k2.grad.numpy())

    # ProtoNCE: the reference's code after the sampling line, through a subclass
    import src.contrastor.contrastive_loss as ref_loss

    class _FixedSample:
        """stands in for ``random.sample`` at contrastive_loss.py:109"""
        def __call__(self, population, r):
            return sorted(population)[:r]

    from proto_inputs import make_inputs, checksum

    # (name, N, D, clusters per set, r, compact): N + r = 22 / 84 (ragged: not a multiple of 8), 88 (aligned), and the
    # reference's own shapes -- batch 128, 128-d, 3072 negatives (config.yaml:7,29-30,87) -> 3200 columns.  `compact`
    # fixtures store q, index and the outputs only; the test regenerates the cluster sets from the seed.
    for name, n, d, ncl, r, compact in [("n16_d32", 16, 32, [24, 40], 6, False), ("n64_d128", 64, 128, [96], 20, False),
                                        ("n64_d128_r24", 64, 128, [96, 160], 24, False),
                                        ("n128_d128_r3072", 128, 128, [4096, 6144], 3072, True)]:
        q, index, cluster_result = make_inputs(n, d, ncl, seed=1337)
        q.requires_grad_(True)
        crit = ref_loss.NCELoss({"temperature": 0.05, "cluster": {"num_cluster": ncl, "num_neg_proto": r}})
        saved = ref_loss.sample
        ref_loss.sample = _FixedSample()
        try:
            loss = crit._compute_proto_loss(q, cluster_result, index)
        finally:
            ref_loss.sample = saved
        loss.backward()
        out = dict(q=q.detach().numpy(), index=index.numpy(), loss=loss.detach().numpy(), dq=q.grad.numpy(),
                   num_sets=np.int64(len(ncl)), num_neg_proto=np.int64(r))
        if compact:
            out.update(seed=np.int64(1337), num_cluster=np.array(ncl, dtype=np.int64), checksum=checksum(cluster_result))
        else:
            for s, c in enumerate(ncl):
                out[f"emb2cluster{s}"] = cluster_result["emb2cluster"][s].numpy()
                out[f"centroids{s}"] = cluster_result["centroids"][s].numpy()
                out[f"density{s}"] = cluster_result["density"][s].numpy()
        np.savez_compressed(os.path.join(HERE, f"proto_{name}.npz"), **out)


def gen_closest_docs():
    sys.modules.setdefault("pexpect", types.ModuleType("pexpect"))   # corenlp_tokenizer.py:14
    sys.path.insert(0, os.path.join(REF, "preprocessing"))
    import scipy.sparse as sp
    from drqa.retriever.tfidf_doc_ranker import TfidfDocRanker      # the reference

    rng = np.random.RandomState(1337)
    hash_size, ndocs = 64, 200
    dense = rng.rand(hash_size, ndocs) * (rng.rand(hash_size, ndocs) < 0.5)
    dense = np.round(dense * 8) / 8                                  # dyadic grid: exact in fp32 and f64, many ties
    doc_mat = sp.csr_matrix(dense)
    ranker = TfidfDocRanker.__new__(TfidfDocRanker)
    ranker.doc_mat = doc_mat
    ranker.doc_dict = ({str(i): i for i in range(ndocs)}, [str(i) for i in range(ndocs)])
    qs, ks, ids, scs = [], [], [], []
    for t in range(12):
        qv = np.zeros(hash_size)
        nz = rng.choice(hash_size, size=4, replace=False)
        qv[nz] = np.round(rng.rand(4) * 8) / 8 + 0.125
        k = int(rng.choice([1, 5, 10, 300]))
        ranker.text2spvec = lambda query, _q=qv: sp.csr_matrix(_q[None, :])
        doc_ids, doc_scores = ranker.closest_docs("ignored", k)      # tfidf_doc_ranker.py:60-75
        qs.append(qv)
        ks.append(k)
        ids.append(np.array([int(i) for i in doc_ids] + [-1] * (300 - len(doc_ids))))
        scs.append(np.concatenate([doc_scores, np.full(300 - len(doc_scores), np.nan)]))
    np.savez_compressed(os.path.join(HERE, "closest_docs.npz"), doc_mat=dense, queries=np.stack(qs),
                        k=np.array(ks), ids=np.stack(ids), scores=np.stack(scs))


def gen_pairs():
    from sklearn.metrics.pairwise import cosine_similarity          # the reference's call (:50)
    import scipy.sparse as sp

    def ref_loop(doc_tfidf):
        # preprocessing/build_docs_sentence_similarity.py:50-65, line for line
        similarity = cosine_similarity(doc_tfidf, doc_tfidf)
        sent_pair_score = []
        if doc_tfidf.shape[0] == 1:
            sent_pair_score.append(((0, 0), similarity[0][0]))
        for i in range(similarity.shape[0]):
            for j in range(i + 1, similarity.shape[0]):
                sent_pair_score.append(((i, j), similarity[i][j]))
        sent_pair_score.sort(key=lambda x: x[1], reverse=True)
        return sent_pair_score

    rng = np.random.RandomState(1337)
    out = {}
    sizes = [1, 2, 3, 7, 16, 33]
    for d, n in enumerate(sizes):
        x = rng.rand(n, 96) * (rng.rand(n, 96) < 0.2)
        if n == 7:
            x[3] = x[1]                                             # duplicate sentence -> exact ties
            x[5] = 0.0                                              # empty sentence -> zero row
        pairs = ref_loop(sp.csr_matrix(x))
        out[f"x{d}"] = x
        out[f"pairs{d}"] = np.array([[p[0][0], p[0][1]] for p in pairs], dtype=np.int64)
        out[f"scores{d}"] = np.array([p[1] for p in pairs], dtype=np.float64)
    out["ndocs"] = np.int64(len(sizes))
    np.savez_compressed(os.path.join(HERE, "pairs.npz"), **out)


def gen_paired():
    """The dense diagnostic the reference left commented out, src/evaluation.py:110-115, evaluated literally:
    `(clm_vec * evdn_vec).sum(dim=-1).mean()` on L2-normalised vectors shaped like `ctx2vec` output
    (contrastive_module.py:96-112: [B, 128], normalised)."""
    g = torch.Generator().manual_seed(1337)
    clm_vec = _unit(torch.randn(64, 128, generator=g))
    evdn_vec = _unit(clm_vec + 0.5 * torch.randn(64, 128, generator=g))
    per_pair = (clm_vec * evdn_vec).sum(dim=-1)                      # :112
    np.savez_compressed(os.path.join(HERE, "paired.npz"), clm=clm_vec.numpy(), evdn=evdn_vec.numpy(),
                        per_pair=per_pair.numpy(), mean=np.float32(per_pair.mean().item()))


if __name__ == "__main__":
    torch.manual_seed(1337)
    gen_infonce()
    gen_closest_docs()
    gen_pairs()
    gen_paired()
    print("golden vectors written to", HERE)
