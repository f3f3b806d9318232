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
* ``src.contrastor.utils.run_kmeans`` over numpy stand-ins for the faiss objects (faiss is absent): the assignment
  read-out and the concentration estimate of :67-101 are the reference's own code                -> kmeans_density.npz
* ``RetrievalModelWrapper._dequeue_and_enqueue`` / ``_momentum_update_key_encoder``
  (src/contrastor/contrastive_module.py:42-68), called unbound on a stand-in object           -> queue_maintenance.npz
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


def gen_kmeans():
    """The reference's ``run_kmeans`` (src/contrastor/utils.py:50-105) ITSELF, executed over numpy stand-ins for the
    faiss objects it builds (faiss is neither vendored nor installed): ``faiss.Clustering.train`` is the oracle's
    Lloyd iteration, ``GpuIndexFlatL2.search`` an exact float64 squared-L2 argmin.  Everything after ``clus.train`` --
    the assignment read-out, the per-cluster distance lists, the concentration estimate, its percentile clamp and
    rescale, the centroid normalisation, the tensor conversions -- is the reference's own code, and that is what the
    fixture pins.  The loader / model pair is a stand-in that hands out fixed embeddings (``extract_all_emb``, :11-25,
    stacks anchor then positive embeddings per batch)."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import kmeans as okm

    class _Index:
        def __init__(self, res, d, cfg):
            self.d, self.mat = d, np.zeros((0, d), np.float32)

        def add(self, x):
            self.mat = np.vstack([self.mat, np.asarray(x, np.float32)])

        def reset(self):
            self.mat = np.zeros((0, self.d), np.float32)

        def search(self, x, k):
            assert k == 1
            d, i = okm.assign(np.asarray(x, np.float32), self.mat)
            return d[:, None], i[:, None]

    class _Clustering:
        def __init__(self, d, k):
            self.d, self.k, self.centroids = d, k, None

        def train(self, x, index):
            c, _ = okm.lloyd(x, okm.init_centroids(x, self.k, self.seed), self.niter)
            self.centroids = c.reshape(-1)
            index.reset()
            index.add(c)

    fake = types.ModuleType("faiss")
    fake.Clustering = _Clustering
    fake.GpuIndexFlatL2 = _Index
    fake.StandardGpuResources = lambda: None
    fake.GpuIndexFlatConfig = lambda: types.SimpleNamespace(useFloat16=False, device=0)
    fake.vector_to_array = lambda v: np.asarray(v)
    sys.modules["faiss"] = fake
    sys.modules["fastcluster"] = types.ModuleType("fastcluster")
    sys.path.insert(0, REF)
    cur = torch.cuda.current_device
    torch.cuda.current_device = lambda: 0              # utils.py:59 asks for it; no GPU in the build container
    try:
        from src.contrastor import utils as ref_utils   # the reference
    finally:
        pass
    g = torch.Generator().manual_seed(1337)
    n_half, d = 300, 32
    centers = _unit(torch.randn(12, d, generator=g))
    lab = torch.randint(0, 12, (n_half,), generator=g)
    anchor = _unit(centers[lab] + 0.15 * torch.randn(n_half, d, generator=g))
    positive = _unit(centers[lab] + 0.15 * torch.randn(n_half, d, generator=g))
    anchor[7] = anchor[3]                               # duplicates: zero distances and a cluster of identical points
    batches = [(torch.arange(i, min(i + 64, n_half)), anchor[i:i + 64], positive[i:i + 64]) for i in range(0, n_half, 64)]

    class _Model:
        def bert_extract(self, a, p, device):
            return a, p

        def seq2vec(self, t):
            return t

    cfg = {"temperature": 0.05,
           "cluster": {"num_cluster": [8, 16, 150], "verbose": False, "niter": 5, "nredo": 1,
                       "max_points_per_centroid": 1000, "min_points_per_centroid": 1}}
    try:
        res = ref_utils.run_kmeans(cfg, batches, _Model(), "cpu")
    finally:
        torch.cuda.current_device = cur
    x = ref_utils.extract_all_emb(batches, _Model(), "cpu")
    out = {"x": x.astype(np.float32), "temperature": np.float32(cfg["temperature"]),
           "num_cluster": np.asarray(cfg["cluster"]["num_cluster"]), "niter": np.int64(cfg["cluster"]["niter"])}
    for s_, k in enumerate(cfg["cluster"]["num_cluster"]):
        c, _ = okm.lloyd(x, okm.init_centroids(x, k, s_), cfg["cluster"]["niter"])   # what the stand-in trained (seed = position, :58)
        out[f"raw_centroids_{s_}"] = c
        out[f"centroids_{s_}"] = res["centroids"][s_].numpy()
        out[f"density_{s_}"] = res["density"][s_].numpy()
        out[f"emb2cluster_{s_}"] = res["emb2cluster"][s_].numpy()
    np.savez_compressed(os.path.join(HERE, "kmeans_density.npz"), **out)


def gen_queue():
    """``RetrievalModelWrapper._dequeue_and_enqueue`` and ``_momentum_update_key_encoder``
    (src/contrastor/contrastive_module.py:42-68), the reference's own methods, called unbound on a stand-in object
    (the constructor downloads BERT): a sequence of key batches, one of a size that does not divide the queue (the
    reference then leaves the queue untouched, :59), wrapping around the end; two momentum steps on small encoders."""
    sys.path.insert(0, REF)
    from src.contrastor.contrastive_module import RetrievalModelWrapper as W   # the reference

    g = torch.Generator().manual_seed(1337)
    dim, size = 16, 48
    me = types.SimpleNamespace(loss_config={"queue_size": size, "momentum": 0.9}, use_LSTM=False,
                               queue=_unit(torch.randn(dim, size, generator=g), dim=0), queue_ptr=torch.zeros(1, dtype=torch.long))
    out = {"queue0": me.queue.numpy().copy(), "momentum": np.float64(0.9)}
    sizes = [12, 12, 7, 24, 12, 16]                     # 7 does not divide 48: skipped by the reference
    for i, b in enumerate(sizes):
        keys = _unit(torch.randn(b, dim, generator=g))
        W._dequeue_and_enqueue(me, keys)
        out[f"keys_{i}"] = keys.numpy()
        out[f"queue_{i}"] = me.queue.numpy().copy()
        out[f"ptr_{i}"] = np.int64(int(me.queue_ptr))
    enc_q = torch.nn.Sequential(torch.nn.Linear(8, 5), torch.nn.Linear(5, 3))
    enc_k = torch.nn.Sequential(torch.nn.Linear(8, 5), torch.nn.Linear(5, 3))
    with torch.no_grad():
        for p_ in list(enc_q.parameters()) + list(enc_k.parameters()):
            p_.copy_(torch.randn(p_.shape, generator=g))
    me.encoder_q, me.encoder_k = enc_q, enc_k
    for j, p_ in enumerate(enc_q.parameters()):
        out[f"pq_{j}"] = p_.detach().numpy().copy()
    for j, p_ in enumerate(enc_k.parameters()):
        out[f"pk_{j}"] = p_.detach().numpy().copy()
    for step in range(2):
        W._momentum_update_key_encoder(me)
        for j, p_ in enumerate(enc_k.parameters()):
            out[f"pk_{j}_after{step}"] = p_.detach().numpy().copy()
    np.savez_compressed(os.path.join(HERE, "queue_maintenance.npz"), **out)


if __name__ == "__main__":
    torch.manual_seed(1337)
    gen_queue()
    gen_kmeans()
    gen_infonce()
    gen_closest_docs()
    gen_pairs()
    gen_paired()
    print("golden vectors written to", HERE)
