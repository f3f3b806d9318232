"""Dense claim x corpus retrieval: scores and top-k ids.

Host-side mirror of the reference's score-and-select interface.  The reference's dense call
site is the commented-out block at src/evaluation.py:105-116 (``ctx2vec`` embeddings, dot
products); its live score+select signature is ``TfidfDocRanker.closest_docs(query, k) ->
(doc_ids, doc_scores)`` and ``batch_closest_docs``
(preprocessing/drqa/retriever/tfidf_doc_ranker.py:60-84).  ``DenseIndex`` keeps those names,
argument meaning and return shapes; ``search`` is the tensor-level call underneath.

All arithmetic runs in the CUDA extension (csrc/): a tcgen05 score GEMM -- bf16 / fp16 operands, or fp32
operands as a 3 x TF32 split for exact comparison (an FFMA kernel is its checker) -- with the top-k select
fused into its epilogue.  There is no CPU path.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib

_DTYPES = {torch.float32: _lib.DRS_F32, torch.bfloat16: _lib.DRS_BF16, torch.float16: _lib.DRS_F16}

MAX_CLAIMS_PER_PASS = 1 << 18   # claims scored per engine call by `search` (larger batches are sliced)

# one cached workspace per (device, stream) so steady-state searches allocate nothing
_workspaces: dict = {}


def _workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def _check_matrix(name: str, t: torch.Tensor):
    if not isinstance(t, torch.Tensor) or t.dim() != 2:
        raise ValueError(f"{name} must be a 2-D tensor, got {type(t).__name__} with shape {getattr(t, 'shape', None)}")
    if t.dtype not in _DTYPES:
        raise TypeError(f"{name} must be float32, bfloat16 or float16, got {t.dtype}")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must live on a CUDA device: drs_b200 has no CPU path")


def search(queries: torch.Tensor, corpus: torch.Tensor, k: int, *, id_base: int = 0, profile: Optional[list] = None):
    """Top-k corpus rows per query by dot product.

    queries [nq, D], corpus [Nc, D]: CUDA, row-major, same dtype (bf16 or fp16 -> tcgen05 path, fp32
    accumulation, scores within 2e-2 relative of fp32; fp32 -> 3 x TF32 on tcgen05 / FFMA kernel, 1e-5).  Rows are expected to be
    L2-normalised the way ``seq2vec`` does (src/contrastor/contrastive_module.py:111), so the dot
    product is the cosine; nothing here depends on that.

    Returns (scores fp32 [nq, k'], ids int64 [nq, k']) with k' = min(k, Nc) (like closest_docs,
    tfidf_doc_ranker.py:67-68, fewer rows than k returns them all), sorted by score descending,
    ties broken by the lower row index; ids = row index + id_base.

    ``profile``: optional list; a (start, end) pair of CUDA events bracketing the scan kernel (the
    fused score GEMM + top-k) on the current stream is appended to it, for roofline measurement.
    """
    _check_matrix("queries", queries)
    _check_matrix("corpus", corpus)
    if queries.device != corpus.device:
        raise RuntimeError(f"queries ({queries.device}) and corpus ({corpus.device}) must be on the same device")
    if queries.shape[1] != corpus.shape[1]:
        raise ValueError(f"dimension mismatch: queries {tuple(queries.shape)} vs corpus {tuple(corpus.shape)}")
    if queries.dtype != corpus.dtype:
        queries = queries.to(corpus.dtype)
    if k <= 0:
        raise ValueError(f"k must be positive, got {k}")
    nq, dim = queries.shape
    nc = corpus.shape[0]
    kk = min(int(k), nc)
    dev = queries.device
    if nq == 0 or nc == 0:
        return (torch.empty(nq, kk, dtype=torch.float32, device=dev), torch.empty(nq, kk, dtype=torch.int64, device=dev))
    if kk > _lib.DRS_MAX_K:
        raise RuntimeError(f"k={kk} exceeds the engine limit of {_lib.DRS_MAX_K}")
    queries = queries.contiguous()
    corpus = corpus.contiguous()
    if nq > MAX_CLAIMS_PER_PASS:
        # the candidate workspace grows with the claim count (~10 KB per claim): very large batches go through in
        # slices (the scan is tensor-bound at these sizes, so re-streaming the corpus per slice costs nothing)
        parts = [search(queries[a:a + MAX_CLAIMS_PER_PASS], corpus, k, id_base=id_base)
                 for a in range(0, nq, MAX_CLAIMS_PER_PASS)]
        return torch.cat([p[0] for p in parts]), torch.cat([p[1] for p in parts])
    lib = _lib.load()
    dt = _DTYPES[corpus.dtype]
    with torch.cuda.device(dev):
        need = ctypes.c_size_t(0)
        _lib.check(lib.drs_search_workspace_bytes(nq, nc, dim, kk, dt, ctypes.byref(need)))
        ws = _workspace(dev, need.value)
        scores = torch.empty(nq, kk, dtype=torch.float32, device=dev)
        ids = torch.empty(nq, kk, dtype=torch.int64, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        if profile is None or kk > 16:      # the scan/select split serves single-pass searches
            _lib.check(lib.drs_search(queries.data_ptr(), nq, corpus.data_ptr(), nc, dim, dt, kk, int(id_base),
                                      scores.data_ptr(), ids.data_ptr(), ws.data_ptr(), ws.numel(), stream))
        else:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            _lib.check(lib.drs_search_scan(queries.data_ptr(), nq, corpus.data_ptr(), nc, dim, dt, kk,
                                           ws.data_ptr(), ws.numel(), stream))
            ev1.record()
            _lib.check(lib.drs_search_select(ws.data_ptr(), nq, nc, dim, dt, kk, int(id_base), scores.data_ptr(),
                                             ids.data_ptr(), stream))
            profile.append((ev0, ev1))
    return scores, ids


def open_claims_per_pass(device=None):
    """Debug: claims each pass of the last k > 32 search on this device/stream left open (a list of 8
    counters; all zero = the first pass finished every claim)."""
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    ws = _workspaces.get((dev.index, torch.cuda.current_stream(dev).cuda_stream))
    if ws is None:
        raise RuntimeError("no search has run on this device/stream yet")
    out = (ctypes.c_uint * 8)()
    with torch.cuda.device(dev):
        _lib.check(_lib.load().drs_debug_open_claims(ws.data_ptr(), ctypes.byref(out),
                                                     torch.cuda.current_stream(dev).cuda_stream))
    return list(out)


def flat_l2_search(x: torch.Tensor, centroids: torch.Tensor, k: int = 1, *, id_base: int = 0):
    """Exact squared-L2 nearest neighbours: what ``faiss.GpuIndexFlatL2.search(x, k)`` returns at
    src/contrastor/utils.py:64-67 (the k-means assignment of ``run_kmeans``).

    x [n, D], centroids [Nc, D]: CUDA, same dtype (fp32 -> the exact path (3 x TF32 on tcgen05) like faiss's
    ``useFloat16 = False`` at utils.py:44; bf16 -> tcgen05 path).  Returns (D fp32 [n, k'] ascending
    squared distances, I int64 [n, k']), k' = min(k, Nc), ties -> lower index.
    PARITY UNPINNED against faiss itself (unpinned, not vendored, not installed)."""
    _check_matrix("x", x)
    _check_matrix("centroids", centroids)
    if x.device != centroids.device:
        raise RuntimeError("x and centroids must be on the same device")
    if x.shape[1] != centroids.shape[1]:
        raise ValueError(f"dimension mismatch: x {tuple(x.shape)} vs centroids {tuple(centroids.shape)}")
    if x.dtype != centroids.dtype:
        x = x.to(centroids.dtype)
    if k <= 0:
        raise ValueError(f"k must be positive, got {k}")
    n, dim = x.shape
    nc = centroids.shape[0]
    kk = min(int(k), nc)
    dev = x.device
    if n == 0 or nc == 0:
        return (torch.empty(n, kk, dtype=torch.float32, device=dev), torch.empty(n, kk, dtype=torch.int64, device=dev))
    if kk > _lib.DRS_MAX_K:
        raise RuntimeError(f"k={kk} exceeds the engine limit of {_lib.DRS_MAX_K}")
    x = x.contiguous()
    centroids = centroids.contiguous()
    lib = _lib.load()
    dt = _DTYPES[centroids.dtype]
    with torch.cuda.device(dev):
        need = ctypes.c_size_t(0)
        _lib.check(lib.drs_search_l2_workspace_bytes(n, nc, dim, kk, dt, ctypes.byref(need)))
        ws = _workspace(dev, need.value)
        dist = torch.empty(n, kk, dtype=torch.float32, device=dev)
        ids = torch.empty(n, kk, dtype=torch.int64, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.drs_search_l2(x.data_ptr(), n, centroids.data_ptr(), nc, dim, dt, kk, int(id_base),
                                     dist.data_ptr(), ids.data_ptr(), ws.data_ptr(), ws.numel(), stream))
    return dist, ids


class FlatL2Index:
    """The slice of ``faiss.GpuIndexFlatL2`` that ``run_kmeans`` uses (src/contrastor/utils.py:39-47,
    :64-67): ``add`` vectors, ``search(x, k) -> (D, I)`` as numpy arrays, ``ntotal``, ``reset``."""

    def __init__(self, d: int, device=None, dtype: torch.dtype = torch.float32):
        self.d = int(d)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.dtype = dtype                      # fp32 == faiss cfg.useFloat16 = False (utils.py:44)
        self._chunks = []
        self._mat = None

    @property
    def ntotal(self) -> int:
        return sum(c.shape[0] for c in self._chunks)

    def reset(self):
        self._chunks, self._mat = [], None

    def add(self, x):
        x = torch.as_tensor(x)
        if x.dim() != 2 or x.shape[1] != self.d:
            raise ValueError(f"add expects [n, {self.d}] vectors")
        self._chunks.append(x.to(device=self.device, dtype=self.dtype).contiguous())
        self._mat = None

    def search(self, x, k: int = 1):
        if not self._chunks:
            raise RuntimeError("search on an empty index")
        if self._mat is None:
            self._mat = self._chunks[0] if len(self._chunks) == 1 else torch.cat(self._chunks)
            self._chunks = [self._mat]
        d, i = flat_l2_search(torch.as_tensor(x).to(device=self.device, dtype=self.dtype), self._mat, k)
        return d.cpu().numpy(), i.cpu().numpy()


def rerank(queries: torch.Tensor, corpus: torch.Tensor, cand_ids: torch.Tensor, k: int):
    """Score each claim against its OWN candidate rows and keep the best k -- the dense stage of the
    report's "TF-IDF top-100 -> contrastive re-rank -> top-15" (report.pdf section 3.2), to be placed at
    src/evaluation.py:105-116 after the sparse candidates of documents_filtering (:57-83).

    queries [nq, D], corpus [Nc, D] (CUDA, same dtype: bf16 or fp32); cand_ids [nq, m] integer corpus
    rows, negative (or >= Nc) entries are padding.  Returns (scores fp32 [nq, k'], ids int64 [nq, k'])
    with k' = min(k, m), sorted descending, ties -> lower row, a row listed twice reported once,
    (-inf, -1) where a claim has fewer than k' valid candidates."""
    _check_matrix("queries", queries)
    _check_matrix("corpus", corpus)
    if queries.device != corpus.device:
        raise RuntimeError(f"queries ({queries.device}) and corpus ({corpus.device}) must be on the same device")
    if queries.shape[1] != corpus.shape[1]:
        raise ValueError(f"dimension mismatch: queries {tuple(queries.shape)} vs corpus {tuple(corpus.shape)}")
    if not isinstance(cand_ids, torch.Tensor) or cand_ids.dim() != 2 or cand_ids.shape[0] != queries.shape[0]:
        raise ValueError("cand_ids must be a [nq, m] integer tensor")
    if cand_ids.dtype.is_floating_point:
        raise TypeError("cand_ids must hold integer row numbers")
    if k <= 0:
        raise ValueError(f"k must be positive, got {k}")
    if queries.dtype != corpus.dtype:
        queries = queries.to(corpus.dtype)
    nq, dim = queries.shape
    m = cand_ids.shape[1]
    kk = min(int(k), m)
    dev = queries.device
    scores = torch.empty(nq, kk, dtype=torch.float32, device=dev)
    ids = torch.empty(nq, kk, dtype=torch.int64, device=dev)
    if nq == 0 or kk == 0:
        return scores, ids
    if corpus.shape[0] == 0:
        return scores.fill_(float("-inf")), ids.fill_(-1)
    queries = queries.contiguous()
    corpus = corpus.contiguous()
    cand = cand_ids.to(device=dev, dtype=torch.int64).contiguous()
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(_lib.load().drs_rerank(queries.data_ptr(), nq, corpus.data_ptr(), corpus.shape[0], dim,
                                          _DTYPES[corpus.dtype], cand.data_ptr(), m, kk, scores.data_ptr(),
                                          ids.data_ptr(), stream))
    return scores, ids


def paired_scores(a: torch.Tensor, b: torch.Tensor):
    """``(clm_vec * evdn_vec).sum(dim=-1)`` of the commented dense evaluation (src/evaluation.py:112,115):
    row-wise dot products of two [n, D] CUDA matrices -> fp32 [n] (the caller takes ``.mean()``)."""
    _check_matrix("a", a)
    _check_matrix("b", b)
    if a.shape != b.shape:
        raise ValueError(f"shape mismatch: {tuple(a.shape)} vs {tuple(b.shape)}")
    if a.device != b.device:
        raise RuntimeError("a and b must be on the same device")
    if a.dtype != b.dtype:
        a, b = a.float(), b.float()
    n, dim = a.shape
    out = torch.empty(n, dtype=torch.float32, device=a.device)
    if n == 0:
        return out
    if dim == 0:
        return out.zero_()
    a, b = a.contiguous(), b.contiguous()
    with torch.cuda.device(a.device):
        stream = torch.cuda.current_stream(a.device).cuda_stream
        _lib.check(_lib.load().drs_pair_scores(a.data_ptr(), b.data_ptr(), n, dim, _DTYPES[a.dtype], out.data_ptr(), stream))
    return out


def merge_shards(scores: torch.Tensor, ids: torch.Tensor):
    """[g, nq, k] per-shard lists (id < 0 = empty) -> [nq, k] by (score desc, id asc)."""
    if scores.dim() != 3 or scores.shape != ids.shape:
        raise ValueError("scores and ids must both be [num_shards, nq, k]")
    if not scores.is_cuda:
        raise RuntimeError("merge_shards needs CUDA tensors: drs_b200 has no CPU path")
    g, nq, k = scores.shape
    scores = scores.contiguous().float()
    ids = ids.contiguous().long()
    out_s = torch.empty(nq, k, dtype=torch.float32, device=scores.device)
    out_i = torch.empty(nq, k, dtype=torch.int64, device=scores.device)
    if nq == 0 or k == 0:
        return out_s, out_i
    with torch.cuda.device(scores.device):
        stream = torch.cuda.current_stream(scores.device).cuda_stream
        _lib.check(_lib.load().drs_merge_shards(scores.data_ptr(), ids.data_ptr(), g, nq, k, out_s.data_ptr(),
                                                out_i.data_ptr(), stream))
    return out_s, out_i


def shard_bounds(num_rows: int, rank: int, world_size: int):
    """contiguous row shards, rank r owns [r*ceil(N/g), min(N, (r+1)*ceil(N/g)))."""
    per = -(-num_rows // world_size)
    lo = min(num_rows, rank * per)
    return lo, min(num_rows, lo + per)


class DenseIndex:
    """A resident corpus of embeddings with the ``TfidfDocRanker`` query surface.

    embeddings : [Nc, D] tensor (moved to ``device``; stored as ``dtype``, default bf16)
    doc_ids    : optional list mapping row -> doc id, the ``doc_dict[1]`` of the reference
                 (tfidf_doc_ranker.py:52-58); defaults to the row number
    id_base    : global row number of this shard's first row (row-sharded corpora)
    """

    def __init__(self, embeddings: torch.Tensor, doc_ids: Optional[Sequence] = None, *, device=None,
                 dtype: torch.dtype = torch.bfloat16, id_base: int = 0):
        if device is None:
            device = embeddings.device if embeddings.is_cuda else torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        self.embeddings = embeddings.to(device=self.device, dtype=dtype).contiguous()
        self.id_base = int(id_base)
        self.num_docs = self.embeddings.shape[0]
        if doc_ids is not None and len(doc_ids) != self.num_docs:
            raise ValueError("doc_ids must have one entry per corpus row")
        self.doc_dict = None
        if doc_ids is not None:
            self.doc_dict = ({d: i for i, d in enumerate(doc_ids)}, list(doc_ids))

    # -- tfidf_doc_ranker.py:52-58.  Rows are GLOBAL on both sides (a shard's first row is `id_base`), so
    #    get_doc_index(get_doc_id(r)) == r for every id `search` returns, sharded or not.
    def get_doc_index(self, doc_id):
        return (self.doc_dict[0][doc_id] + self.id_base) if self.doc_dict else int(doc_id)

    def get_doc_id(self, doc_index):
        local = int(doc_index) - self.id_base
        if not 0 <= local < self.num_docs:
            raise IndexError(f"row {doc_index} is not in this shard [{self.id_base}, {self.id_base + self.num_docs})")
        return self.doc_dict[1][local] if self.doc_dict else int(doc_index)

    def search(self, queries: torch.Tensor, k: int = 1, profile: Optional[list] = None):
        """queries [nq, D] on any device (host tensors are copied in) -> device (scores, ids)."""
        q = queries.to(device=self.device, dtype=self.embeddings.dtype, non_blocking=True)
        return search(q, self.embeddings, k, id_base=self.id_base, profile=profile)

    def rerank(self, queries: torch.Tensor, cand_ids, k: int = 15):
        """Dense re-rank of per-claim candidate rows (report.pdf section 3.2: sparse top-100 -> dense top-15).
        ``cand_ids`` [nq, m]: GLOBAL row numbers (this shard's ``id_base`` is subtracted; rows outside the
        shard and negative entries are ignored).  Returns device (scores, global ids)."""
        q = queries.to(device=self.device, dtype=self.embeddings.dtype, non_blocking=True)
        cand = torch.as_tensor(cand_ids).to(device=self.device, dtype=torch.int64)
        if self.id_base:
            cand = torch.where(cand >= 0, cand - self.id_base, cand)
        s, i = rerank(q, self.embeddings, cand, k)
        if self.id_base:
            i = torch.where(i >= 0, i + self.id_base, i)
        return s, i

    def save(self, filename: str, metadata: Optional[dict] = None):
        """Write this corpus to an index file (store.save_dense_index); the doc-id map travels as
        ``metadata['doc_dict']`` like the reference's (retriever/utils.py:21-29)."""
        from .store import save_dense_index
        meta = dict(metadata or {})
        if self.doc_dict is not None:
            meta.setdefault("doc_dict", self.doc_dict)
        save_dense_index(filename, self.embeddings, meta, dtype=self.embeddings.dtype)

    def closest_docs(self, query: torch.Tensor, k: int = 1):
        """tfidf_doc_ranker.py:60-75 -- one query vector [D] -> (list of doc ids, np.ndarray scores)."""
        res = self.batch_closest_docs(query.reshape(1, -1), k)
        return res[0]

    def batch_closest_docs(self, queries: torch.Tensor, k: int = 1, num_workers=None):
        """tfidf_doc_ranker.py:77-84 -- a batch of query vectors [nq, D] -> list of
        (doc_ids, doc_scores).  ``num_workers`` is accepted and ignored (one GPU launch)."""
        scores, ids = self.search(queries, k)
        scores = scores.cpu().numpy().astype(np.float64)      # closest_docs returns f64 scores
        ids = ids.cpu().numpy()
        out = []
        for s_row, i_row in zip(scores, ids):
            keep = i_row >= 0
            out.append(([self.get_doc_id(int(i)) for i in i_row[keep]], s_row[keep]))
        return out


class DenseDocRanker(DenseIndex):
    """``TfidfDocRanker`` with dense vectors: built from an index FILE (tfidf_doc_ranker.py:33-50 loads
    ``tfidf_path``; here ``store.save_dense_index`` wrote it) and queried with TEXT -- ``encoder`` turns a
    list of strings into [n, D] embeddings, i.e. ``lambda texts: model.ctx2vec(texts, device)``
    (src/contrastor/contrastive_module.py:96-100), the call the commented block at
    src/evaluation.py:110-111 makes.  Tensors are accepted too (then no encoder is needed)."""

    def __init__(self, index_path: str, encoder=None, *, device=None, rank: int = 0, world_size: int = 1, strict: bool = True,
                 verify: bool = False):
        from .store import load_dense_index
        index, meta = load_dense_index(index_path, device=device, rank=rank, world_size=world_size, verify=verify)
        self.__dict__.update(index.__dict__)
        self.metadata = meta
        self.encoder = encoder
        self.strict = strict

    def text2vec(self, queries):
        if self.encoder is None:
            raise RuntimeError("DenseDocRanker was built without an encoder: pass embeddings, or encoder=model.ctx2vec")
        with torch.no_grad():
            return self.encoder(list(queries))

    def closest_docs(self, query, k: int = 1):
        if isinstance(query, str):
            query = self.text2vec([query])
        return super().closest_docs(query, k)

    def batch_closest_docs(self, queries, k: int = 1, num_workers=None):
        if not isinstance(queries, torch.Tensor):
            queries = self.text2vec(queries)
        return super().batch_closest_docs(queries, k, num_workers)


class _PeerBuffer:
    """One zero-filled device buffer per rank, mapped into every other rank of the group over CUDA IPC
    (csrc/exchange_api.inc::drs_peer_*): ``ptrs[r]`` is rank r's buffer as seen from this process."""

    def __init__(self, group, device, world, rank, nbytes):
        import torch.distributed as dist
        lib = _lib.load()
        self.device, self.rank, self.world, self.group = device, rank, world, group
        self.ptrs, self._own, self._opened = [], None, []
        with torch.cuda.device(device):
            ptr, handle = ctypes.c_void_p(), (ctypes.c_ubyte * 64)()
            _lib.check(lib.drs_peer_alloc(nbytes, ctypes.byref(ptr), ctypes.byref(handle)))
            self._own = ptr.value
            mine = torch.tensor(list(handle), dtype=torch.uint8, device=device)
            everyone = torch.empty(world * 64, dtype=torch.uint8, device=device)
            dist.all_gather_into_tensor(everyone, mine, group=group)
            handles = everyone.cpu().view(world, 64)
            err = None
            for r in range(world):
                if r == rank:
                    self.ptrs.append(self._own)
                    continue
                h = (ctypes.c_ubyte * 64)(*handles[r].tolist())
                p = ctypes.c_void_p()
                try:
                    _lib.check(lib.drs_peer_open(ctypes.byref(h), ctypes.byref(p)))
                except RuntimeError as e:             # keep the collectives below balanced, fail afterwards
                    err = err or e
                    self.ptrs.append(0)
                    continue
                self._opened.append(p.value)
                self.ptrs.append(p.value)
            # doubles as the barrier: every rank's buffer is zeroed and mapped before anyone publishes
            ok = torch.tensor([0 if err else 1], device=device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if not int(ok.item()):
                self.close()
                raise RuntimeError(f"peer buffers could not be mapped on every rank ({err or 'a peer failed'})")

    def close(self):
        """Collective: call on every rank, after the last exchange that used this buffer."""
        import torch.distributed as dist
        if self._own is None:
            return
        lib = _lib.load()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)            # no peer is still reading or writing
            for p in self._opened:
                lib.drs_peer_close(p)
            dist.barrier(group=self.group)            # every mapping is closed before the owner frees
            lib.drs_peer_free(self._own)
        self._own, self._opened, self.ptrs = None, [], []


class _PeerExchange:
    """Peer-mapped gather buffers and flags for the fused select + exchange + merge kernel
    (csrc/exchange.cuh): every rank allocates the same layout and maps every peer's copy over NVLink.

        [ flags: claim blocks x world u32 | parity 0: scores, ids | parity 1: scores, ids ]
    """

    MAX_NQ = 1 << 17

    def __init__(self, group, device, world, rank, capacity):
        self.world, self.rank, self.capacity = world, rank, int(capacity)         # capacity: nq * k entries
        need = ctypes.c_size_t(0)
        _lib.check(_lib.load().drs_exchange_flag_bytes(self.MAX_NQ, world, ctypes.byref(need)))
        self.flag_bytes = need.value
        self.s_bytes = (world * self.capacity * 4 + 255) // 256 * 256
        self.i_bytes = (world * self.capacity * 8 + 255) // 256 * 256
        total = self.flag_bytes + 2 * (self.s_bytes + self.i_bytes)
        self.buf = _PeerBuffer(group, device, world, rank, total)
        self.ptrs = self.buf.ptrs
        self.calls = torch.zeros(1, dtype=torch.int32, device=device)      # this rank's completed exchange calls
        arr = ctypes.c_void_p * world
        base = self.flag_bytes                                             # parity-0 buffers; parity 1 = + parity_stride
        self.parity_stride = self.s_bytes + self.i_bytes
        self.score_ptrs = arr(*[p + base for p in self.ptrs])
        self.id_ptrs = arr(*[p + base + self.s_bytes for p in self.ptrs])
        self.flag_ptrs = arr(*self.ptrs)

    def close(self):
        self.buf.close()


class _SlicedExchange:
    """Peer-mapped buffer for the query-sliced exchange (csrc/exchange.cuh::exchange_sliced_kernel): per rank

        [ flags1 | flags2 | parity 0: gather scores, gather ids, result scores, result ids | parity 1: ... ]

    laid out by the library (drs_exchange_sliced_bytes) from (max_nq, max_entries, world)."""

    MAX_NQ = 1 << 18

    def __init__(self, group, device, world, rank, max_entries):
        self.world, self.rank, self.max_entries = world, rank, int(max_entries)
        need = ctypes.c_size_t(0)
        _lib.check(_lib.load().drs_exchange_sliced_bytes(self.MAX_NQ, self.max_entries, world, ctypes.byref(need)))
        self.buf = _PeerBuffer(group, device, world, rank, need.value)
        self.bases = (ctypes.c_void_p * world)(*self.buf.ptrs)
        self.calls = torch.zeros(1, dtype=torch.int32, device=device)

    def close(self):
        self.buf.close()


class ShardedDenseIndex:
    """Row-sharded corpus over the ranks of a torch.distributed group (one process per GPU).

    Each rank holds rows ``shard_bounds(N, rank, world)`` and runs the same fused scan kernel on them.
    The per-rank (score, global id) lists are then exchanged and merged by (score desc, id asc), which
    makes the result identical to a single-GPU search of the whole corpus.  Two exchanges:

    * ``exchange='p2p'`` (NVLink peer memory, csrc/exchange.cuh; buffers mapped across the ranks with CUDA IPC, so
      one process per GPU of ONE box; needs a CUDA-capable backend (nccl) for the handle exchange, <= 8 ranks, no
      empty shard).  k <= 16: ONE kernel per search selects the shard's top-k, stores it into every peer's
      buffer, waits on per-claim-block flags and merges.  Larger k (BASELINE configs[4]: top-100): the QUERY-SLICED
      kernel -- rank s receives only the lists of its slice of the claims, merges them and stores the final lists
      into every rank's result buffer (world x fewer bytes and merges per GPU than gathering everything everywhere);
    * ``exchange='nccl'``: select, ``all_gather_into_tensor`` of scores and ids, merge kernel (any backend).

    ``exchange='auto'`` (default) takes 'p2p' when its conditions hold, else 'nccl'.
    """

    def __init__(self, local_embeddings: torch.Tensor, total_rows: int, *, group=None, device=None,
                 dtype: torch.dtype = torch.bfloat16, exchange: str = "auto"):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        lo, hi = shard_bounds(total_rows, self.rank, self.world)
        if local_embeddings.shape[0] != hi - lo:
            raise ValueError(f"rank {self.rank} must hold rows [{lo}, {hi}) = {hi - lo} rows, got {local_embeddings.shape[0]}")
        if exchange not in ("auto", "p2p", "nccl"):
            raise ValueError("exchange must be 'auto', 'p2p' or 'nccl'")
        self.total_rows = total_rows
        self.local = DenseIndex(local_embeddings, device=device, dtype=dtype, id_base=lo)
        # the same decision on every rank: it depends only on group-wide facts
        smallest = min(b - a for a, b in (shard_bounds(total_rows, r, self.world) for r in range(self.world)))
        p2p_ok = (self.world > 1 and self.world <= 8 and smallest > 0 and self.local.embeddings.is_cuda
                  and dist.get_backend(group) == "nccl")
        if exchange == "p2p" and not p2p_ok:
            raise RuntimeError("exchange='p2p' needs the nccl backend, 2..8 ranks, CUDA shards and no empty shard")
        self.exchange = "p2p" if (p2p_ok and exchange != "nccl") else "nccl"
        self._exchange_requested = exchange
        self._peer = None
        self._sliced = None

    def _peer_exchange(self, entries: int):
        """The symmetric buffers, (re)allocated collectively -- every rank sees the same `entries`.  Returns
        None (and switches this index to the NCCL exchange on ALL ranks) if any rank fails to map its peers."""
        if self._peer is None or self._peer.capacity < entries:
            cap = max(entries, 1 << 17) if self._peer is None else max(entries, 2 * self._peer.capacity)
            ok, err = 1, None
            try:
                peer = _PeerExchange(self.group, self.local.device, self.world, self.rank, cap)
            except Exception as e:  # noqa: BLE001  (no peer access, symmetric memory unavailable, ...)
                ok, err, peer = 0, e, None
            flag = torch.tensor([ok], device=self.local.device)
            self.dist.all_reduce(flag, op=self.dist.ReduceOp.MIN, group=self.group)
            if not int(flag.item()):
                if self._exchange_requested == "p2p":
                    raise RuntimeError(f"exchange='p2p': peer memory could not be mapped on every rank ({err})")
                self.exchange, self._peer = "nccl", None
                return None
            if self._peer is not None:
                self._peer.close()                    # collective: every rank regrows at the same call
            self._peer = peer
        return self._peer

    def _sliced_exchange(self, entries: int):
        """The query-sliced exchange buffers, (re)allocated collectively like `_peer_exchange`."""
        if self._sliced is None or self._sliced.max_entries < entries:
            cap = max(entries, 1 << 20) if self._sliced is None else max(entries, 2 * self._sliced.max_entries)
            ok, err = 1, None
            try:
                sl = _SlicedExchange(self.group, self.local.device, self.world, self.rank, cap)
            except Exception as e:  # noqa: BLE001
                ok, err, sl = 0, e, None
            flag = torch.tensor([ok], device=self.local.device)
            self.dist.all_reduce(flag, op=self.dist.ReduceOp.MIN, group=self.group)
            if not int(flag.item()):
                if self._exchange_requested == "p2p":
                    raise RuntimeError(f"exchange='p2p': peer memory could not be mapped on every rank ({err})")
                self.exchange, self._sliced = "nccl", None
                return None
            if self._sliced is not None:
                self._sliced.close()
            self._sliced = sl
        return self._sliced

    def close(self):
        """Release the peer-mapped exchange buffers (collective; optional -- process exit releases them too)."""
        for ex in (self._peer, self._sliced):
            if ex is not None:
                ex.close()
        self._peer = self._sliced = None

    def search(self, queries: torch.Tensor, k: int = 1, profile: Optional[list] = None, timing: Optional[dict] = None):
        """``timing``: optional dict; CUDA event pairs are stored under 'local' (scan + select of this shard) and
        'exchange' (the exchange + merge) for the paths that run them as separate launches (k > 16, NCCL)."""
        kk = min(int(k), self.total_rows)
        nq = queries.shape[0]
        if self.exchange == "p2p" and 0 < kk <= 16 and 0 < nq <= _PeerExchange.MAX_NQ:
            if self._peer_exchange(nq * kk) is not None:
                return self._search_p2p(queries, kk, profile)
        dev = self.local.device

        def mark():
            if timing is None:
                return None
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(torch.cuda.current_stream(dev))
            return ev

        t0 = mark()
        slot_s, slot_i = self.local_lists(queries, kk, profile)
        t1 = mark()
        out = self.exchange_lists(slot_s, slot_i)
        t2 = mark()
        if timing is not None:
            timing.setdefault("local", []).append((t0, t1))
            timing.setdefault("exchange", []).append((t1, t2))
        return out

    def local_lists(self, queries: torch.Tensor, kk: int, profile: Optional[list] = None):
        """This shard's sorted (score, GLOBAL id) lists, [nq, kk] each, padded with (-inf, -1) when the shard holds
        fewer than kk rows -- fixed-size slots, so every rank contributes the same number of bytes."""
        nq = queries.shape[0]
        dev = self.local.device
        s, i = (self.local.search(queries, min(kk, max(self.local.num_docs, 1)), profile=profile)
                if self.local.num_docs else (None, None))
        if s is not None and s.shape[1] == kk:
            return s, i
        slot_s = torch.full((nq, kk), float("-inf"), dtype=torch.float32, device=dev)
        slot_i = torch.full((nq, kk), -1, dtype=torch.int64, device=dev)
        if s is not None:
            slot_s[:, : s.shape[1]] = s
            slot_i[:, : i.shape[1]] = i
        return slot_s, slot_i

    def exchange_lists(self, slot_s: torch.Tensor, slot_i: torch.Tensor):
        """The one exchange step on per-shard lists (a collective): query-sliced peer-memory kernel, or NCCL
        all-gather + merge kernel.  Returns the merged [nq, kk] (scores, ids), identical on every rank."""
        nq, kk = slot_s.shape
        dev = self.local.device
        if self.world == 1:
            return slot_s, slot_i
        per = -(-nq // self.world)
        if self.exchange == "p2p" and 0 < nq <= _SlicedExchange.MAX_NQ and kk <= _lib.DRS_MAX_K:
            sl = self._sliced_exchange(per * self.world * kk)
            if sl is not None:
                out_s = torch.empty(nq, kk, dtype=torch.float32, device=dev)
                out_i = torch.empty(nq, kk, dtype=torch.int64, device=dev)
                with torch.cuda.device(dev):
                    _lib.check(_lib.load().drs_exchange_sliced(slot_s.data_ptr(), slot_i.data_ptr(), nq, kk, self.rank, self.world,
                                                               sl.bases, sl.MAX_NQ, sl.max_entries, sl.calls.data_ptr(),
                                                               out_s.data_ptr(), out_i.data_ptr(),
                                                               torch.cuda.current_stream(dev).cuda_stream))
                return out_s, out_i
        all_s, all_i = all_gather_topk(slot_s.contiguous(), slot_i.contiguous(), self.group)
        return merge_shards(all_s, all_i)

    def _search_p2p(self, queries: torch.Tensor, kk: int, profile: Optional[list]):
        loc = self.local
        dev = loc.device
        q = queries.to(device=dev, dtype=loc.embeddings.dtype, non_blocking=True).contiguous()
        _check_matrix("queries", q)
        if q.shape[1] != loc.embeddings.shape[1]:
            raise ValueError(f"dimension mismatch: queries {tuple(q.shape)} vs corpus {tuple(loc.embeddings.shape)}")
        nq, dim = q.shape
        nc = loc.num_docs
        peer = self._peer
        lib = _lib.load()
        dt = _DTYPES[loc.embeddings.dtype]
        with torch.cuda.device(dev):
            need = ctypes.c_size_t(0)
            _lib.check(lib.drs_search_workspace_bytes(nq, nc, dim, kk, dt, ctypes.byref(need)))
            ws = _workspace(dev, need.value)
            scores = torch.empty(nq, kk, dtype=torch.float32, device=dev)
            ids = torch.empty(nq, kk, dtype=torch.int64, device=dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            ev0 = ev1 = None
            if profile is not None:
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
            _lib.check(lib.drs_search_sharded_p2p(q.data_ptr(), nq, loc.embeddings.data_ptr(), nc, dim, dt, kk, loc.id_base,
                                                  self.rank, self.world, peer.score_ptrs, peer.id_ptrs, peer.flag_ptrs,
                                                  peer.parity_stride, peer.calls.data_ptr(), scores.data_ptr(),
                                                  ids.data_ptr(), ws.data_ptr(), ws.numel(), stream))
            if profile is not None:
                ev1.record()
                profile.append((ev0, ev1))
        return scores, ids


def all_gather_topk(slot_s: torch.Tensor, slot_i: torch.Tensor, group=None):
    """The one exchange step of the sharded path: every rank contributes its [nq, k] scores and
    global ids; returns the stacked [world, nq, k] tensors (same on every rank)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    nq, k = slot_s.shape
    # rank-major concatenation along dim 0 (the layout every backend accepts), viewed as [world, nq, k]
    all_s = torch.empty(world * nq, k, dtype=slot_s.dtype, device=slot_s.device)
    all_i = torch.empty(world * nq, k, dtype=slot_i.dtype, device=slot_i.device)
    dist.all_gather_into_tensor(all_s, slot_s.contiguous(), group=group)
    dist.all_gather_into_tensor(all_i, slot_i.contiguous(), group=group)
    return all_s.view(world, nq, k), all_i.view(world, nq, k)
