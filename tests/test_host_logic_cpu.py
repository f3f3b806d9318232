"""Host-side logic that needs no GPU: shard planning, argument validation (the product path must
refuse CPU tensors instead of falling back), the reference-shaped module surface, and the
world_size-2 exchange step over gloo."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import drs_b200
from oracle import dense_topk

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_and_partition():
    for n in (0, 1, 7, 100, 25_000_000):
        for g in (1, 2, 3, 4, 8):
            spans = [drs_b200.shard_bounds(n, r, g) for r in range(g)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a0 <= a1
            per = -(-n // g) if n else 0
            assert all(hi - lo <= per for lo, hi in spans)


def test_cpu_tensors_are_refused_not_routed_to_a_fallback():
    q = torch.randn(4, 16)
    c = torch.randn(32, 16)
    with pytest.raises(RuntimeError, match="no CPU path"):
        drs_b200.search(q, c, 3)
    with pytest.raises(RuntimeError, match="no CPU path"):
        drs_b200.info_nce_loss(q, q, None, 0.05)
    with pytest.raises(RuntimeError, match="no CPU path"):
        drs_b200.merge_shards(torch.zeros(2, 4, 3), torch.zeros(2, 4, 3, dtype=torch.long))


def test_search_argument_validation():
    with pytest.raises(ValueError):
        drs_b200.search(torch.randn(4), torch.randn(8, 4), 1)
    with pytest.raises(TypeError):
        drs_b200.search(torch.zeros(2, 4, dtype=torch.int32), torch.randn(8, 4), 1)


def test_nceloss_module_surface_matches_reference():
    """contrastive_loss.py:47-54,137: ctor takes the loss_config dict; no parameters or buffers,
    so RetrievalModelWrapper.state_dict() gains no keys (strict load, src/model.py:93)."""
    crit = drs_b200.NCELoss({"temperature": 0.05, "cluster": {"num_cluster": [4, 8], "num_neg_proto": 2}})
    assert crit.T == 0.05 and crit.num_cluster == [4, 8] and crit.num_neg_proto == 2
    assert len(crit.state_dict()) == 0 and len(list(crit.parameters())) == 0
    import inspect
    assert list(inspect.signature(crit.forward).parameters) == ["q", "k", "queue", "cluster_result", "index"]


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "information-retrieval-with-contrastive-learning_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inc")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


# ----------------------------------------------------------------------------- world_size 2, gloo
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, nq, nc, dim, k, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import drs_b200 as drs
    from oracle import dense_topk as dt
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(1337)
    corpus = torch.nn.functional.normalize(torch.randn(nc, dim, generator=g), dim=1)
    corpus[nc - 1] = corpus[3]                       # a tie that straddles the two shards
    queries = torch.nn.functional.normalize(torch.randn(nq, dim, generator=g), dim=1)
    queries[0] = corpus[3]
    lo, hi = drs.shard_bounds(nc, rank, world)
    # the local scorer is the oracle here (no GPU in this test): what is under test is the
    # shard planning, the fixed-size slots, the gather layout and the merge order
    v, i = dt.search(queries, corpus[lo:hi], k, index_base=lo)
    slot_s = torch.full((nq, k), float("-inf"))
    slot_i = torch.full((nq, k), -1, dtype=torch.int64)
    slot_s[:, : v.shape[1]] = v
    slot_i[:, : i.shape[1]] = i
    all_s, all_i = drs.all_gather_topk(slot_s, slot_i)
    assert all_s.shape == (world, nq, k)
    mv = torch.empty(nq, 0)
    mi = torch.empty(nq, 0, dtype=torch.int64)
    for r in range(world):
        keep = all_i[r] >= 0
        assert keep.all() or hi - lo < k
        mv, mi = dt._merge_topk(mv, mi, all_s[r], all_i[r], k)
    rv, ri = dt.search(queries, corpus, k)
    ok = torch.equal(mi, ri) and torch.allclose(mv, rv)
    np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([int(ok), int(mi[0, 0]), int(mi[0, 1])]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sharded_exchange_world2_gloo(tmp_path):
    port = _free_port()
    nc = 1001
    mp.spawn(_rank_main, args=(2, port, 9, nc, 32, 5, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        ok, first, second = np.load(tmp_path / f"ok{r}.npy")
        assert ok == 1
        assert (first, second) == (3, nc - 1)        # tie broken by the lower global id across shards


def test_pair_builder_host_logic_and_no_cpu_path():
    """n(n-1)/2 pairs per document, 1 for a one-sentence document (build_docs_sentence_similarity.py:54-57),
    0 for an empty one; without a CUDA device the builder refuses instead of computing on the CPU."""
    from importlib import import_module
    pairs = import_module(drs_b200.__name__ + ".pairs")
    assert pairs.pair_counts(np.array([0, 1, 2, 3, 10])).tolist() == [0, 1, 1, 3, 45]
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            drs_b200.docs_sentence_pairs([np.eye(3)])


def test_default_vectoriser_is_the_references_and_says_so_when_nltk_is_missing():
    """build_docs_sentence_similarity.py:27-43: the default vectoriser needs nltk (WordNet lemmas, English stop
    words); where nltk is absent the builder must say so instead of substituting another tokenizer."""
    try:
        import nltk  # noqa: F401
    except ImportError:
        with pytest.raises(RuntimeError, match="vectorizer"):
            drs_b200.get_docs_sents_similarity([["a b"]], [["a b"]])


def test_moco_queue_and_momentum_update_match_the_reference_methods():
    """contrastive_module.py:42-68, against outputs of the reference's own methods (tests/golden/queue_maintenance.npz)."""
    import os

    import numpy as np
    import torch

    import drs_b200

    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "queue_maintenance.npz"))
    queue, ptr = torch.from_numpy(z["queue0"].copy()), torch.zeros(1, dtype=torch.long)
    i = 0
    while f"keys_{i}" in z.files:
        drs_b200.dequeue_and_enqueue(queue, ptr, torch.from_numpy(z[f"keys_{i}"]))
        np.testing.assert_array_equal(queue.numpy(), z[f"queue_{i}"])
        assert int(ptr) == int(z[f"ptr_{i}"])
        i += 1
    assert i == 6
    pq = [torch.nn.Parameter(torch.from_numpy(z[f"pq_{j}"].copy())) for j in range(4)]
    pk = [torch.nn.Parameter(torch.from_numpy(z[f"pk_{j}"].copy()), requires_grad=False) for j in range(4)]
    for step in range(2):
        drs_b200.momentum_update(pq, pk, float(z["momentum"]))
        for j in range(4):
            np.testing.assert_array_equal(pk[j].detach().numpy(), z[f"pk_{j}_after{step}"])
    q, p = drs_b200.new_queue(16, 48)
    assert q.shape == (16, 48) and int(p) == 0
    torch.testing.assert_close(q.norm(dim=0), torch.ones(48))


def test_clustering_has_no_cpu_path():
    import pytest
    import torch

    import drs_b200

    with pytest.raises(RuntimeError, match="no CPU path"):
        drs_b200.update_centroids(torch.randn(10, 8), torch.zeros(10, dtype=torch.int64), torch.randn(3, 8))


def test_triangle_walk_covers_every_symmetric_tile_once():
    """The symmetric InfoNCE GEMMs (forward LSE, gradient-of-logits) compute only the tiles (m, t), t >= m, of the
    square tile grid, dealt to the clusters by TriangleWalk (gemm_tc.cuh) -- run here on the host through
    drs_debug_triangle_walk: every tile exactly once, for both orders; contiguous pieces differ by at most one tile
    and stay in row-major order; round-robin hands tile i to cluster i mod parts."""
    import ctypes
    from drs_b200 import _lib
    lib = _lib.load()
    for tiles in (1, 2, 3, 8, 31, 32, 64):
        live = tiles * (tiles + 1) // 2
        order_all = [(m, t) for m in range(tiles) for t in range(m, tiles)]
        for parts in (1, 7, 74, 600):
            for order in (0, 1):
                seen, sizes = [], []
                for part in range(parts):
                    buf = (ctypes.c_int * (2 * live))()
                    cnt = ctypes.c_int(0)
                    _lib.check(lib.drs_debug_triangle_walk(tiles, parts, part, order, buf, live, ctypes.byref(cnt)))
                    walk = [(buf[2 * i], buf[2 * i + 1]) for i in range(cnt.value)]
                    sizes.append(len(walk))
                    if order == 0:
                        assert walk == sorted(walk)
                    else:
                        assert walk == order_all[part::parts]
                    seen += walk
                assert sorted(seen) == order_all, (tiles, parts, order)
                assert max(sizes) - min(sizes) <= 1
