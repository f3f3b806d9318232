"""B200-native dense-retrieval scoring engine for the FEVER contrastive-IR project.

Drop-in for the query-by-document similarity-and-select path of
PM25/Information-Retrieval-with-Contrastive-Learning:

* ``search`` / ``DenseIndex`` / ``ShardedDenseIndex`` -- dense claim x corpus scores + top-k ids
  (src/evaluation.py:105-116 call site, ``TfidfDocRanker.closest_docs`` signature);
* ``NCELoss`` / ``info_nce_loss`` -- in-batch InfoNCE logits + softmax-CE forward/backward
  (src/contrastor/contrastive_loss.py:47-141).

Everything computes in hand-written sm_100a CUDA behind the C ABI of include/drs_b200.h
(libdrs_b200.so, built in-tree by ``build.build()``); there is no CPU fallback.
"""
from . import _lib, build
from ._lib import get_option, set_option
from .loss import InfoNCE, NCELoss, info_nce_loss, proto_nce_loss
from .pairs import doc_sentence_pairs_arrays, docs_sentence_pairs, get_docs_sents_similarity
from .store import load_dense_index, save_dense_index
from .moco import dequeue_and_enqueue, momentum_update, new_queue
from .clustering import Clustering, cluster_density, run_kmeans, update_centroids, vector_to_array
from .retrieval import (DenseDocRanker, DenseIndex, FlatL2Index, ShardedDenseIndex, all_gather_topk, flat_l2_search, merge_shards,
                        paired_scores, rerank, search, shard_bounds)

__all__ = [
    "search", "rerank", "paired_scores", "flat_l2_search", "FlatL2Index", "merge_shards", "DenseIndex", "ShardedDenseIndex", "all_gather_topk", "shard_bounds",
    "DenseDocRanker", "save_dense_index", "load_dense_index", "docs_sentence_pairs", "doc_sentence_pairs_arrays", "get_docs_sents_similarity",
    "NCELoss", "InfoNCE", "info_nce_loss", "proto_nce_loss", "set_option", "get_option", "build",
    "dequeue_and_enqueue", "momentum_update", "new_queue", "Clustering", "cluster_density", "run_kmeans", "update_centroids", "vector_to_array",
]
