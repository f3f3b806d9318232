/*
 * drs_b200 -- C ABI of the B200 dense-retrieval scoring engine.
 *
 * The reference (PM25/Information-Retrieval-with-Contrastive-Learning) is pure Python and has no
 * FFI of its own; these entry points are what a binding for its similarity-and-select path binds
 * (see INTEGRATION.md for the ctypes stubs).  Each function names the reference call it replaces.
 *
 * Conventions
 *   - every pointer marked "device" is a CUDA device pointer owned by the caller; the library
 *     allocates nothing and keeps no state between calls except the options set below;
 *   - matrices are row-major and contiguous; `stream` is a cudaStream_t passed as void*
 *     (NULL = the legacy default stream); calls are asynchronous on that stream;
 *   - workspace: ask `*_workspace_bytes`, pass a 256-byte aligned device buffer at least that large;
 *   - return value: DRS_OK (0) or an error code; `drs_last_error()` has the message
 *     (thread local).  Nothing is thrown across the boundary.
 */
#ifndef DRS_B200_H_
#define DRS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { DRS_OK = 0, DRS_ERR_INVALID = 1, DRS_ERR_CUDA = 2, DRS_ERR_UNSUPPORTED = 3, DRS_ERR_WORKSPACE = 4 };
enum { DRS_F32 = 0, DRS_BF16 = 1, DRS_F16 = 2 };  /* DRS_F16: IEEE half operands on the same tcgen05 path as DRS_BF16 */
#define DRS_MAX_K 256  /* k > 16: adaptive passes over the corpus (at most ceil(k/16), usually one) */

int drs_version(void);
const char* drs_last_error(void);

/* Test / tuning knobs.  "search.cta_group": 0 auto, 1 single CTA (128x256 MMA), 2 CTA pair
 * (256x256 MMA).  "search.num_ctas": 0 auto (all SMs), else the persistent grid size.
 * "search.splits": 0 auto, else the number of corpus splits.
 * "tune.cooperative" (default 1): the kernels that spin on grid-wide flags (the scan's round barrier, the fused
 * exchange) are launched cooperatively, so the driver guarantees that the whole grid is resident or refuses; the
 * scan additionally asks cudaOccupancyMaxActiveClusters and runs WITHOUT the barrier when the grid cannot be
 * co-resident ("debug.coop_fallbacks" counts those launches).  Calls on one workspace must be issued from one
 * stream at a time.
 * "tune.k_split" (default 0 = auto; 1 = off; n > 1 = n slices): the loss gradients wrt q through a long K
 * (dq = Hq x queue, dq = W x prototypes: one or two output tiles, K = queue length) are computed as K slices on
 * many clusters and summed in slice order (deterministic).  Set it BEFORE the *_workspace_bytes query of the call.
 * "tune.symmetric_lse" (default 1 = when it pays, from 2N = 6144 rows at dim = 768; 0 = never; 2 = always): the bf16
 * InfoNCE forward computes only the tiles of F F^T on and above the diagonal (whole 256-row tiles) against one bounded
 * reference; logits whose span exceeds the bound (checked on the device) take the full-matrix schedule inside the same
 * launch.  Changes the workspace size: set it BEFORE drs_infonce_workspace_bytes.  "tune.symmetric_grad" (default 2): the
 * same for the gradient-of-logits matrix of the backward (0 = full matrix).  "tune.triangle_order": 0 contiguous pieces
 * of the tile triangle per cluster (default), 1 round-robin. */
int drs_set_option(const char* name, int value);
/* Debug: {flag, tag, block, thread, parity, extra} of the last pipeline wait that timed out (a
 * kernel whose mbarrier wait exceeds a few seconds records this in mapped host memory and traps,
 * so a broken pipeline surfaces as a CUDA error instead of a hang).  flag == 0: none. */
int drs_debug_hang_report(unsigned int out[6]);
int drs_get_option(const char* name, int* value);
/* Debug: how many clusters of `cluster_size` CTAs of the scan kernel (one CTA per SM, ~198 KB of shared memory)
 * the current device can hold at once (cudaOccupancyMaxActiveClusters). */
int drs_debug_max_clusters(int cluster_size, int* out);
/* Debug (host only, no GPU needed): the tiles (m, t), t >= m, of a `tiles` x `tiles` symmetric tile grid that cluster `part`
 * of `parts` walks in the symmetric InfoNCE GEMMs (order 0: contiguous pieces, 1: round-robin; "tune.triangle_order").
 * Writes up to `capacity` (m, t) pairs to out_mt and the true number to *count. */
int drs_debug_triangle_walk(int tiles, int parts, int part, int order, int* out_mt, int capacity, int* count);

/*
 * Dense claim x corpus scoring with fused top-k select.
 * Replaces: the dense scoring intended at src/evaluation.py:105-116 (dot products of ctx2vec
 * embeddings) + the select of TfidfDocRanker.closest_docs
 * (preprocessing/drqa/retriever/tfidf_doc_ranker.py:60-75), batched over claims like
 * batch_closest_docs (:77-84).
 *
 *   queries  device [nq, dim]   dtype DRS_BF16 / DRS_F16 (tcgen05, fp32 accumulate) or DRS_F32 (the exact-comparison path: fp32
 *                                operands on tcgen05 as a 3 x TF32 split, <= 4e-6 relative; FFMA kernel when dim % 4 != 0 or
 *                                "search.fp32_mode" = 1)
 *   corpus   device [nc, dim]   same dtype
 *   out_scores device [nq, k] fp32, descending;  out_ids device [nq, k] int64 = row + id_base
 *   ties are broken by the lower row index; when nc < k the tail is (-inf, -1).
 * NaN scores (a NaN in a claim or corpus row) are never selected -- they rank below every number, the way numpy's
 * argpartition / argsort order NaN last in closest_docs (tfidf_doc_ranker.py:70-71); torch.topk would rank them
 * FIRST.  drs_rerank and drs_search_l2 follow the same rule; a claim whose scores are all NaN returns (-inf, -1).
 * DRS_BF16 / DRS_F16 need dim % 8 == 0 and 16-byte aligned base pointers (TMA); k <= DRS_MAX_K.
 * The running top-k lists hold 16 (or, for 17 <= k on corpora with few splits, 32) entries per (claim, corpus
 * split) in registers.  k beyond the list capacity stays exact: the select emits picks only while no
 * split's full list could hide a better row, and a claim that stops early is continued by a rescan strictly
 * below its last pick.  At most ceil(k/capacity) passes are enqueued; a pass with no open claim returns at
 * once on the device (one pass does all the work unless more than 16 of a claim's top-k fall into one of
 * the ~74+ splits).
 */
int drs_search_workspace_bytes(int64_t nq, int64_t nc, int dim, int k, int dtype, size_t* bytes);
int drs_search(const void* queries, int64_t nq, const void* corpus, int64_t nc, int dim, int dtype, int k,
               int64_t id_base, float* out_scores, int64_t* out_ids, void* workspace, size_t workspace_bytes,
               void* stream);

/* Debug: after a k > 32 drs_search / drs_search_l2 on `workspace`, out[p] = the number of claims that
 * pass p left open (needing a rescan below their last certain pick).  Synchronises the stream. */
int drs_debug_open_claims(const void* workspace, unsigned int out[8], void* stream);

/* The two phases of drs_search, separately launchable (drs_search == scan then select on the same
 * stream): `scan` is the fused score GEMM + per-row running top-k that leaves the per-split
 * candidate keys in the workspace; `select` reduces them to the final k.  Same arguments. */
int drs_search_scan(const void* queries, int64_t nq, const void* corpus, int64_t nc, int dim, int dtype, int k,
                    void* workspace, size_t workspace_bytes, void* stream);
int drs_search_select(const void* workspace, int64_t nq, int64_t nc, int dim, int dtype, int k, int64_t id_base,
                      float* out_scores, int64_t* out_ids, void* stream);

/*
 * Exact squared-L2 nearest neighbours (same kernel, ranked value 2 x.c - |c|^2, reported as
 * |x|^2 + |c|^2 - 2 x.c, ascending; ties -> lower row).
 * Replaces: faiss GpuIndexFlatL2.search as used for the k-means assignment at
 * src/contrastor/utils.py:64-67 (`D, I = index.search(x, 1)`).  PARITY UNPINNED: faiss is an
 * unpinned dependency that is not part of the reference tree.
 */
int drs_search_l2_workspace_bytes(int64_t nq, int64_t nc, int dim, int k, int dtype, size_t* bytes);
int drs_search_l2(const void* queries, int64_t nq, const void* corpus, int64_t nc, int dim, int dtype, int k,
                  int64_t id_base, float* out_dist, int64_t* out_ids, void* workspace, size_t workspace_bytes,
                  void* stream);

/*
 * The other half of a k-means iteration and the distance statistics of the prototype "concentration" estimate.
 * Replaces: the centroid update inside faiss.Clustering.train (src/contrastor/utils.py:28-36, :64 -- third-party,
 * unpinned, absent: PARITY UNPINNED for the training) and the per-cluster distance lists of :73-83 (the reference's
 * own numpy loops; pinned by tests/golden/kmeans_density.npz).
 *
 *   x        device [n, dim] fp32 samples (may be NULL when centroids is NULL)
 *   order    device [n] int64: sample indices sorted by assigned cluster, STABLE (members in ascending sample order)
 *   offsets  device [k + 1] int64: cluster c owns order[offsets[c] .. offsets[c + 1])
 *   dist     device [n] fp32 squared distances to the assigned centroid (drs_search_l2 output), or NULL
 *   centroids device [k, dim] fp32, in/out, or NULL: centroid c = mean of its members, summed in fp64 in member order
 *            (deterministic; equal to a sequential float64 sum over the samples); an EMPTY cluster keeps its value
 *   sum_sqrt_dist device [k] fp32 out (with dist): sum over the members of sqrt(dist)
 */
int drs_cluster_update(const float* x, int64_t n, int dim, const int64_t* order, const int64_t* offsets, int64_t k,
                       const float* dist, float* centroids, float* sum_sqrt_dist, void* stream);

/*
 * Sharded search with the exchange fused in: scan of this rank's shard, then ONE kernel that selects the
 * shard's top-k, stores it into every peer's gather buffer over NVLink peer memory, waits (per block of 32
 * claims, acquire/release flags at system scope) for the other shards' lists and merges them.  Replaces the
 * select + 2 x all-gather + merge sequence of the NCCL form; the result is the same on every rank and equal to
 * a single-GPU search of the whole corpus.  (The reference is single-device: nothing upstream to cite.)
 *   peer_scores / peer_ids / peer_flags: HOST arrays of `world` device pointers -- rank r's parity-0 gather
 *       buffers ([world][nq][k] fp32 and int64) and flag array (drs_exchange_flag_bytes, zeroed once at
 *       allocation), all mapped into this process (cudaIpc / symmetric memory); entry [rank] is this rank's own.
 *   parity_stride_bytes: distance from a parity-0 gather buffer to its parity-1 twin (a rank may run one call
 *       ahead of a peer that is still merging the previous one, so calls alternate between two buffer sets).
 *   call_counter: device uint32 of THIS rank, zero at allocation; the library bumps it once per call and derives
 *       the call's epoch and buffer parity from it on the device -- no argument changes between calls, so the
 *       sequence can be captured in a CUDA graph and replayed.
 * All ranks must make the same sequence of calls (it is a collective).  k <= 16, world <= 8.
 */
int drs_exchange_flag_bytes(int64_t max_nq, int world, size_t* bytes);
int drs_search_sharded_p2p(const void* queries, int64_t nq, const void* corpus, int64_t nc_local, int dim, int dtype,
                           int k, int64_t id_base, int rank, int world, void* const* peer_scores,
                           void* const* peer_ids, void* const* peer_flags, size_t parity_stride_bytes,
                           uint32_t* call_counter, float* out_scores, int64_t* out_ids, void* workspace,
                           size_t workspace_bytes, void* stream);

/*
 * Query-sliced exchange of per-shard top-k lists for ANY k <= DRS_MAX_K (the fused kernel above serves k <= 16):
 * rank s owns the claims [s * ceil(nq/world), ...); every rank stores the lists of slice s into rank s's gather
 * buffer over NVLink peer memory, rank s merges the `world` sorted runs of each of its claims and stores the final
 * list into every rank's result buffer -- one kernel, three flag-ordered phases, world x fewer bytes and merges per
 * GPU than gathering every list everywhere (BASELINE configs[4]: 65 536 claims, top-100, 8 GPUs).
 *   local_scores / local_ids: device [nq, k], this shard's sorted lists with GLOBAL ids (drs_search with id_base;
 *       (-inf, -1) where the shard has fewer than k rows)
 *   peer_bases: HOST array of `world` device pointers to each rank's exchange buffer (drs_exchange_sliced_bytes
 *       bytes, zeroed once, peer-mapped; the same max_nq / max_entries on every rank).  Needs
 *       ceil(nq/world) * world * k <= max_entries and nq <= max_nq.
 *   call_counter: device uint32 of this rank, zero at allocation (epochs and buffer parity derive from it).
 * Collective: every rank makes the same calls.  The result is identical on every rank and equal to the
 * (score desc, id asc) merge of all shards' lists, i.e. to a single-GPU search of the whole corpus.
 */
int drs_exchange_sliced_bytes(int64_t max_nq, int64_t max_entries, int world, size_t* bytes);
int drs_exchange_sliced(const float* local_scores, const int64_t* local_ids, int64_t nq, int k, int rank, int world,
                        void* const* peer_bases, int64_t max_nq, int64_t max_entries, uint32_t* call_counter,
                        float* out_scores, int64_t* out_ids, void* stream);

/*
 * Peer-mapped device buffers for the two exchange kernels above, over CUDA IPC: every rank (one process per GPU of
 * one box) allocates its buffer with drs_peer_alloc -- zero-filled, which is the state both kernels' flags and epochs
 * start from -- hands the 64-byte handle to its peers (any transport: the Python side all-gathers them through
 * torch.distributed), and maps each peer's buffer with drs_peer_open (peer access is enabled on first use).
 * Close every mapping and free the buffer only after all ranks have finished their last exchange call.
 */
int drs_peer_alloc(size_t bytes, void** ptr, unsigned char handle[64]);
int drs_peer_open(const unsigned char handle[64], void** ptr);
int drs_peer_close(void* ptr);
int drs_peer_free(void* ptr);

/*
 * Candidate-restricted re-rank: score each claim against ITS OWN candidate rows only and keep the best k.
 * Replaces: the dense stage of the report's pipeline "TF-IDF top-100 -> contrastive re-rank -> top-15"
 * (report.pdf section 3.2) at the call site src/evaluation.py:105-116, fed by the sparse candidates of
 * documents_filtering (src/evaluation.py:57-83) / closest_docs (tfidf_doc_ranker.py:60-75).
 *   cand_ids device [nq, num_cand] int64 corpus rows; id < 0 or >= nc is padding and is ignored
 *   out_scores device [nq, k] fp32 descending; out_ids device [nq, k] int64 (the candidate's row);
 *   ties -> lower row; a row listed twice is reported once; missing entries are (-inf, -1).
 * HBM-bound gather (one corpus row read per pair); no workspace.  1 <= k <= num_cand.
 */
int drs_rerank(const void* queries, int64_t nq, const void* corpus, int64_t nc, int dim, int dtype,
               const int64_t* cand_ids, int num_cand, int k, float* out_scores, int64_t* out_ids, void* stream);

/*
 * out[i] = a[i] . b[i] for two [n, dim] matrices.
 * Replaces: `(clm_vec * evdn_vec).sum(dim=-1)` of the commented dense evaluation, src/evaluation.py:112,115.
 */
int drs_pair_scores(const void* a, const void* b, int64_t n, int dim, int dtype, float* out, void* stream);

/*
 * Per-document sentence-pair similarity, a whole batch of documents per call.
 * Replaces: the body of get_docs_sents_similarity, preprocessing/build_docs_sentence_similarity.py:48-66 --
 * sklearn cosine_similarity(doc_tfidf, doc_tfidf) (:50), the strict-upper-triangle walk (:59-63; a
 * one-sentence document yields the single pair (0,0), :54-57) and the stable descending sort (:65).
 *   indptr [num_sentences+1] int64, indices [nnz] int32 (sorted within a row), data [nnz] float64:
 *       the CSR rows `vectorizer.transform(doc)` (:49) yields, stacked over all documents (device)
 *   doc_offsets  [num_docs+1] int64: sentence range of each document (device)
 *   pair_offsets [num_docs+1] int64: output range of each document = prefix sum of n(n-1)/2
 *       (1 for n == 1, 0 for n == 0) (device)
 *   out_i, out_j [total_pairs] int32, out_score [total_pairs] float64: per document, pairs by score
 *       descending, equal scores in (i, j) lexicographic order (device)
 * float64 throughout, operation order = the reference's: results are bit-identical to it.
 */
int drs_doc_pairs_workspace_bytes(int64_t nnz, int64_t total_pairs, size_t* bytes);
int drs_doc_sentence_pairs(const int64_t* indptr, const int32_t* indices, const double* data, int64_t num_sentences,
                           const int64_t* doc_offsets, int64_t num_docs, const int64_t* pair_offsets,
                           int64_t total_pairs, int64_t nnz, int32_t* out_i, int32_t* out_j, double* out_score,
                           void* workspace, size_t workspace_bytes, void* stream);

/*
 * Merge the per-shard top-k lists of a row-sharded corpus (after the all-gather):
 *   scores device [num_shards, nq, k], ids device [num_shards, nq, k] (id < 0 = empty slot)
 *   -> out_scores / out_ids device [nq, k], ordered by (score desc, id asc).
 * No reference counterpart (the reference is single-device); SURVEY.md section 8(e).
 */
int drs_merge_shards(const float* scores, const int64_t* ids, int num_shards, int64_t nq, int k, float* out_scores,
                     int64_t* out_ids, void* stream);

/*
 * In-batch InfoNCE (SimCLR form) with fused logits + softmax cross-entropy.
 * Replaces: NCELoss._compute_info_loss (src/contrastor/contrastive_loss.py:56-93).
 *   q, k   device [n, dim] fp32 (L2-normalised by the caller, contrastive_module.py:111)
 *   queue  device [dim, queue_len] fp32 or NULL (MoCo negatives, :77-82; rows n..2n-1 reuse q's
 *          queue logits exactly like the reference's .repeat(2, 1))
 *   precision DRS_F32: fp32 FFMA logits;  DRS_BF16: inputs rounded to bf16, tcgen05, fp32 accumulate
 *   loss   device [1] fp32 = sum_i CE_i / 2;   lse device [2n] fp32 (saved for backward)
 * Backward: dq, dk device [n, dim] fp32 = grad_loss[0] * dloss/d{q,k}; grad_loss device [1].
 */
int drs_infonce_workspace_bytes(int64_t n, int dim, int64_t queue_len, int precision, size_t* bytes);
int drs_infonce_forward(const float* q, const float* k, const float* queue, int64_t n, int dim, int64_t queue_len,
                        float inv_temperature, int precision, float* loss, float* lse, void* workspace,
                        size_t workspace_bytes, void* stream);
int drs_infonce_backward(const float* q, const float* k, const float* queue, int64_t n, int dim, int64_t queue_len,
                         float inv_temperature, int precision, const float* lse, const float* grad_loss, float* dq,
                         float* dk, void* workspace, size_t workspace_bytes, void* stream);
/* Same, for a caller that kept the forward's workspace untouched (same q, k, queue, shapes, precision): the packed
 * operands and the row LSEs the forward left there are reused instead of being staged again. */
int drs_infonce_backward_staged(const float* q, const float* k, const float* queue, int64_t n, int dim,
                                int64_t queue_len, float inv_temperature, int precision, const float* lse,
                                const float* grad_loss, float* dq, float* dk, void* workspace, size_t workspace_bytes,
                                void* stream);

/*
 * MoCo-form InfoNCE.  Replaces: InfoNCE.forward (src/contrastor/contrastive_loss.py:26-44):
 * logits = [q.k | q @ queue] / T, label 0, CrossEntropyLoss with MEAN reduction.
 *   q, k device [n, dim] fp32; queue device [dim, queue_len] fp32 (required; no gradient, :32)
 *   loss device [1]; lse device [n] (saved for backward); dq, dk device [n, dim].
 */
int drs_moco_workspace_bytes(int64_t n, int dim, int64_t queue_len, int precision, size_t* bytes);
int drs_moco_forward(const float* q, const float* k, const float* queue, int64_t n, int dim, int64_t queue_len,
                     float inv_temperature, int precision, float* loss, float* lse, void* workspace,
                     size_t workspace_bytes, void* stream);
int drs_moco_backward(const float* q, const float* k, const float* queue, int64_t n, int dim, int64_t queue_len,
                      float inv_temperature, int precision, const float* lse, const float* grad_loss, float* dq,
                      float* dk, void* workspace, size_t workspace_bytes, void* stream);

/*
 * ProtoNCE, one cluster set.  Replaces the arithmetic of NCELoss._compute_proto_loss
 * (src/contrastor/contrastive_loss.py:112-131) after the prototype selection (:99-110, host side):
 *   protos device [num_protos, dim] fp32 = cat(pos_prototypes (n rows), neg_prototypes)  (:112)
 *   inv_temps device [num_protos] fp32 = 1 / density of each selected prototype          (:122-124)
 *   logits = q @ protos^T * inv_temps; label of row i is i (:118-119); loss[0] = CE SUM (:131).
 * Backward: dq (+)= grad_loss[0] * dloss/dq  (accumulate != 0 adds into dq: sets are summed, :129-131).
 */
int drs_proto_workspace_bytes(int64_t n, int dim, int64_t num_protos, int precision, size_t* bytes);
int drs_proto_forward(const float* q, const float* protos, const float* inv_temps, int64_t n, int dim,
                      int64_t num_protos, int precision, float* loss, float* lse, void* workspace,
                      size_t workspace_bytes, void* stream);
int drs_proto_backward(const float* q, const float* protos, const float* inv_temps, int64_t n, int dim,
                       int64_t num_protos, int precision, const float* lse, const float* grad_loss, float* dq,
                       int accumulate, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DRS_B200_H_ */
