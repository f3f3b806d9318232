// C ABI of the engine (include/drs_b200.h).  Single translation unit: kernels come in through
// the .cuh files.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -shared ...
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <type_traits>

#include "../../include/drs_b200.h"
#include "epilogues.cuh"
#include "exchange.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "infonce.cuh"
#include "kmeans.cuh"
#include "merge.cuh"
#include "pairs.cuh"
#include "rerank.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// mapped pinned host memory the kernels write their hang report to (survives a trapped context)
drs::HangReport* g_hang_host = nullptr;

// A launch/runtime error.  If a kernel trapped on an mbarrier timeout, say which wait it was.
int cuda_fail(cudaError_t e, const char* what) {
  if (g_hang_host && g_hang_host->flag) {
    const drs::HangReport rep = *g_hang_host;
    return fail(DRS_ERR_CUDA, "%s: %s (pipeline wait timed out: tag=%u block=%u thread=%u parity=%u extra=%u)", what,
                cudaGetErrorString(e), rep.tag, rep.block, rep.thread, rep.parity, rep.extra);
  }
  return fail(DRS_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}
#define DRS_CUDA(call)                                  \
  do {                                                  \
    cudaError_t e_ = (call);                            \
    if (e_ != cudaSuccess) return cuda_fail(e_, #call); \
  } while (0)

struct Options {
  int cta_group = 0;
  int num_ctas = 0;
  int splits = 0;
  int infonce_cta_group = 0;
  int debug_flags = 0;
  int b_hint = 0;
  int stagger_cycles = 0;
  int round_barrier = 1;
  int seed_thresholds = 1;
  int symmetric_grad = 2;   // InfoNCE backward: 0 full H, 1 upper tiles computed + mirrored stores, 2 upper tiles only (dF reads transposed)
  int symmetric_lse = 1;    // InfoNCE forward: only the tiles of S = F F^T on and above the diagonal (whole 256-row tiles, bounded logits): 0 off, 1 when it pays, 2 always
  int triangle_order = 0;   // symmetric InfoNCE GEMMs: 0 contiguous pieces of the tile triangle per cluster, 1 round-robin (TriangleWalk)
  int tma_store = 1;       // InfoNCE backward: the gradient-of-logits tiles leave through TMA stores
  int fp32_tile = 128;     // B-tile width of the fp32 tensor-core scan: 128 (two accumulator stages: 0.66 ms on config 0) or 256 (one: 0.73 ms)
  int m_block = 0;         // search: A tiles per super-block of the unit order: 0 auto (= clusters), -1 off
  int df_tile = 0;         // InfoNCE dF = H F GEMM tile width: 0 auto, 256, 192
  int fp32_mode = 0;       // DRS_F32 search: 0 = 3 x TF32 on tcgen05 (when dim % 4 == 0 and 16-byte aligned), 1 = FFMA kernel
  int k_split = 0;         // loss gradients wrt q through a long K (dq = Hq x queue, dq = W x prototypes): 0 auto, 1 off, > 1 that many K slices
  int cooperative = 1;     // launch the kernels that spin on grid-wide flags cooperatively (driver-checked co-residency)
  int coop_fallbacks = 0;  // read-only counter: launches that gave up the round barrier (grid not co-resident)
} g_opt;

// DRS_COOPERATIVE=0/1 overrides "tune.cooperative" from the environment (read once).  Nsight Compute cannot replay a
// cooperative launch of a clustered kernel on this driver (the profiled process dies at that launch), so when a
// profiler's injection library is announced in the environment and nothing was said, the attribute is left out: the
// barrier then relies on the occupancy check alone, as it did before cooperative launches were introduced.
void init_options_from_env() {
  static bool done = false;
  if (done) return;
  done = true;
  if (const char* v = getenv("DRS_COOPERATIVE")) g_opt.cooperative = atoi(v) != 0;
  else if (getenv("CUDA_INJECTION64_PATH") || getenv("NV_NSIGHT_INJECTION_PORT_BASE") || getenv("NVTX_INJECTION64_PATH"))
    g_opt.cooperative = 0;
}

struct DeviceInfo {
  int device = -1;
  int num_sms = 0;
  int cc_major = 0, cc_minor = 0;
};
int get_device_info(DeviceInfo* out) {
  static thread_local DeviceInfo cache[64];
  init_options_from_env();
  int dev = 0;
  DRS_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(DRS_ERR_INVALID, "device ordinal %d out of range", dev);
  if (cache[dev].device != dev) {
    cache[dev].device = dev;
    DRS_CUDA(cudaDeviceGetAttribute(&cache[dev].num_sms, cudaDevAttrMultiProcessorCount, dev));
    DRS_CUDA(cudaDeviceGetAttribute(&cache[dev].cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    DRS_CUDA(cudaDeviceGetAttribute(&cache[dev].cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
    if (!g_hang_host) {
      void* h = nullptr;
      if (cudaHostAlloc(&h, sizeof(drs::HangReport), cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess) {
        memset(h, 0, sizeof(drs::HangReport));
        g_hang_host = static_cast<drs::HangReport*>(h);
      }
    }
    if (g_hang_host) {
      void* d = nullptr;
      if (cudaHostGetDevicePointer(&d, g_hang_host, 0) == cudaSuccess)
        DRS_CUDA(cudaMemcpyToSymbol(drs::g_hang_ptr, &d, sizeof(d)));
    }
  }
  *out = cache[dev];
  return DRS_OK;
}

// ------------------------------------------------------------------ TMA descriptors
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
// [rows, dim] bf16 row-major; box = 64 (K) x box_rows, 128-byte swizzle, zero fill out of bounds
inline bool is_16bit(int dtype) { return dtype == DRS_BF16 || dtype == DRS_F16; }

// `pitch` = elements from one row to the next (0: dim).  Columns in [dim, pitch) are never read: a box that
// reaches past `dim` is zero-filled, which is how ragged K extents (prototype counts that are not a multiple
// of 8) get a 16-byte row pitch without a padded copy.
int make_tmap_bf16(CUtensorMap* m, const void* base, int64_t rows, int dim, int box_rows, bool f16 = false, int64_t pitch = 0) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(DRS_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(dim), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(pitch > 0 ? pitch : dim) * 2};
  cuuint32_t box[2] = {64u, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(DRS_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d)", static_cast<int>(r));
  return DRS_OK;
}

// [rows, dim] fp32 row-major; box = 32 (K, one 128-byte swizzle row) x box_rows
int make_tmap_f32(CUtensorMap* m, const void* base, int64_t rows, int dim, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(DRS_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(dim), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(dim) * 4};
  cuuint32_t box[2] = {32u, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(DRS_ERR_CUDA, "cuTensorMapEncodeTiled (fp32) failed (CUresult %d)", static_cast<int>(r));
  return DRS_OK;
}

// a bf16 OUTPUT matrix [rows, cols] (row pitch in elements) for the epilogue's TMA stores: box 32 x 32, 64-byte swizzle
int make_tmap_out_bf16(CUtensorMap* m, void* base, int64_t rows, int64_t cols, int64_t pitch) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(DRS_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(pitch) * 2};
  cuuint32_t box[2] = {32u, 32u};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(DRS_ERR_CUDA, "cuTensorMapEncodeTiled (output) failed (CUresult %d)", static_cast<int>(r));
  return DRS_OK;
}

// ------------------------------------------------------------------ work decomposition
// units = num_m_tiles * splits, dealt round-robin to `groups` persistent clusters; a unit is `tiles_per_split`
// consecutive B tiles against one A tile.
//   search (long_units = false): the smallest split count that makes the units a multiple of the groups
//     (every cluster gets the same number of equally long units), capped by the number of B tiles; many
//     medium units also feed the threshold seeding.
//   square GEMMs of the losses (long_units = true): the makespan rounds * tiles_per_split is minimised and
//     ties go to the LONGER unit.  A cluster that changes its A tile fetches it across the die-to-die link
//     instead of from the near L2: on 8192^2 x 768 one-tile units ran the mainloop in 111 us, two-tile units
//     (same makespan, half the unit starts) in 78 us -- cuBLAS takes 83 us for that GEMM.
drs::GemmShape plan_shape(int64_t rows_a, int64_t rows_b, int dim_k_blocks, int rows_per_m_tile, int rows_per_b_tile,
                          int groups, int forced_splits, int col_groups, bool long_units = false) {
  drs::GemmShape s;
  s.rows_a = static_cast<int>(rows_a);
  s.rows_b = static_cast<int>(rows_b);
  s.num_k_blocks = dim_k_blocks;
  s.num_m_tiles = static_cast<int>((rows_a + rows_per_m_tile - 1) / rows_per_m_tile);
  s.total_b_tiles = static_cast<int>((rows_b + rows_per_b_tile - 1) / rows_per_b_tile);
  int splits = forced_splits > 0 ? forced_splits : groups / std::gcd(s.num_m_tiles, groups);
  if (forced_splits <= 0 && long_units) {
    long long best = -1;
    for (int cand = std::min(s.total_b_tiles, 4 * groups); cand >= 1; --cand) {  // descending: the last tie wins
      const int tps = (s.total_b_tiles + cand - 1) / cand;
      const int real = (s.total_b_tiles + tps - 1) / tps;
      const long long units = static_cast<long long>(s.num_m_tiles) * real;
      const long long makespan = (units + groups - 1) / groups * tps;
      if (best < 0 || makespan <= best) { best = makespan; splits = cand; }
    }
  }
  splits = std::max(1, std::min(splits, s.total_b_tiles));
  s.tiles_per_split = (s.total_b_tiles + splits - 1) / splits;
  s.num_splits = (s.total_b_tiles + s.tiles_per_split - 1) / s.tiles_per_split;  // no empty split
  s.col_groups = col_groups;
  s.debug_flags = g_opt.debug_flags;
  s.b_hint = g_opt.b_hint;
  s.stagger_cycles = g_opt.stagger_cycles;
  s.round_counter = nullptr;
  s.active = nullptr;
  s.active_mode = 0;
  s.active_scale = 0.f;
  s.f16_operands = 0;
  s.m_block = 0;
  s.epi_tma_store = 0;
  s.a_sym = 0;
  s.skip_below_diagonal = 0;
  s.tri_order = g_opt.triangle_order;
  s.k_splits = 1;
  s.kb_per_split = dim_k_blocks;
  return s;
}

// Split-K for a GEMM whose output is a tile or two but whose K is long (the loss gradients wrt q through the MoCo queue
// or the prototypes: [n x K] x [K x dim] with K = 12 544 at n = dim = 128): without it one cluster walks all of K (67 us
// at the reference's shapes) while 73 idle.  `total_kb` = K blocks of the kernel that will run (64 wide on the
// tensor-core path, 16 on the FFMA path); slices of >= `min_kb` blocks, at most 64 of them (the partial products are
// [slices][rows_a][rows_b] fp32 in the workspace, summed in slice order by splitk_reduce_kernel).
void plan_k_split(drs::GemmShape* s, int total_kb, int groups, int min_kb) {
  s->k_splits = 1;
  s->kb_per_split = total_kb;
  const int base = s->num_m_tiles * s->num_splits;
  int want = 1;
  if (g_opt.k_split > 1) want = g_opt.k_split;
  else if (g_opt.k_split == 0 && 2 * base <= groups && total_kb >= 4 * min_kb) want = std::min(groups / base, total_kb / min_kb);
  want = std::max(1, std::min({want, 64, total_kb}));
  if (want > 1) {
    s->kb_per_split = (total_kb + want - 1) / want;
    s->k_splits = (total_kb + s->kb_per_split - 1) / s->kb_per_split;   // no empty slice
  }
}
inline size_t k_split_bytes(const drs::GemmShape& s) {
  return s.k_splits > 1 ? static_cast<size_t>(s.k_splits) * s.rows_a * s.rows_b * sizeof(float) : 0;
}

constexpr int kTcColGroups = drs::GemmCfg<1>::EPI_GROUPS;
inline size_t align256s(size_t x) { return (x + 255) & ~size_t(255); }
constexpr size_t kWsRoundBytes = 256;   // round-barrier counter, zeroed before every scan
constexpr size_t kWsHeaderBytes = 512;  // + per-pass "claims still open" counters of the adaptive multi-pass search
inline int num_slots(const drs::GemmShape& s) { return s.num_splits * s.col_groups; }

struct SearchPlan {
  int dtype;
  int cg;          // bf16: CTA group
  int kcap;        // 16 or 32
  int k;
  int passes;
  int grid;        // CTAs
  drs::GemmShape shape;
  size_t bound_bytes;
  size_t cand_bytes;
  size_t seed_bytes;  // threshold seeds [nq][seed_slots] (epilogues.cuh), zeroed before every scan
  int seed_slots;     // certificates kept per claim: ceil(k / seed_group)
  int seed_group;     // rows certified by one published value (epilogues.cuh)
  int seed_chunks;    // values a unit publishes: seed_chunks * seed_group <= entries selected per pass
  bool f32_tc;        // DRS_F32 on the tensor cores (3 x TF32): claims and corpus are split into hi + lo first
  int f32_bn;         // B-tile width of the fp32 tensor-core scan: 256 (one accumulator stage) or 128 (two)
  bool f32_prepared;  // the split of this call has been enqueued (multi-pass searches split once)
  size_t f32_a_bytes; // one padded [a_rows, dim] fp32 copy of the claims (two are kept: hi, lo)
  size_t f32_b_bytes; // the corpus residual [nc, dim] fp32
  size_t pad_bytes;   // bf16: zero-padded copy of the claims when nq is not a multiple of the A tile (see scan_pass)
  int64_t a_rows;     // rows of the A operand as the tensor map sees it (nq rounded up when padded)
  size_t ws_bytes;
};

int plan_search(int64_t nq, int64_t nc, int dim, int k, int dtype, SearchPlan* p, bool allow_f32_tc = true) {
  if (nq <= 0 || nc <= 0 || dim <= 0) return fail(DRS_ERR_INVALID, "nq, nc, dim must be positive (got %lld, %lld, %d)", (long long)nq, (long long)nc, dim);
  if (k <= 0 || k > DRS_MAX_K) return fail(DRS_ERR_UNSUPPORTED, "k must be in [1, %d] (got %d)", DRS_MAX_K, k);
  if (nq > 0x7fffffffLL - 512 || nc > 0x7fffffffLL - 512) return fail(DRS_ERR_UNSUPPORTED, "nq and nc must fit in int32 (shard larger corpora)");
  DeviceInfo di;
  if (int rc = get_device_info(&di)) return rc;
  p->dtype = dtype;
  p->k = k;
  p->f32_tc = dtype == DRS_F32 && allow_f32_tc && g_opt.fp32_mode == 0 && di.cc_major == 10 && dim % 4 == 0;
  p->f32_prepared = false;
  p->f32_a_bytes = p->f32_b_bytes = 0;
  if (is_16bit(dtype) || p->f32_tc) {
    if (di.cc_major != 10) return fail(DRS_ERR_UNSUPPORTED, "the bf16 path needs an sm_100 device (tcgen05/TMEM); this is sm_%d%d", di.cc_major, di.cc_minor);
    if (is_16bit(dtype) && dim % 8 != 0) return fail(DRS_ERR_INVALID, "bf16 path: dim must be a multiple of 8 (TMA 16-byte row pitch), got %d", dim);
    // CTA pair (256-row A tiles) for large batches; single CTAs (128-row tiles) up to 128 claims, and near the
    // ridge of the two roofs (<= 1024 claims) whenever they pad the batch less: 384 claims are 3 x 128 exactly
    // but 2 x 256 with a quarter of every MMA wasted (15.3 vs 16.9 ms over 25M rows, 3.5 vs 4.4 ms over 6.25M).
    // With equal padding (192, 256, 512 claims) the pair is the better choice on the big corpus (less L2 -> SM
    // traffic per flop): 9.1 / 9.6 / 18.1 ms against 10.2 / 11.9 / 20.9 ms over 25M rows.
    int cg = 2;
    if (nq <= 128) cg = 1;
    else if (nq <= 1024 && (nq + 127) / 128 * 128 < (nq + 255) / 256 * 256) cg = 1;
    if (g_opt.cta_group) cg = g_opt.cta_group;
    if (cg != 1 && cg != 2) return fail(DRS_ERR_INVALID, "search.cta_group must be 0, 1 or 2");
    int ctas = g_opt.num_ctas > 0 ? g_opt.num_ctas : di.num_sms;
    ctas = std::max(cg, ctas - ctas % cg);
    p->cg = cg;
    p->grid = ctas;
    p->f32_bn = (p->f32_tc && g_opt.fp32_tile != 256) ? 128 : 256;
    p->shape = plan_shape(nq, nc, p->f32_tc ? (dim + 31) / 32 : (dim + 63) / 64, 128 * cg, p->f32_tc ? p->f32_bn : 256, ctas / cg, g_opt.splits, kTcColGroups);
    p->shape.f16_operands = dtype == DRS_F16;
    p->shape.m_block = g_opt.m_block < 0 ? 0 : (g_opt.m_block > 0 ? g_opt.m_block : ctas / cg);   // A super-blocks (gemm_tc.cuh::unit_to_tile)
  } else if (dtype == DRS_F32) {
    int ctas = g_opt.num_ctas > 0 ? g_opt.num_ctas : 2 * di.num_sms;
    p->cg = 1;
    p->grid = ctas;
    p->shape = plan_shape(nq, nc, 0, 128, 128, ctas, g_opt.splits, 1);
  } else {
    return fail(DRS_ERR_INVALID, "unknown dtype %d", dtype);
  }
  // List capacity per (claim, slot) and the pass count.  k <= 16: lists of 16, one pass.  Larger k: adaptive
  // passes (merge.cuh) -- each pass is certain of >= kcap picks, so at most ceil(k / kcap) are enqueued and,
  // when a claim's neighbours are spread over the slots, the first one finishes.  Lists of 16 cost about
  // half of what lists of 32 cost in the epilogue (registers, insert depth), so they are used whenever the
  // expected share of a slot, k / slots, is small (<= 4: P(> 16 in one slot) ~ 1e-6 for spread neighbours;
  // clustered neighbours only add passes, never errors).  17 <= k <= 32 with few slots: lists of 32, one pass.
  {
    const int slots = num_slots(p->shape);
    p->kcap = (k <= 16 || k <= 4 * slots) ? 16 : 32;
    p->passes = (k + p->kcap - 1) / p->kcap;
  }
  // passes > 1: per claim a continuation bound (u64) and the count of picks already emitted (int)
  p->bound_bytes = p->passes > 1 ? align256s(static_cast<size_t>(nq) * sizeof(uint64_t)) + align256s(static_cast<size_t>(nq) * sizeof(int)) : 0;
  p->cand_bytes = align256s(static_cast<size_t>(nq) * num_slots(p->shape) * p->kcap * sizeof(uint64_t));
  p->pad_bytes = 0;
  p->a_rows = nq;
  if (p->f32_tc) {
    // the claims are always staged (hi and lo copies, zero-padded to whole A tiles); the corpus keeps its own
    // words as the hi part and gets a residual array
    const int64_t tile = 128 * p->cg;
    p->a_rows = (nq + tile - 1) / tile * tile;
    p->f32_a_bytes = align256s(static_cast<size_t>(p->a_rows) * dim * 4);
    p->f32_b_bytes = align256s(static_cast<size_t>(nc) * dim * 4);
  }
  if (is_16bit(dtype)) {
    // TMA boxes that hang over the end of the claims matrix are zero-filled correctly but SLOWLY (measured:
    // 1 claim in a 128-row box streams the corpus at 4.7 TB/s, the same claim zero-padded in memory at
    // 6.7 TB/s), and with the round barrier one slow A tile holds every cluster back.  Stage a padded copy.
    const int64_t tile = 128 * p->cg;
    const int64_t padded = (nq + tile - 1) / tile * tile;
    if (padded != nq) {
      p->a_rows = padded;
      p->pad_bytes = align256s(static_cast<size_t>(padded) * dim * 2);
    }
  }
  {
    const int k_pass = p->passes > 1 ? p->kcap : k;   // entries a unit's list is asked for
    p->seed_chunks = std::min(5, k_pass);
    p->seed_group = k_pass / p->seed_chunks;
    p->seed_slots = (k + p->seed_group - 1) / p->seed_group;
  }
  if (p->seed_slots > drs::kMaxSeedSlots) return fail(DRS_ERR_UNSUPPORTED, "internal: %d seed slots exceed the limit", p->seed_slots);
  p->seed_bytes = align256s(static_cast<size_t>(nq) * (p->seed_slots + 1) * sizeof(uint32_t));   // + the floor word
  p->ws_bytes = kWsHeaderBytes + p->bound_bytes + p->cand_bytes + p->pad_bytes + p->seed_bytes + 2 * p->f32_a_bytes + p->f32_b_bytes;
  return DRS_OK;
}

// ------------------------------------------------------------------ generic GEMM launchers
// Launch contract for the round barrier (gemm_tc.cuh): the producers spin on a grid-wide counter, which is only safe
// when every CTA of the grid is resident at the same time.  That is not inferred from "grid <= SM count": the
// occupancy API is asked how many clusters of this kernel the device holds (it knows about GPC shapes, MIG slices
// and carve-outs), and the kernel is launched COOPERATIVELY, so the driver either places the whole grid at once --
// waiting for a kernel on another stream to leave if it must -- or refuses the launch.  If either check fails the
// scan runs without the barrier (correct, more DRAM traffic) and `debug.coop_fallbacks` counts it.
// PREC = 1 (fp32 operands as hi + lo, three kind::tf32 MMAs per K step): a / b are the hi parts -- any fp32 matrix,
// the tensor core reads the top 19 bits of every word -- and a_lo / b_lo the matching residuals (split_f32_kernel).
template <class Epi, class = void> struct epi_has_tma_output : std::false_type {};
template <class Epi> struct epi_has_tma_output<Epi, std::enable_if_t<Epi::kTmaOutput>> : std::true_type {};

template <int CG, class Epi, int BN = 256, int PREC = 0>
int launch_gemm_tc(const void* a, const void* b, int kdim, drs::GemmShape shp, int grid,
                   const typename Epi::Params& ep, cudaStream_t st, int64_t a_rows_alloc = 0, int64_t pitch_a = 0,
                   int64_t pitch_b = 0, const void* a_lo = nullptr, const void* b_lo = nullptr) {
  using Cfg = drs::GemmCfg<CG, BN, PREC>;
  CUtensorMap ta, tb, ta_lo, tb_lo;
  const int64_t rows_a_map = a_rows_alloc > 0 ? a_rows_alloc : shp.rows_a;
  if constexpr (PREC == 0) {
    if (int rc = make_tmap_bf16(&ta, a, rows_a_map, kdim, Cfg::BM, shp.f16_operands != 0, pitch_a)) return rc;
    if (int rc = make_tmap_bf16(&tb, b, shp.rows_b, kdim, Cfg::BN_CTA, shp.f16_operands != 0, pitch_b)) return rc;
    ta_lo = ta;
    tb_lo = tb;
    if (shp.epi_tma_store) {   // the epilogue's output matrix (GradLogitEpilogue: ep.out, rows_a x rows_b, pitch ld_out)
      if constexpr (epi_has_tma_output<Epi>::value) {
        if (int rc = make_tmap_out_bf16(&tb_lo, ep.out, ep.rows_a, ep.rows_b, ep.ld_out)) return rc;
      } else {
        return fail(DRS_ERR_INVALID, "internal: epi_tma_store needs an epilogue with an output matrix");
      }
    }
    if (shp.a_sym) {   // the transposed view of the (square, symmetric) A operand: 64 x 64 boxes of the same matrix
      if (CG != 2 || shp.rows_a != kdim) return fail(DRS_ERR_INVALID, "internal: a_sym needs the CTA-pair kernel and a square A");
      if (int rc = make_tmap_bf16(&ta_lo, a, rows_a_map, kdim, 64, shp.f16_operands != 0, pitch_a)) return rc;
    }
  } else {
    if (!a_lo || !b_lo) return fail(DRS_ERR_INVALID, "fp32 tensor-core GEMM needs the lo parts of both operands");
    if (int rc = make_tmap_f32(&ta, a, rows_a_map, kdim, Cfg::BM)) return rc;
    if (int rc = make_tmap_f32(&tb, b, shp.rows_b, kdim, Cfg::BN_CTA)) return rc;
    if (int rc = make_tmap_f32(&ta_lo, a_lo, rows_a_map, kdim, Cfg::BM)) return rc;
    if (int rc = make_tmap_f32(&tb_lo, b_lo, shp.rows_b, kdim, Cfg::BN_CTA)) return rc;
  }
  if (shp.k_splits > 1 && (shp.a_sym || shp.skip_below_diagonal || shp.round_counter != nullptr))
    return fail(DRS_ERR_INVALID, "internal: split-K is not combined with the symmetric schedules or the round barrier");
  auto kern = drs::gemm_nt_tc_kernel<CG, Epi, BN, PREC>;
  static thread_local bool attr_set[64] = {};
  static thread_local int max_clusters[64] = {};
  int dev = 0;
  DRS_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(DRS_ERR_INVALID, "device ordinal %d out of range", dev);
  if (!attr_set[dev]) {
    DRS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set[dev] = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(Cfg::THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (shp.round_counter != nullptr) {
    if (max_clusters[dev] == 0) {
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { n = -1; (void)cudaGetLastError(); }
      max_clusters[dev] = n == 0 ? -1 : n;
    }
    if (max_clusters[dev] * CG < grid) {   // not all CTAs can be resident at once: no spinning on a grid-wide counter
      shp.round_counter = nullptr;
      ++g_opt.coop_fallbacks;
    } else if (g_opt.cooperative) {
      attr[1].id = cudaLaunchAttributeCooperative;
      attr[1].val.cooperative = 1;
      cfg.numAttrs = 2;
    }
  }
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, ta_lo, tb_lo, shp, ep);
  if (e != cudaSuccess && cfg.numAttrs == 2 &&
      (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorNotSupported || e == cudaErrorInvalidValue ||
       e == cudaErrorInvalidConfiguration)) {
    (void)cudaGetLastError();              // the driver refused to co-schedule the grid: run without the barrier
    shp.round_counter = nullptr;
    cfg.numAttrs = 1;
    ++g_opt.coop_fallbacks;
    e = cudaLaunchKernelEx(&cfg, kern, ta, tb, ta_lo, tb_lo, shp, ep);
  }
  DRS_CUDA(e);
  return DRS_OK;
}
template <class Epi, int BN = 256>
int launch_gemm_tc_cg(int cg, const void* a, const void* b, int kdim, const drs::GemmShape& shp, int grid,
                      const typename Epi::Params& ep, cudaStream_t st, int64_t pitch_a = 0, int64_t pitch_b = 0) {
  return cg == 2 ? launch_gemm_tc<2, Epi, BN>(a, b, kdim, shp, grid, ep, st, 0, pitch_a, pitch_b)
                 : launch_gemm_tc<1, Epi, BN>(a, b, kdim, shp, grid, ep, st, 0, pitch_a, pitch_b);
}
template <class Epi, bool B_KN>
int launch_gemm_simt(const float* a, long long lda, const float* b, long long ldb, int kdim,
                     const drs::GemmShape& shp, int grid, const typename Epi::Params& ep, cudaStream_t st) {
  auto kern = drs::gemm_simt_f32_kernel<Epi, B_KN>;
  DRS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, drs::SimtCfg::SMEM_BYTES));
  kern<<<grid, drs::SimtCfg::THREADS, drs::SimtCfg::SMEM_BYTES, st>>>(a, b, lda, ldb, kdim, shp, ep);
  DRS_CUDA(cudaGetLastError());
  return DRS_OK;
}

template <int CG, int KCAP, int PREC = 0, int BN = 256>
int launch_search_tc(const SearchPlan& p, const void* queries, const void* corpus, int dim, uint64_t* ws, int k_pass,
                     const uint64_t* bound, const float* col_bias, uint32_t* seeds, cudaStream_t st,
                     const void* queries_lo = nullptr, const void* corpus_lo = nullptr) {
  if (col_bias != nullptr) {
    using Epi = drs::TopKEpilogue<KCAP, true>;
    typename Epi::Params ep{ws, p.shape.rows_a, p.shape.rows_b, num_slots(p.shape), k_pass, bound, col_bias, 2.0f, seeds, p.seed_slots, p.seed_group, p.seed_chunks};
    return launch_gemm_tc<CG, Epi, BN, PREC>(queries, corpus, dim, p.shape, p.grid, ep, st, p.a_rows, 0, 0, queries_lo, corpus_lo);
  }
  using Epi = drs::TopKEpilogue<KCAP>;
  typename Epi::Params ep{ws, p.shape.rows_a, p.shape.rows_b, num_slots(p.shape), k_pass, bound, nullptr, 1.0f, seeds, p.seed_slots, p.seed_group, p.seed_chunks};
  return launch_gemm_tc<CG, Epi, BN, PREC>(queries, corpus, dim, p.shape, p.grid, ep, st, p.a_rows, 0, 0, queries_lo, corpus_lo);
}

template <int KCAP>
int launch_search_f32(const SearchPlan& p, const void* queries, const void* corpus, int dim, uint64_t* ws, int k_pass,
                      const uint64_t* bound, const float* col_bias, uint32_t* seeds, cudaStream_t st) {
  if (col_bias != nullptr) {
    using Epi = drs::TopKEpilogue<KCAP, true>;
    typename Epi::Params ep{ws, p.shape.rows_a, p.shape.rows_b, num_slots(p.shape), k_pass, bound, col_bias, 2.0f, seeds, p.seed_slots, p.seed_group, p.seed_chunks};
    return launch_gemm_simt<Epi, false>(static_cast<const float*>(queries), dim, static_cast<const float*>(corpus),
                                        dim, dim, p.shape, p.grid, ep, st);
  }
  using Epi = drs::TopKEpilogue<KCAP>;
  typename Epi::Params ep{ws, p.shape.rows_a, p.shape.rows_b, num_slots(p.shape), k_pass, bound, nullptr, 1.0f, seeds, p.seed_slots, p.seed_group, p.seed_chunks};
  return launch_gemm_simt<Epi, false>(static_cast<const float*>(queries), dim, static_cast<const float*>(corpus), dim,
                                      dim, p.shape, p.grid, ep, st);
}

int launch_merge_keys(const uint64_t* ws, int64_t nq, int ncand, int k, int64_t id_base, float* out_scores,
                      int64_t* out_ids, int ld_out, uint64_t* bound_out, const float* row_term, cudaStream_t st) {
  const int warps_per_block = 4;
  const int blocks = static_cast<int>((nq + warps_per_block - 1) / warps_per_block);
  long long* ids = reinterpret_cast<long long*>(out_ids);
  if (ncand <= 32 * 8)
    drs::merge_keys_kernel<8><<<blocks, 128, 0, st>>>(ws, (int)nq, ncand, k, id_base, out_scores, ids, ld_out, bound_out, row_term);
  else if (ncand <= 32 * 24)
    drs::merge_keys_kernel<24><<<blocks, 128, 0, st>>>(ws, (int)nq, ncand, k, id_base, out_scores, ids, ld_out, bound_out, row_term);
  else if (ncand <= 32 * 40)
    drs::merge_keys_kernel<40><<<blocks, 128, 0, st>>>(ws, (int)nq, ncand, k, id_base, out_scores, ids, ld_out, bound_out, row_term);
  else
    drs::merge_keys_kernel<0><<<blocks, 128, 0, st>>>(ws, (int)nq, ncand, k, id_base, out_scores, ids, ld_out, bound_out, row_term);
  DRS_CUDA(cudaGetLastError());
  return DRS_OK;
}

// single-pass select: k-way merge of the sorted per-slot runs; unsorted re-scan only beyond 320 slots
int launch_select(const uint64_t* ws, int64_t nq, int nslots, int kcap, int k, int64_t id_base, float* out_scores,
                  int64_t* out_ids, const float* row_term, cudaStream_t st) {
  const int blocks = static_cast<int>((nq + 3) / 4);
  long long* ids = reinterpret_cast<long long*>(out_ids);
#define DRS_SEL(SL) drs::select_runs_kernel<SL><<<blocks, 128, 0, st>>>(ws, (int)nq, nslots, kcap, k, id_base, out_scores, ids, row_term)
  if (nslots <= 32) DRS_SEL(1);
  else if (nslots <= 64) DRS_SEL(2);
  else if (nslots <= 96) DRS_SEL(3);
  else if (nslots <= 160) DRS_SEL(5);
  else if (nslots <= 320) DRS_SEL(10);
  else return launch_merge_keys(ws, nq, nslots * kcap, k, id_base, out_scores, out_ids, k, nullptr, row_term, st);
#undef DRS_SEL
  DRS_CUDA(cudaGetLastError());
  return DRS_OK;
}

template <int SL>
int launch_merge_runs_sl(const uint64_t* ws, int64_t nq, int nslots, int kcap, int k, int64_t id_base, float* out_scores,
                         int64_t* out_ids, uint64_t* bound, int* done, const unsigned int* active_in,
                         unsigned int* remaining_out, const float* row_term, cudaStream_t st) {
  const int blocks = static_cast<int>((nq + 3) / 4);
  drs::merge_runs_kernel<SL><<<blocks, 128, 0, st>>>(ws, (int)nq, nslots, kcap, k, id_base, out_scores,
                                                          reinterpret_cast<long long*>(out_ids), k, bound, done,
                                                          active_in, remaining_out, row_term);
  DRS_CUDA(cudaGetLastError());
  return DRS_OK;
}
int launch_merge_runs(const uint64_t* ws, int64_t nq, int nslots, int kcap, int k, int64_t id_base, float* out_scores,
                      int64_t* out_ids, uint64_t* bound, int* done, const unsigned int* active_in,
                      unsigned int* remaining_out, const float* row_term, cudaStream_t st) {
#define DRS_RUNS(SL) return launch_merge_runs_sl<SL>(ws, nq, nslots, kcap, k, id_base, out_scores, out_ids, bound, done, active_in, remaining_out, row_term, st)
  if (nslots <= 32) DRS_RUNS(1);
  if (nslots <= 64) DRS_RUNS(2);
  if (nslots <= 96) DRS_RUNS(3);
  if (nslots <= 160) DRS_RUNS(5);
  if (nslots <= 320) DRS_RUNS(10);
#undef DRS_RUNS
  return fail(DRS_ERR_UNSUPPORTED, "k > 32 with %d corpus splits per claim (limit 320)", nslots);
}

}  // namespace

extern "C" {

int drs_version(void) { return 100; }
const char* drs_last_error(void) { return g_err; }

int drs_debug_hang_report(unsigned int out[6]) {
  if (!out) return fail(DRS_ERR_INVALID, "out is null");
  memset(out, 0, 6 * sizeof(unsigned int));
  if (g_hang_host) memcpy(out, g_hang_host, sizeof(drs::HangReport));
  return DRS_OK;
}

int drs_debug_triangle_walk(int tiles, int parts, int part, int order, int* out_mt, int capacity, int* count) {
  if (tiles <= 0 || parts <= 0 || part < 0 || part >= parts || !count || (capacity > 0 && !out_mt))
    return fail(DRS_ERR_INVALID, "drs_debug_triangle_walk: bad arguments");
  int n = 0;
  for (drs::TriangleWalk w(tiles, static_cast<unsigned>(part), static_cast<unsigned>(parts), order); w.valid(); w.next()) {
    if (n < capacity) { out_mt[2 * n] = w.m; out_mt[2 * n + 1] = w.t; }
    ++n;
  }
  *count = n;
  return DRS_OK;
}

int drs_debug_max_clusters(int cluster_size, int* out) {
  if (!out || cluster_size < 1 || cluster_size > 16) return fail(DRS_ERR_INVALID, "bad argument");
  using Cfg = drs::GemmCfg<2>;
  auto kern = drs::gemm_nt_tc_kernel<2, drs::TopKEpilogue<16>>;
  DRS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  if (cluster_size > 8) DRS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(cluster_size * 64);
  cfg.blockDim = dim3(Cfg::THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_size;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  DRS_CUDA(cudaOccupancyMaxActiveClusters(out, kern, &cfg));
  return DRS_OK;
}

int drs_set_option(const char* name, int value) {
  if (!name) return fail(DRS_ERR_INVALID, "null option name");
  init_options_from_env();
  if (!strcmp(name, "search.cta_group")) g_opt.cta_group = value;
  else if (!strcmp(name, "search.num_ctas")) g_opt.num_ctas = value;
  else if (!strcmp(name, "search.splits")) g_opt.splits = value;
  else if (!strcmp(name, "infonce.cta_group")) g_opt.infonce_cta_group = value;
  else if (!strcmp(name, "debug.flags")) g_opt.debug_flags = value;
  else if (!strcmp(name, "tune.b_hint")) g_opt.b_hint = value;
  else if (!strcmp(name, "tune.stagger_cycles")) g_opt.stagger_cycles = value;
  else if (!strcmp(name, "tune.round_barrier")) g_opt.round_barrier = value;
  else if (!strcmp(name, "tune.seed_thresholds")) g_opt.seed_thresholds = value;
  else if (!strcmp(name, "tune.symmetric_grad")) g_opt.symmetric_grad = value;
  else if (!strcmp(name, "tune.symmetric_lse")) g_opt.symmetric_lse = value;
  else if (!strcmp(name, "tune.triangle_order")) g_opt.triangle_order = value;
  else if (!strcmp(name, "tune.k_split")) g_opt.k_split = value;
  else if (!strcmp(name, "tune.cooperative")) g_opt.cooperative = value;
  else if (!strcmp(name, "search.fp32_mode")) g_opt.fp32_mode = value;
  else if (!strcmp(name, "infonce.df_tile")) g_opt.df_tile = value;
  else if (!strcmp(name, "search.m_block")) g_opt.m_block = value;
  else if (!strcmp(name, "search.fp32_tile")) g_opt.fp32_tile = value;
  else if (!strcmp(name, "infonce.tma_store")) g_opt.tma_store = value;
  else if (!strcmp(name, "debug.coop_fallbacks")) g_opt.coop_fallbacks = value;
  else return fail(DRS_ERR_INVALID, "unknown option '%s'", name);
  return DRS_OK;
}
int drs_get_option(const char* name, int* value) {
  if (!name || !value) return fail(DRS_ERR_INVALID, "null argument");
  if (!strcmp(name, "search.cta_group")) *value = g_opt.cta_group;
  else if (!strcmp(name, "search.num_ctas")) *value = g_opt.num_ctas;
  else if (!strcmp(name, "search.splits")) *value = g_opt.splits;
  else if (!strcmp(name, "infonce.cta_group")) *value = g_opt.infonce_cta_group;
  else if (!strcmp(name, "debug.flags")) *value = g_opt.debug_flags;
  else if (!strcmp(name, "tune.b_hint")) *value = g_opt.b_hint;
  else if (!strcmp(name, "tune.stagger_cycles")) *value = g_opt.stagger_cycles;
  else if (!strcmp(name, "tune.round_barrier")) *value = g_opt.round_barrier;
  else if (!strcmp(name, "tune.seed_thresholds")) *value = g_opt.seed_thresholds;
  else if (!strcmp(name, "tune.symmetric_grad")) *value = g_opt.symmetric_grad;
  else if (!strcmp(name, "tune.symmetric_lse")) *value = g_opt.symmetric_lse;
  else if (!strcmp(name, "tune.triangle_order")) *value = g_opt.triangle_order;
  else if (!strcmp(name, "tune.k_split")) *value = g_opt.k_split;
  else if (!strcmp(name, "tune.cooperative")) *value = g_opt.cooperative;
  else if (!strcmp(name, "search.fp32_mode")) *value = g_opt.fp32_mode;
  else if (!strcmp(name, "infonce.df_tile")) *value = g_opt.df_tile;
  else if (!strcmp(name, "search.m_block")) *value = g_opt.m_block;
  else if (!strcmp(name, "search.fp32_tile")) *value = g_opt.fp32_tile;
  else if (!strcmp(name, "infonce.tma_store")) *value = g_opt.tma_store;
  else if (!strcmp(name, "debug.coop_fallbacks")) *value = g_opt.coop_fallbacks;
  else return fail(DRS_ERR_INVALID, "unknown option '%s'", name);
  return DRS_OK;
}

int drs_search_workspace_bytes(int64_t nq, int64_t nc, int dim, int k, int dtype, size_t* bytes) {
  if (!bytes) return fail(DRS_ERR_INVALID, "bytes is null");
  SearchPlan p;
  if (int rc = plan_search(nq, nc, dim, k, dtype, &p)) return rc;
  *bytes = p.ws_bytes;
  return DRS_OK;
}

namespace {
// one pass: candidates for the k_pass best keys below bound[row] (bound == nullptr: no bound)
int scan_pass(SearchPlan& p, const void* queries, const void* corpus, int dim, void* workspace, int k_pass,
              const uint64_t* bound, cudaStream_t st, const float* col_bias = nullptr, size_t extra_bytes = 0,
              const unsigned int* active = nullptr) {
  p.shape.active = active;
  char* base = static_cast<char*>(workspace);
  uint32_t* seeds = nullptr;
  // seeds pay only when a claim's units run one after another: with a single round (units <= groups, e.g. one A
  // tile split 148 ways) nobody ever reads them, and 296 lists publishing for the same 128 claims at the same
  // moment contend on the compare-and-swap (measured +100 us on a 370 us scan)
  const bool tc = is_16bit(p.dtype) || p.f32_tc;
  const int groups = tc ? p.grid / p.cg : p.grid;
  if (g_opt.seed_thresholds && p.shape.num_splits > 1 && p.shape.num_m_tiles * p.shape.num_splits > groups)
    seeds = reinterpret_cast<uint32_t*>(base + kWsHeaderBytes + p.bound_bytes + extra_bytes + p.cand_bytes + p.pad_bytes);
  uint64_t* ws = reinterpret_cast<uint64_t*>(base + kWsHeaderBytes + p.bound_bytes + extra_bytes);
  p.shape.round_counter = nullptr;
  char* pad = nullptr;
  if (tc) {
    DeviceInfo di;
    if (int rc = get_device_info(&di)) return rc;
    // requested here -- only when some cluster runs more than one unit, else nobody would ever wait on it -- and
    // granted by launch_gemm_tc only if the whole grid is co-resident (occupancy query + cooperative launch)
    if (g_opt.round_barrier && p.grid <= di.num_sms && p.shape.num_m_tiles * p.shape.num_splits > groups)
      p.shape.round_counter = static_cast<unsigned int*>(workspace);
    if (p.pad_bytes) pad = reinterpret_cast<char*>(ws) + p.cand_bytes;  // zero-padded claims (plan_search explains why)
  }
  // one staging launch: round counter, seeds, padded claims (merge.cuh::scan_prep_kernel)
  if (p.shape.round_counter != nullptr || seeds != nullptr || pad != nullptr) {
    const size_t seed_vec = seeds ? p.seed_bytes / 16 : 0;
    const size_t live = pad ? static_cast<size_t>(p.shape.rows_a) * dim * 2 : 0;   // dim % 8 == 0: a multiple of 16
    const size_t pad_vec = pad ? static_cast<size_t>(p.a_rows) * dim * 2 / 16 : 0;
    const size_t work = std::max<size_t>(std::max(seed_vec, pad_vec), 16);
    const int blocks = static_cast<int>(std::min<size_t>((work + 255) / 256, 2048));
    drs::scan_prep_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<uint4*>(p.shape.round_counter), reinterpret_cast<uint4*>(seeds),
                                                  seed_vec, static_cast<const uint4*>(queries), reinterpret_cast<uint4*>(pad),
                                                  live / 16, pad_vec);
    DRS_CUDA(cudaGetLastError());
    if (pad) queries = pad;
  }
  if (p.f32_tc) {
    // operands as hi + lo: [.. seeds | claims hi | claims lo | corpus lo]; split once per call (every pass reuses it)
    char* f32 = reinterpret_cast<char*>(ws) + p.cand_bytes + p.pad_bytes + p.seed_bytes;
    float* a_hi = reinterpret_cast<float*>(f32);
    float* a_lo = reinterpret_cast<float*>(f32 + p.f32_a_bytes);
    float* b_lo = reinterpret_cast<float*>(f32 + 2 * p.f32_a_bytes);
    if (!p.f32_prepared) {
      const size_t na = static_cast<size_t>(p.a_rows) * dim, live = static_cast<size_t>(p.shape.rows_a) * dim;
      const size_t nb = static_cast<size_t>(p.shape.rows_b) * dim;
      drs::split_f32_kernel<<<static_cast<int>(std::min<size_t>((na / 4 + 255) / 256, 4096)), 256, 0, st>>>(
          static_cast<const float4*>(queries), live / 4, na / 4, reinterpret_cast<float4*>(a_hi), reinterpret_cast<float4*>(a_lo));
      drs::split_f32_kernel<<<static_cast<int>(std::min<size_t>((nb / 4 + 255) / 256, 8192)), 256, 0, st>>>(
          static_cast<const float4*>(corpus), nb / 4, nb / 4, nullptr, reinterpret_cast<float4*>(b_lo));
      DRS_CUDA(cudaGetLastError());
      p.f32_prepared = true;
    }
#define DRS_F32TC(CGV, KC, BNV) launch_search_tc<CGV, KC, 1, BNV>(p, a_hi, corpus, dim, ws, k_pass, bound, col_bias, seeds, st, a_lo, b_lo)
    if (p.f32_bn == 128) {   // 128-wide tiles: two accumulator stages (main + correction each) fit TMEM, the epilogue overlaps the MMAs
      if (p.cg == 1) return p.kcap == 16 ? DRS_F32TC(1, 16, 128) : DRS_F32TC(1, 32, 128);
      return p.kcap == 16 ? DRS_F32TC(2, 16, 128) : DRS_F32TC(2, 32, 128);
    }
    if (p.cg == 1) return p.kcap == 16 ? DRS_F32TC(1, 16, 256) : DRS_F32TC(1, 32, 256);
    return p.kcap == 16 ? DRS_F32TC(2, 16, 256) : DRS_F32TC(2, 32, 256);
#undef DRS_F32TC
  }
  if (is_16bit(p.dtype)) {
    if (p.cg == 1) return p.kcap == 16 ? launch_search_tc<1, 16>(p, queries, corpus, dim, ws, k_pass, bound, col_bias, seeds, st)
                                       : launch_search_tc<1, 32>(p, queries, corpus, dim, ws, k_pass, bound, col_bias, seeds, st);
    return p.kcap == 16 ? launch_search_tc<2, 16>(p, queries, corpus, dim, ws, k_pass, bound, col_bias, seeds, st)
                        : launch_search_tc<2, 32>(p, queries, corpus, dim, ws, k_pass, bound, col_bias, seeds, st);
  }
  return p.kcap == 16 ? launch_search_f32<16>(p, queries, corpus, dim, ws, k_pass, bound, col_bias, seeds, st)
                      : launch_search_f32<32>(p, queries, corpus, dim, ws, k_pass, bound, col_bias, seeds, st);
}
int check_search_args(const SearchPlan& p, const void* queries, const void* corpus, void* workspace,
                      size_t workspace_bytes) {
  if (!queries || !corpus) return fail(DRS_ERR_INVALID, "null pointer argument");
  if (!workspace || workspace_bytes < p.ws_bytes)
    return fail(DRS_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", p.ws_bytes, workspace_bytes);
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return fail(DRS_ERR_INVALID, "workspace must be 256-byte aligned");
  if ((is_16bit(p.dtype) || p.f32_tc) && ((reinterpret_cast<uintptr_t>(queries) & 15) || (reinterpret_cast<uintptr_t>(corpus) & 15)))
    return fail(DRS_ERR_INVALID, "tensor-core path: queries and corpus must be 16-byte aligned");
  return DRS_OK;
}
}  // namespace

int drs_search_scan(const void* queries, int64_t nq, const void* corpus, int64_t nc, int dim, int dtype, int k,
                    void* workspace, size_t workspace_bytes, void* stream) {
  SearchPlan p;
  if (int rc = plan_search(nq, nc, dim, k, dtype, &p)) return rc;
  if (p.passes > 1) return fail(DRS_ERR_UNSUPPORTED, "drs_search_scan/select serve single-pass searches (k <= 16, or <= 32 on small corpora); use drs_search for k = %d", k);
  if (int rc = check_search_args(p, queries, corpus, workspace, workspace_bytes)) return rc;
  return scan_pass(p, queries, corpus, dim, workspace, k, nullptr, static_cast<cudaStream_t>(stream));
}

int drs_search_select(const void* workspace, int64_t nq, int64_t nc, int dim, int dtype, int k, int64_t id_base,
                      float* out_scores, int64_t* out_ids, void* stream) {
  if (!workspace || !out_scores || !out_ids) return fail(DRS_ERR_INVALID, "null pointer argument");
  SearchPlan p;
  if (int rc = plan_search(nq, nc, dim, k, dtype, &p)) return rc;
  if (p.passes > 1) return fail(DRS_ERR_UNSUPPORTED, "drs_search_scan/select serve single-pass searches (k <= 16, or <= 32 on small corpora); use drs_search for k = %d", k);
  return launch_select(reinterpret_cast<const uint64_t*>(static_cast<const char*>(workspace) + kWsHeaderBytes), nq,
                       num_slots(p.shape), p.kcap, k, id_base, out_scores, out_ids, nullptr, static_cast<cudaStream_t>(stream));
}

namespace {
// Scan + select, shared by the dot-product and the squared-L2 searches.
//   k <= 32: one scan (register lists of 16 or 32 per claim and split) + one select.
//   k  > 32: adaptive passes (merge.cuh::merge_runs_kernel): all ceil(k/32) passes are enqueued, a pass
//            whose predecessor left no claim open returns at once on the device (no host sync).
int run_search(SearchPlan& p, const void* queries, const void* corpus, int dim, int64_t nq, int k, int64_t id_base,
               float* out_vals, int64_t* out_ids, void* workspace, const float* col_bias, const float* row_term,
               size_t extra_bytes, cudaStream_t st) {
  char* base = static_cast<char*>(workspace);
  const uint64_t* cand = reinterpret_cast<const uint64_t*>(base + kWsHeaderBytes + p.bound_bytes + extra_bytes);
  if (p.passes == 1) {
    if (int rc = scan_pass(p, queries, corpus, dim, workspace, k, nullptr, st, col_bias, extra_bytes)) return rc;
    return launch_select(cand, nq, num_slots(p.shape), p.kcap, k, id_base, out_vals, out_ids, row_term, st);
  }
  uint64_t* bound = reinterpret_cast<uint64_t*>(base + kWsHeaderBytes);
  int* done = reinterpret_cast<int*>(base + kWsHeaderBytes + align256s(static_cast<size_t>(nq) * sizeof(uint64_t)));
  unsigned int* open_claims = reinterpret_cast<unsigned int*>(base + kWsRoundBytes);  // [pass]
  DRS_CUDA(cudaMemsetAsync(open_claims, 0, kWsHeaderBytes - kWsRoundBytes, st));
  DRS_CUDA(cudaMemsetAsync(bound, 0xFF, static_cast<size_t>(nq) * sizeof(uint64_t), st));
  DRS_CUDA(cudaMemsetAsync(done, 0, static_cast<size_t>(nq) * sizeof(int), st));
  for (int pass = 0; pass < p.passes; ++pass) {
    const unsigned int* active = pass ? open_claims + (pass - 1) : nullptr;
    if (int rc = scan_pass(p, queries, corpus, dim, workspace, p.kcap, pass ? bound : nullptr, st, col_bias, extra_bytes, active)) return rc;
    if (int rc = launch_merge_runs(cand, nq, num_slots(p.shape), p.kcap, k, id_base, out_vals, out_ids, bound, done, active,
                                   open_claims + pass, row_term, st)) return rc;
  }
  return DRS_OK;
}
}  // namespace

int drs_search(const void* queries, int64_t nq, const void* corpus, int64_t nc, int dim, int dtype, int k,
               int64_t id_base, float* out_scores, int64_t* out_ids, void* workspace, size_t workspace_bytes,
               void* stream) {
  if (!out_scores || !out_ids) return fail(DRS_ERR_INVALID, "null pointer argument");
  SearchPlan p;
  if (int rc = plan_search(nq, nc, dim, k, dtype, &p)) return rc;
  if (int rc = check_search_args(p, queries, corpus, workspace, workspace_bytes)) return rc;
  return run_search(p, queries, corpus, dim, nq, k, id_base, out_scores, out_ids, workspace, nullptr, nullptr, 0,
                    static_cast<cudaStream_t>(stream));
}

int drs_debug_open_claims(const void* workspace, unsigned int out[8], void* stream) {
  if (!workspace || !out) return fail(DRS_ERR_INVALID, "null pointer argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DRS_CUDA(cudaMemcpyAsync(out, static_cast<const char*>(workspace) + kWsRoundBytes, 8 * sizeof(unsigned int),
                           cudaMemcpyDeviceToHost, st));
  DRS_CUDA(cudaStreamSynchronize(st));
  return DRS_OK;
}

namespace {
size_t l2_extra_bytes(int64_t nq, int64_t nc) { return align256s(nc * sizeof(float)) + align256s(nq * sizeof(float)); }
}

int drs_search_l2_workspace_bytes(int64_t nq, int64_t nc, int dim, int k, int dtype, size_t* bytes) {
  if (!bytes) return fail(DRS_ERR_INVALID, "bytes is null");
  SearchPlan p;
  if (int rc = plan_search(nq, nc, dim, k, dtype, &p)) return rc;
  *bytes = p.ws_bytes + l2_extra_bytes(nq, nc);
  return DRS_OK;
}

int drs_search_l2(const void* queries, int64_t nq, const void* corpus, int64_t nc, int dim, int dtype, int k,
                  int64_t id_base, float* out_dist, int64_t* out_ids, void* workspace, size_t workspace_bytes,
                  void* stream) {
  if (!out_dist || !out_ids) return fail(DRS_ERR_INVALID, "null pointer argument");
  SearchPlan p;
  if (int rc = plan_search(nq, nc, dim, k, dtype, &p)) return rc;
  const size_t extra = l2_extra_bytes(nq, nc);
  p.ws_bytes += extra;
  if (int rc = check_search_args(p, queries, corpus, workspace, workspace_bytes)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* base = static_cast<char*>(workspace);
  // layout: [header | bound, done | -|c|^2 per corpus row | |x|^2 per query | candidates]
  float* col_bias = reinterpret_cast<float*>(base + kWsHeaderBytes + p.bound_bytes);
  float* row_term = reinterpret_cast<float*>(base + kWsHeaderBytes + p.bound_bytes + align256s(nc * sizeof(float)));
  const int bc = static_cast<int>((nc + 7) / 8), bq = static_cast<int>((nq + 7) / 8);
  if (dtype == DRS_F16) {
    drs::row_sqnorm_kernel<__half><<<bc, 256, 0, st>>>(static_cast<const __half*>(corpus), nc, dim, -1.f, col_bias);
    drs::row_sqnorm_kernel<__half><<<bq, 256, 0, st>>>(static_cast<const __half*>(queries), nq, dim, 1.f, row_term);
  } else if (dtype == DRS_BF16) {
    drs::row_sqnorm_kernel<__nv_bfloat16><<<bc, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(corpus), nc, dim, -1.f, col_bias);
    drs::row_sqnorm_kernel<__nv_bfloat16><<<bq, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(queries), nq, dim, 1.f, row_term);
  } else {
    drs::row_sqnorm_kernel<float><<<bc, 256, 0, st>>>(static_cast<const float*>(corpus), nc, dim, -1.f, col_bias);
    drs::row_sqnorm_kernel<float><<<bq, 256, 0, st>>>(static_cast<const float*>(queries), nq, dim, 1.f, row_term);
  }
  DRS_CUDA(cudaGetLastError());
  return run_search(p, queries, corpus, dim, nq, k, id_base, out_dist, out_ids, workspace, col_bias, row_term, extra, st);
}

int drs_merge_shards(const float* scores, const int64_t* ids, int num_shards, int64_t nq, int k, float* out_scores,
                     int64_t* out_ids, void* stream) {
  if (!scores || !ids || !out_scores || !out_ids) return fail(DRS_ERR_INVALID, "null pointer argument");
  if (num_shards <= 0 || nq <= 0 || k <= 0) return fail(DRS_ERR_INVALID, "num_shards, nq, k must be positive");
  const int warps_per_block = 4;
  const int blocks = static_cast<int>((nq + warps_per_block - 1) / warps_per_block);
  drs::merge_pairs_kernel<<<blocks, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      scores, reinterpret_cast<const long long*>(ids), num_shards, (int)nq, k, out_scores,
      reinterpret_cast<long long*>(out_ids));
  DRS_CUDA(cudaGetLastError());
  return DRS_OK;
}

#include "infonce_api.inc"
#include "contrast_api.inc"
#include "rerank_api.inc"
#include "pairs_api.inc"
#include "exchange_api.inc"
#include "kmeans_api.inc"

}  // extern "C"
