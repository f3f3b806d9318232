// In-batch InfoNCE: epilogue functors and helper kernels.
// Replaces NCELoss._compute_info_loss (src/contrastor/contrastive_loss.py:56-93): the 2N x 2N
// logits, their masked/gathered copies (:65-85) and the softmax cross-entropy (:91-92) are never
// materialised in the forward pass; each row keeps a running (max, sum-exp) while the score tiles
// stream through the GEMM epilogue.
//
// Notation: F = cat(q, k) (:61), S = F F^T (:62), y = S * inv_T * log2(e) (log2 domain),
// pos(i) = (i + N) mod 2N (:57-58,:71), the diagonal is excluded (:65-68).
#pragma once
#include <cuda_bf16.h>

#include "ptx.cuh"
#include "topk.cuh"

namespace drs {

static constexpr float kLog2e = 1.4426950408889634f;
static constexpr float kLn2 = 0.6931471805599453f;

// 2^x on the MUFU pipe, one instruction (exp2f() wraps it in range fix-ups that double the epilogue cost;
// the arguments here are y - max <= 0 or y - lse: flushing denormal results to zero is harmless).
__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ------------------------------------------------------------------------- forward: row LSE partials
// Per (row, slot): running max m and l = sum 2^(y - m) over the slot's columns (log2 domain).
//   pos_mode 0: no positive, nothing masked           (queue operand, contrastive_loss.py:32,:79)
//   pos_mode 1: in-batch pairs: positive (i + half) mod 2*half, diagonal masked      (:57-75)
//   pos_mode 2: the diagonal IS the positive (ProtoNCE: label of row i is i, :118-119)
//   col_scale : optional per-column inverse temperature (ProtoNCE densities, :122-124)
struct LseEpilogue {
  struct Params {
    float2* part;     // [rows_a][num_slots]  (m, l)
    float* pos;       // [rows_a] positive logit in natural units, or nullptr
    int rows_a;
    int rows_b;
    int num_slots;
    int pos_mode;
    int half;
    float scale_log2;        // inv_T * log2(e)   (when col_scale == nullptr)
    float inv_t;
    const float* col_scale;  // [rows_b] or nullptr
  };
  static constexpr bool kUsesScratch = false;
  static constexpr bool kStagesColumns = false;
  float m, l;

  __device__ __forceinline__ void begin_unit(const Params&, int, int, int) {
    m = -INFINITY;
    l = 0.f;
  }
  __device__ __forceinline__ void chunk(const Params& p, int row, int col0, const uint32_t (&v)[32]) {
    const int valid = p.rows_b - col0;
    if (valid <= 0) return;
    const int diag = (p.pos_mode == 1) ? row - col0 : -1;  // column offset to skip
    int posj = -1;
    if (p.pos_mode == 1) {
      const int pc = row < p.half ? row + p.half : row - p.half;
      posj = pc - col0;
    } else if (p.pos_mode == 2) {
      posj = row - col0;
    }
    if (p.col_scale == nullptr && valid >= 32) {
      // Fast path (all but ~2 of a row's 256 chunks): no masked column, no positive, one temperature.
      // max over the raw scores (scale > 0 keeps the order), then one FFMA + one MUFU per score.
      const bool special = (static_cast<unsigned>(diag) < 32u) || (static_cast<unsigned>(posj) < 32u);
      if (!__any_sync(0xffffffffu, special)) {
        float c0 = -INFINITY, c1 = -INFINITY, c2 = -INFINITY, c3 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          c0 = fmaxf(c0, __uint_as_float(v[j + 0]));
          c1 = fmaxf(c1, __uint_as_float(v[j + 1]));
          c2 = fmaxf(c2, __uint_as_float(v[j + 2]));
          c3 = fmaxf(c3, __uint_as_float(v[j + 3]));
        }
        const float m_new = fmaxf(m, fmaxf(fmaxf(c0, c1), fmaxf(c2, c3)) * p.scale_log2);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          a0 += fast_ex2(fmaf(__uint_as_float(v[j + 0]), p.scale_log2, -m_new));
          a1 += fast_ex2(fmaf(__uint_as_float(v[j + 1]), p.scale_log2, -m_new));
          a2 += fast_ex2(fmaf(__uint_as_float(v[j + 2]), p.scale_log2, -m_new));
          a3 += fast_ex2(fmaf(__uint_as_float(v[j + 3]), p.scale_log2, -m_new));
        }
        l = l * fast_ex2(m - m_new) + ((a0 + a1) + (a2 + a3));
        m = m_new;
        return;
      }
    }
    float y[32];
    float cmax = -INFINITY;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float s = __uint_as_float(v[j]);
      const bool ok = (j < valid) && (j != diag);
      float sc = p.inv_t;
      if (p.col_scale != nullptr) sc = (j < valid) ? __ldg(p.col_scale + col0 + j) : 0.f;
      y[j] = ok ? s * (p.col_scale != nullptr ? sc * kLog2e : p.scale_log2) : -INFINITY;
      cmax = fmaxf(cmax, y[j]);
      if (j == posj && row < p.rows_a && p.pos) p.pos[row] = s * sc;
    }
    if (cmax == -INFINITY) return;
    const float m_new = fmaxf(m, cmax);
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) acc += exp2f(y[j] - m_new);
    l = l * exp2f(m - m_new) + acc;
    m = m_new;
  }
  __device__ __forceinline__ void end_unit(const Params& p, int row, int, int slot) {
    if (row < p.rows_a) p.part[static_cast<size_t>(row) * p.num_slots + slot] = make_float2(m, l);
  }
};

// ------------------------------------------------------------------------- forward, symmetric schedule
// S = F F^T is symmetric: with whole 256-row tiles only the tiles on and above the diagonal are computed
// (GemmShape::skip_below_diagonal, 528 of 1024 at 2N = 8192) and a tile above the diagonal serves BOTH its rows (sums
// along the rows, as in LseEpilogue) and, transposed, the rows of its mirror image (sums down the columns).  A column
// sum adds up values held by different threads, so all terms need ONE reference instead of per-row running maxima:
//   M = scale * max_i |f_i|^2 * (1 + 2^-10) >= every y_ij (Cauchy-Schwarz),   term = 2^(y_ij - M) <= 1.
// Nothing overflows; nothing is flushed to zero as long as y_ij - M >= -2M >= -kBoundedSpan, which the launch checks
// on the device against the row bound (normalised rows: M = 28.9 at the reference's T = 0.05).  When the check fails
// the same launch walks the full matrix instead (GemmShape::active_mode 1) and this functor hands every call to an
// LseEpilogue (running maxima).
//
// Control words in the workspace (zeroed by the pack kernel): the loss-reduction counter and the row bound.
enum : int { kCtlLossCounter = 0, kCtlBoundBits = 1 };

// max_i |f_i|^2 over the bf16 operand rows (what the tensor core multiplies), a warp per row, one fire-and-forget
// atomic max per block (non-negative floats order like their bit patterns).  The launches that follow derive M and
// the choice of schedule from the word themselves (gemm_tc.cuh: bounded_reference, launch_gate_open).  dim % 8 == 0.
__global__ void __launch_bounds__(256)
row_bound_kernel(const __nv_bfloat16* __restrict__ f, int rows, int dim, unsigned int* __restrict__ ctl) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  float acc0 = 0.f, acc1 = 0.f;
  if (row < rows) {
    const uint4* src = reinterpret_cast<const uint4*>(f + static_cast<size_t>(row) * dim);
    const int n16 = dim / 8;
#pragma unroll 4
    for (int i = lane; i < n16; i += 32) {
      const uint4 w = __ldg(src + i);
      const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float lo = __uint_as_float(ws[t] << 16), hi = __uint_as_float(ws[t] & 0xffff0000u);
        acc0 = fmaf(lo, lo, acc0);
        acc1 = fmaf(hi, hi, acc1);
      }
    }
  }
  float acc = acc0 + acc1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ float warp_max[8];
  if (lane == 0) warp_max[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float mx = 0.f;
    for (int w = 0; w < 8; ++w) mx = fmaxf(mx, warp_max[w]);   // (a NaN row drops out here; its NaN logits still reach the loss)
    atomicMax(ctl + kCtlBoundBits, __float_as_uint(mx));
  }
}

struct SymLseEpilogue {
  struct Params {
    float* rowpart;            // [rows][row_slots]  sum_j 2^(y_ij - M) over the columns of slot (tile t, column group)
    float* colpart;            // [tiles * 8][col_pitch]  entry (m * 8 + w, j): sum_i 2^(y_ij - M) over the 32 rows of warp w of A tile m
    int col_pitch;             // rows + 32 (slots of a column not a power of two apart; measured: no different from a pitch of `rows`)
    float* pos;                // [rows] positive logit in natural units
    const unsigned int* ctl;   // control words (M)
    int rows;                  // 2N, a multiple of the 256-row tile
    int row_slots;
    int half;
    float scale_log2;
    float inv_t;
    LseEpilogue::Params full;  // the stand-in when the logits are not bounded: (max, sum) partials of the full matrix
  };
  static constexpr bool kUsesScratch = false;
  static constexpr bool kStagesColumns = false;
  float l, mref;
  int tile_m, tile_t;
  bool bounded;
  LseEpilogue full;

  __device__ __forceinline__ void begin_unit(const Params& p, int row, int m, int t) {
    const unsigned int bits = __ldg(p.ctl + kCtlBoundBits);
    bounded = logits_bounded(p.scale_log2, bits);
    if (!bounded) {
      full.begin_unit(p.full, row, m, t);
      return;
    }
    l = 0.f;
    tile_m = m;
    tile_t = t;
    mref = bounded_reference(p.scale_log2, bits);
  }
  __device__ __forceinline__ void chunk(const Params& p, int row, int col0, const uint32_t (&v)[32]) {
    if (!bounded) {
      full.chunk(p.full, row, col0, v);
      return;
    }
    const int lane = threadIdx.x & 31;
    float e[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) e[j] = fast_ex2(fmaf(__uint_as_float(v[j]), p.scale_log2, -mref));
    const int pc = row < p.half ? row + p.half : row - p.half;
    const int posj = pc - col0, diag = row - col0;
    if (static_cast<unsigned>(posj) < 32u) {   // the positive of this row: also that of row pc, whose own walk never sees column `row`
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j == posj) {
          const float s = __uint_as_float(v[j]) * p.inv_t;
          p.pos[row] = s;
          if (tile_t != tile_m) p.pos[pc] = s;
        }
    }
    if (tile_t == tile_m && static_cast<unsigned>(diag) < 32u) {   // the diagonal is not a logit (contrastive_loss.py:65-68)
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j == diag) e[j] = 0.f;
    }
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      a0 += e[j + 0];
      a1 += e[j + 1];
      a2 += e[j + 2];
      a3 += e[j + 3];
    }
    l += (a0 + a1) + (a2 + a3);
    if (tile_t == tile_m) return;   // a diagonal tile holds both triangles of its block: its rows have seen every column
    // Column sums over the warp's 32 rows: a transposing butterfly.  At distance d a lane keeps the half of its values
    // whose column bit matches its own lane bit, hands the other half to its partner and adds what it receives: 16 + 8 +
    // 4 + 2 + 1 shuffles, after which lane j holds the sum of column col0 + j.
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      const bool upper = (lane & d) != 0;
#pragma unroll
      for (int j = 0; j < d; ++j) {
        const float send = upper ? e[j] : e[j + d];
        const float keep = upper ? e[j + d] : e[j];
        e[j] = keep + __shfl_xor_sync(0xffffffffu, send, d);
      }
    }
    const int warp_in_tile = ((row - lane) - tile_m * 256) >> 5;   // 0..7 over the CTA pair's 256 rows
    p.colpart[static_cast<size_t>(tile_m * 8 + warp_in_tile) * p.col_pitch + col0 + lane] = e[0];
  }
  __device__ __forceinline__ void end_unit(const Params& p, int row, int m, int slot) {
    if (!bounded) {
      full.end_unit(p.full, row, m, slot);
      return;
    }
    p.rowpart[static_cast<size_t>(row) * p.row_slots + slot] = l;
  }
};

// ------------------------------------------------------------------------- backward: dL/dlogits
//   mode 0 (in-batch, symmetrised so that dF = H F, contrastive_loss.py:62 has F on both sides):
//          H[i][j] = c (2^(y_ij - L_i) + 2^(y_ij - L_j) - 2 [j == pos(i)]),  H[i][i] = 0
//   mode 1 (queue of NCELoss, the .repeat(2,1) at :80):  W[i][c] = c (2^(y - L_i) + 2^(y - L_{i+N}))
//   mode 2 (queue of the MoCo InfoNCE, :32):             W[i][c] = c  2^(y - L_i)
//   mode 3 (ProtoNCE, :115-124):                         W[i][c] = c  s_c (2^(y - L_i) - [c == i])
//   c = coef * grad[0];  y = S * scale_log2, or S * col_scale[c] * log2(e) when col_scale is given.
template <typename OutT>
struct GradLogitEpilogue {
  struct Params {
    OutT* out;               // [rows_a][ld_out]
    const float* lse2;       // row LSE in log2 domain
    const float* grad;       // device scalar: upstream dL
    const float* col_scale;  // [rows_b] or nullptr
    long long ld_out;
    int rows_a;
    int rows_b;
    int mode;
    int half;                // mode 0: N
    int n_rows_q;            // mode 1: N (second LSE is lse2[row + N])
    float scale_log2;
    float coef;
    int debug = 0;           // tuning instrumentation (debug.flags): 8 skip the stores, 16 never share one 2^y between the two softmax terms
    int mirror_rows = 0;     // mode 0, H symmetric: rows per A tile (256) when the tiles below the diagonal block are
                             // not computed (GemmShape::skip_below_diagonal): a chunk above it is also written transposed
  };
  static constexpr bool kUsesScratch = sizeof(OutT) == 2;
  static constexpr bool kTmaOutput = sizeof(OutT) == 2;      // `out` can be described by a tensor map (GemmShape::epi_tma_store)
  static constexpr bool kStagesColumns = sizeof(OutT) == 2;   // mode 0 needs the LSE of every COLUMN: staged per tile by the kernel
  uint8_t* scratch = nullptr;  // per-warp smem staging (tensor-core kernel only), see store32_coalesced
  const CUtensorMap* out_map = nullptr;  // when set: chunks leave through TMA stores (store32_tma)
  const float* cols = nullptr; // the 32 staged column values of the current chunk in shared memory (tensor-core kernel only)
  float li, li2, c;
  // One MUFU per score instead of two.  Both softmax terms of a score share 2^y:
  //   mode 0:  c (2^(y - L_i) + 2^(y - L_j)) = 2^(y - L_i) (c + u_i v_j),  u_i = c 2^(L_i - ref),  v_j = 2^(ref - L_j)
  //   mode 1:  c (2^(y - L_i) + 2^(y - L_i')) = 2^(y - L_i) w_i,           w_i = c (1 + 2^(L_i - L_i'))
  // (with two exponentials per score the MUFU pipe alone needs 2 x 32 768 ex2 per 128 x 256 tile at 16 per clock = 4096
  //  clocks against the 3072 the tile's MMAs take; measured: 221 -> 214 us per NCELoss step at 8192 x 768.)  ref = the LSE of row 0 -- it cancels, it only centres the two factors.  The factors are
  // used only while |L - ref| <= kFactorRange: u_i v_j then stays below 2^60, and a first term flushed to zero
  // (y - L_i < -126) hides at most 2^-66 of the second one.  A warp whose rows, or a chunk whose columns, lie further
  // out (temperatures far below the reference's 0.05 on unnormalised rows) takes the two-MUFU form.
  static constexpr float kFactorRange = 30.f;
  float ref = 0.f, ui = 0.f;
  bool rows_centred = false;

  // staged per tile column (mode 0): v_j, or -1 when column j is out of the factor range
  __device__ __forceinline__ float column_value(const Params& p, int col) const {
    if (p.mode != 0 || col >= p.rows_b) return 0.f;
    const float d = ref - __ldg(p.lse2 + col);
    return fabsf(d) <= kFactorRange ? fast_ex2(d) : -1.f;
  }

  __device__ __forceinline__ void begin_unit(const Params& p, int row, int, int) {
    const int r = min(row, p.rows_a - 1);
    li = p.lse2[r];
    li2 = (p.mode == 1) ? p.lse2[r + p.n_rows_q] : 0.f;
    c = p.coef * __ldg(p.grad);
    if (p.mode == 0) {
      if constexpr (kStagesColumns) {   // tensor-core kernel: whole warps call begin_unit, and the columns are staged
        ref = __ldg(p.lse2);
        const float d = li - ref;
        ui = c * fast_ex2(fminf(fmaxf(d, -kFactorRange), kFactorRange));
        rows_centred = __all_sync(0xffffffffu, fabsf(d) <= kFactorRange) && !(p.debug & 16);
      }
    } else if (p.mode == 1) {
      const float d = li - li2;
      rows_centred = fabsf(d) <= 2.f * kFactorRange && !(p.debug & 16);
      ui = c + c * fast_ex2(fminf(fmaxf(d, -2.f * kFactorRange), 2.f * kFactorRange));   // w_i
    }
  }
  static __device__ __forceinline__ void store32(OutT* dst, const float (&h)[32]) {
    if constexpr (sizeof(OutT) == 2) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 pk;
        __nv_bfloat162 t0 = __floats2bfloat162_rn(h[j + 0], h[j + 1]);
        __nv_bfloat162 t1 = __floats2bfloat162_rn(h[j + 2], h[j + 3]);
        __nv_bfloat162 t2 = __floats2bfloat162_rn(h[j + 4], h[j + 5]);
        __nv_bfloat162 t3 = __floats2bfloat162_rn(h[j + 6], h[j + 7]);
        pk.x = *reinterpret_cast<uint32_t*>(&t0);
        pk.y = *reinterpret_cast<uint32_t*>(&t1);
        pk.z = *reinterpret_cast<uint32_t*>(&t2);
        pk.w = *reinterpret_cast<uint32_t*>(&t3);
        *reinterpret_cast<uint4*>(dst + j) = pk;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(h[j], h[j + 1], h[j + 2], h[j + 3]);
    }
  }

  // The same staging, handed to the TMA: the warp writes its 32 x 32 bf16 block into shared memory in the 64-byte
  // swizzle pattern (16-byte chunk j of row r at chunk j ^ ((r >> 1) & 3): conflict-free), one lane issues a bulk
  // tensor store (rows past the end of the matrix are clipped by the tensor map) and the warp moves on; the wait for
  // the TMA to have READ the block happens just before the next chunk is staged, a chunk's worth of arithmetic later.
  // Replaces 4 LDS + 4 STG + their address arithmetic per chunk and takes the store latency off the warp.
  __device__ __forceinline__ void store32_tma(int row, int col0, const float (&h)[32]) const {
    const int lane = threadIdx.x & 31;
    const uint32_t base = smem_u32(scratch);
    if (lane == 0) tma_store_wait_read();
    __syncwarp();
    const uint32_t mine = base + lane * 64;
    const uint32_t sw = (static_cast<uint32_t>(lane) >> 1) & 3u;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      __nv_bfloat162 t0 = __floats2bfloat162_rn(h[j + 0], h[j + 1]);
      __nv_bfloat162 t1 = __floats2bfloat162_rn(h[j + 2], h[j + 3]);
      __nv_bfloat162 t2 = __floats2bfloat162_rn(h[j + 4], h[j + 5]);
      __nv_bfloat162 t3 = __floats2bfloat162_rn(h[j + 6], h[j + 7]);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(mine + (((j >> 3) ^ sw) << 4)),
                   "r"(*reinterpret_cast<uint32_t*>(&t0)), "r"(*reinterpret_cast<uint32_t*>(&t1)),
                   "r"(*reinterpret_cast<uint32_t*>(&t2)), "r"(*reinterpret_cast<uint32_t*>(&t3))
                   : "memory");
    }
    fence_proxy_async_smem();   // generic-proxy writes -> visible to the async proxy (TMA)
    __syncwarp();
    if (lane == 0) tma_store_2d(out_map, scratch, col0, row - lane);
  }
  __device__ __forceinline__ void finish() const {
    if constexpr (kUsesScratch) {
      if (out_map != nullptr && (threadIdx.x & 31) == 0) tma_store_wait_all();
    }
  }

  // A thread owns a row: stored directly, one instruction writes 16 bytes into each of 32 rows (32 half-filled
  // sectors; measured 43 us of a 350 us step at 8192^2).  Staged through shared memory instead, four lanes
  // cover a row's 64 bytes and one instruction writes 8 rows x 2 whole sectors.  bf16 only; all 32 lanes call.
  __device__ __forceinline__ void store32_coalesced(const Params& p, int row, int col0, const float (&h)[32]) const {
    const int lane = threadIdx.x & 31;
    constexpr int kPitch = 80;  // 64 data + 16 pad: conflict-free 16-byte writes, one 2-way conflict on the reads
    // explicit shared-space accesses: through the generic `scratch` pointer these were LD.E / ST.E (long scoreboard)
    const uint32_t base = smem_u32(scratch);
    const uint32_t mine = base + lane * kPitch;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      uint4 pk;
      __nv_bfloat162 t0 = __floats2bfloat162_rn(h[j + 0], h[j + 1]);
      __nv_bfloat162 t1 = __floats2bfloat162_rn(h[j + 2], h[j + 3]);
      __nv_bfloat162 t2 = __floats2bfloat162_rn(h[j + 4], h[j + 5]);
      __nv_bfloat162 t3 = __floats2bfloat162_rn(h[j + 6], h[j + 7]);
      pk.x = *reinterpret_cast<uint32_t*>(&t0);
      pk.y = *reinterpret_cast<uint32_t*>(&t1);
      pk.z = *reinterpret_cast<uint32_t*>(&t2);
      pk.w = *reinterpret_cast<uint32_t*>(&t3);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(mine + 2 * j), "r"(pk.x), "r"(pk.y), "r"(pk.z), "r"(pk.w) : "memory");
    }
    __syncwarp();
    const int row0 = row - lane, piece = lane & 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = (lane >> 2) + 8 * i;
      uint4 val;
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w) : "r"(base + r * kPitch + 16 * piece) : "memory");
      // evict-last: the next kernel (dF = H F) reads this matrix back; it should still be in L2 then
      if (row0 + r < p.rows_a)
        asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p.out + static_cast<long long>(row0 + r) * p.ld_out + col0 + 8 * piece),
                     "r"(val.x), "r"(val.y), "r"(val.z), "r"(val.w), "l"(kEvictLast)
                     : "memory");
    }
    __syncwarp();  // the staging area is reused by the next chunk (store32_mirrored, if any, only reads it)
  }

  // The transposed copy of the chunk store32_coalesced has just staged (H is symmetric): element (r, j) of the
  // 32 x 32 block goes to out[col0 + j][row0 + r].  Four lanes cover an output row's 64 bytes, as above.
  __device__ __forceinline__ void store32_mirrored(const Params& p, int row, int col0) const {
    const int lane = threadIdx.x & 31;
    constexpr int kPitch = 80;
    const int row0 = row - lane, piece = lane & 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int j = (lane >> 2) + 8 * i;  // column of the block = output row
      uint32_t w[4];
#pragma unroll
      for (int h2 = 0; h2 < 4; ++h2) {
        const int r = 8 * piece + 2 * h2;
        const uint32_t lo = *reinterpret_cast<const unsigned short*>(scratch + r * kPitch + 2 * j);
        const uint32_t hi = *reinterpret_cast<const unsigned short*>(scratch + (r + 1) * kPitch + 2 * j);
        w[h2] = lo | (hi << 16);
      }
      // (the host enables mirroring only when the rows are a multiple of the tile: every chunk is full)
      *reinterpret_cast<uint4*>(p.out + static_cast<long long>(col0 + j) * p.ld_out + row0 + 8 * piece) =
            make_uint4(w[0], w[1], w[2], w[3]);
    }
    __syncwarp();
  }

  __device__ __forceinline__ void chunk(const Params& p, int row, int col0, const uint32_t (&v)[32]) {
    const int valid = min(32, p.rows_b - col0);
    if (valid <= 0) return;
    const bool live = row < p.rows_a;
    const int diag = row - col0;
    int posj = -1;
    if (p.mode == 0) {
      const int pc = row < p.half ? row + p.half : row - p.half;
      posj = pc - col0;
    }
    OutT* dst_fast = p.out + static_cast<long long>(row) * p.ld_out + col0;
    if (p.mode <= 2 && p.col_scale == nullptr && valid == 32 && ((p.ld_out * sizeof(OutT)) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(p.out) & 15) == 0) {
      // Fast path: a full, aligned chunk with one temperature -- per score one FFMA + one MUFU per softmax term; a
      // chunk that holds the row's diagonal element or its positive (mode 0) patches that one element afterwards.
      // (Those chunks used to take the general path below: 2 of a row's 8 chunks in a diagonal tile -- every tile at
      // the reference's own batch of 128, where the kernel took 28 us for one 256 x 256 tile.)
      {
        bool staged = false;
        if constexpr (kUsesScratch) staged = scratch != nullptr;
        if (!live && !staged) return;  // (a dead lane still takes part in the staged store of its warp)
        float h[32];
        if (p.mode == 0) {
          // the factors v_j of the chunk's columns come from the kernel's per-tile shared-memory stage (broadcast
          // reads); without a stage (FFMA kernel), or out of the factor range, the column LSEs come from global memory
          bool factored = false;
          if (cols != nullptr)    // (staged: the whole warp is here)
            factored = rows_centred && !__any_sync(0xffffffffu, cols[threadIdx.x & 31] < 0.f);
          if (factored) {
            const float4* vp = reinterpret_cast<const float4*>(cols);
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 vj = vp[j4];
              const float vjs[4] = {vj.x, vj.y, vj.z, vj.w};
#pragma unroll
              for (int t = 0; t < 4; ++t)
                h[4 * j4 + t] = fast_ex2(fmaf(__uint_as_float(v[4 * j4 + t]), p.scale_log2, -li)) * fmaf(ui, vjs[t], c);
            }
          } else {
            const float4* lp = reinterpret_cast<const float4*>(p.lse2 + col0);  // 16-byte aligned
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 lj = __ldg(lp + j4);
              const float ljs[4] = {lj.x, lj.y, lj.z, lj.w};
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float s = __uint_as_float(v[4 * j4 + t]);
                h[4 * j4 + t] = c * (fast_ex2(fmaf(s, p.scale_log2, -li)) + fast_ex2(fmaf(s, p.scale_log2, -ljs[t])));
              }
            }
          }
          if (static_cast<unsigned>(diag) < 32u || static_cast<unsigned>(posj) < 32u) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (j == posj) h[j] -= 2.f * c;   // P_ij + P_ji - 2 at the positive
              if (j == diag) h[j] = 0.f;        // the diagonal is not a logit (contrastive_loss.py:64-67)
            }
          }
        } else if (p.mode == 1) {
          if (rows_centred) {   // per thread: the two LSEs of this row are close enough to share one 2^y
#pragma unroll
            for (int j = 0; j < 32; ++j) h[j] = fast_ex2(fmaf(__uint_as_float(v[j]), p.scale_log2, -li)) * ui;
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float s = __uint_as_float(v[j]);
              h[j] = c * (fast_ex2(fmaf(s, p.scale_log2, -li)) + fast_ex2(fmaf(s, p.scale_log2, -li2)));
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) h[j] = c * fast_ex2(fmaf(__uint_as_float(v[j]), p.scale_log2, -li));
        }
        if (p.debug & 8) return;
        if constexpr (kUsesScratch) {
          if (staged && out_map != nullptr && p.mirror_rows == 0) {
            store32_tma(row, col0, h);
            return;
          }
          if (staged) {
            store32_coalesced(p, row, col0, h);
            // (row - lane) / mirror_rows is the A tile of the whole warp: the condition is warp-uniform
            if (p.mirror_rows > 0 && col0 >= ((row - (threadIdx.x & 31)) / p.mirror_rows + 1) * p.mirror_rows)
              store32_mirrored(p, row, col0);
            return;
          }
        }
        store32(dst_fast, h);
        return;
      }
    }
    if (!live) return;
    float h[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float sc = 1.f;
      float y;
      if (p.col_scale != nullptr) {
        sc = (j < valid) ? __ldg(p.col_scale + col0 + j) : 0.f;
        y = __uint_as_float(v[j]) * sc * kLog2e;
      } else {
        y = __uint_as_float(v[j]) * p.scale_log2;
      }
      float val;
      if (p.mode == 0) {
        const float lj = (j < valid) ? __ldg(p.lse2 + col0 + j) : 0.f;
        val = exp2f(y - li) + exp2f(y - lj) - (j == posj ? 2.f : 0.f);
        if (j == diag) val = 0.f;
      } else if (p.mode == 1) {
        val = exp2f(y - li) + exp2f(y - li2);
      } else if (p.mode == 2) {
        val = exp2f(y - li);
      } else {
        val = sc * (exp2f(y - li) - (j == diag ? 1.f : 0.f));
      }
      h[j] = (j < valid) ? c * val : 0.f;
    }
    OutT* dst = p.out + static_cast<long long>(row) * p.ld_out + col0;
    if constexpr (sizeof(OutT) == 2) {
      if (valid == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint4 pk;
          __nv_bfloat162 t0 = __floats2bfloat162_rn(h[j + 0], h[j + 1]);
          __nv_bfloat162 t1 = __floats2bfloat162_rn(h[j + 2], h[j + 3]);
          __nv_bfloat162 t2 = __floats2bfloat162_rn(h[j + 4], h[j + 5]);
          __nv_bfloat162 t3 = __floats2bfloat162_rn(h[j + 6], h[j + 7]);
          pk.x = *reinterpret_cast<uint32_t*>(&t0);
          pk.y = *reinterpret_cast<uint32_t*>(&t1);
          pk.z = *reinterpret_cast<uint32_t*>(&t2);
          pk.w = *reinterpret_cast<uint32_t*>(&t3);
          *reinterpret_cast<uint4*>(dst + j) = pk;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < valid) dst[j] = __float2bfloat16_rn(h[j]);
      }
    } else {
      if (valid == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(dst + j) = make_float4(h[j], h[j + 1], h[j + 2], h[j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < valid) dst[j] = h[j];
      }
    }
    // the rare chunk that holds a positive, above the diagonal block: its mirror image is not computed either
    if (p.mirror_rows > 0 && col0 >= (row / p.mirror_rows + 1) * p.mirror_rows) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < valid) p.out[static_cast<long long>(col0 + j) * p.ld_out + row] = static_cast<OutT>(h[j]);
    }
  }
  __device__ __forceinline__ void end_unit(const Params&, int, int, int) {}
};

// ------------------------------------------------------------------------- plain fp32 store (dF = H F)
// Rows [0, split_row) go to out0, rows [split_row, rows_a) to out1 (dq / dk of cat(q, k)).
struct StoreEpilogue {
  struct Params {
    float* out0;
    float* out1;
    long long ld_out;
    int rows_a;
    int rows_b;
    int split_row;
    int accumulate;  // 1: out += value
    // split-K launches (GemmShape::k_splits > 1): K slice ks of the product goes, un-accumulated, to
    // partials + ks * rows_a * rows_b as a compact [rows_a][rows_b] matrix; splitk_reduce_kernel sums the slices in a
    // fixed order into out0 / out1 afterwards (deterministic, unlike atomic adds)
    float* partials = nullptr;
    int num_splits = 1;   // GemmShape::num_splits of the launch (begin_unit is handed split + num_splits * ks)
    int debug = 0;        // tuning instrumentation (debug.flags): 32 store every row directly (no staging)
  };
  static constexpr bool kUsesScratch = true;    // tensor-core kernel: a staging strip per warp (store_staged)
  static constexpr bool kStagesColumns = false;
  uint8_t* scratch = nullptr;
  const CUtensorMap* out_map = nullptr;   // (unused: set by the kernel for every functor with a scratch strip)
  float* slice = nullptr;
  __device__ __forceinline__ void finish() const {}
  __device__ __forceinline__ void begin_unit(const Params& p, int, int, int split) {
    if (p.partials != nullptr)
      slice = p.partials + static_cast<long long>(split / p.num_splits) * p.rows_a * p.rows_b;
  }
  __device__ __forceinline__ float* row_ptr(const Params& p, int row) const {
    return slice != nullptr ? slice + static_cast<long long>(row) * p.rows_b
                            : (row < p.split_row ? p.out0 + static_cast<long long>(row) * p.ld_out
                                                 : p.out1 + static_cast<long long>(row - p.split_row) * p.ld_out);
  }
  // A thread owns a row, so a direct store instruction writes 16 bytes into each of 32 rows (32 half-filled sectors).
  // Staged through the warp's shared-memory strip in two halves of 16 columns (32 rows x 64 bytes, pitch 80: the
  // geometry of GradLogitEpilogue::store32_coalesced), four lanes cover a row's 64 bytes and one instruction writes 8
  // rows x 2 whole sectors.  The whole warp calls; rows past the end are skipped on the way out.
  __device__ __forceinline__ void store_staged(const Params& p, int row, int col0, const uint32_t (&v)[32], bool acc) const {
    const int lane = threadIdx.x & 31;
    constexpr int kPitch = 80;
    const uint32_t base = smem_u32(scratch);
    const uint32_t mine = base + lane * kPitch;
    const int row0 = row - lane, piece = lane & 3;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(mine + 4 * j), "r"(v[16 * half + j]), "r"(v[16 * half + j + 1]),
                     "r"(v[16 * half + j + 2]), "r"(v[16 * half + j + 3]) : "memory");
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = (lane >> 2) + 8 * i;
        float4 val;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(val.x), "=f"(val.y), "=f"(val.z), "=f"(val.w) : "r"(base + r * kPitch + 16 * piece) : "memory");
        if (row0 + r < p.rows_a) {
          float4* dst = reinterpret_cast<float4*>(row_ptr(p, row0 + r) + col0 + 16 * half + 4 * piece);
          if (acc) {
            const float4 old = *dst;
            val.x += old.x; val.y += old.y; val.z += old.z; val.w += old.w;
          }
          *dst = val;
        }
      }
      __syncwarp();   // the strip is reused by the next half / chunk
    }
  }
  __device__ __forceinline__ void chunk(const Params& p, int row, int col0, const uint32_t (&v)[32]) {
    const int valid = min(32, p.rows_b - col0);
    if (valid <= 0) return;
    const bool acc = p.accumulate && slice == nullptr;
    if (scratch != nullptr && valid == 32 && !(p.debug & 32)) {   // (uniform over the warp)
      const long long pitch = slice != nullptr ? p.rows_b : p.ld_out;
      const uintptr_t bases = slice != nullptr ? reinterpret_cast<uintptr_t>(slice)
                                               : (reinterpret_cast<uintptr_t>(p.out0) | reinterpret_cast<uintptr_t>(p.out1));
      if ((pitch & 3) == 0 && (bases & 15) == 0) {
        store_staged(p, row, col0, v, acc);
        return;
      }
    }
    if (row >= p.rows_a) return;
    float* dst = slice != nullptr ? slice + static_cast<long long>(row) * p.rows_b + col0
                 : (row < p.split_row ? p.out0 + static_cast<long long>(row) * p.ld_out
                                      : p.out1 + static_cast<long long>(row - p.split_row) * p.ld_out) + col0;
    if (valid == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                               __uint_as_float(v[j + 3]));
        if (acc) {
          const float4 old = *reinterpret_cast<const float4*>(dst + j);
          o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
        }
        *reinterpret_cast<float4*>(dst + j) = o;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < valid) dst[j] = __uint_as_float(v[j]) + (acc ? dst[j] : 0.f);
    }
  }
  __device__ __forceinline__ void end_unit(const Params&, int, int, int) {}
};

// out (+)= sum over the K slices, slice 0 first (a fixed order: the result does not depend on which cluster ran which slice)
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ partials, int k_splits, StoreEpilogue::Params p) {
  const long long total = static_cast<long long>(p.rows_a) * p.rows_b;
  for (long long e = blockIdx.x * 256ll + threadIdx.x; e < total; e += gridDim.x * 256ll) {
    float acc = 0.f;
    for (int ks = 0; ks < k_splits; ++ks) acc += partials[ks * total + e];
    const int row = static_cast<int>(e / p.rows_b), col = static_cast<int>(e - static_cast<long long>(row) * p.rows_b);
    float* dst = (row < p.split_row ? p.out0 + static_cast<long long>(row) * p.ld_out
                                    : p.out1 + static_cast<long long>(row - p.split_row) * p.ld_out) + col;
    *dst = p.accumulate ? *dst + acc : acc;
  }
}

// ------------------------------------------------------------------------- helper kernels
// Operand staging through 32 x 32 shared-memory tiles, so that the row-major copy AND the transposed
// copy are both written with full 128-byte lines (a plain "one thread per element" transpose writes
// 2-byte elements 16 KB apart and took 60 us for the 8192 x 768 operand; this takes ~10).
//   src: [R][C] fp32 row-major; rows [0, split) come from s0, rows [split, R) from s1 (cat(q, k), :61)
//   any of: o32 [R][ld_o] fp32, obf [R][ld_o] bf16, o32_t [C][ld_t] fp32, obf_t [C][ld_t] bf16
//   (ld_o >= C, ld_t >= R: row pitches in elements; the pad columns are never written nor read -- the tensor maps
//    that read these buffers stop at the true extent and zero-fill beyond it)
__global__ void __launch_bounds__(256)
pack_transpose_kernel(const float* __restrict__ s0, const float* __restrict__ s1, long long split, long long R, long long C,
                      float* __restrict__ o32, __nv_bfloat16* __restrict__ obf, float* __restrict__ o32_t,
                      __nv_bfloat16* __restrict__ obf_t, long long ld_o, long long ld_t) {
  __shared__ float tile[32][33];
  const long long tiles_c = (C + 31) / 32, tiles_r = (R + 31) / 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (long long t = blockIdx.x; t < tiles_c * tiles_r; t += gridDim.x) {
    const long long r0 = (t / tiles_c) * 32, c0 = (t % tiles_c) * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long r = r0 + ty + 8 * i, c = c0 + tx;
      float x = 0.f;
      if (r < R && c < C) {
        x = r < split ? s0[r * C + c] : s1[(r - split) * C + c];
        if (o32) o32[r * ld_o + c] = x;
        if (obf) obf[r * ld_o + c] = __float2bfloat16_rn(x);
      }
      tile[ty + 8 * i][tx] = x;
    }
    __syncthreads();
    if (o32_t || obf_t) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long c = c0 + ty + 8 * i, r = r0 + tx;  // consecutive threads -> consecutive r
        if (r < R && c < C) {
          const float x = tile[tx][ty + 8 * i];
          if (o32_t) o32_t[c * ld_t + r] = x;
          if (obf_t) obf_t[c * ld_t + r] = __float2bfloat16_rn(x);
        }
      }
    }
    __syncthreads();
  }
}

// The InfoNCE operands: F = cat(q, k) [R = 2N][C] fp32 -> bf16 F (row-major) and bf16 F^T [C][R], 64 x 64 tiles, every
// global access a full 128-byte line (float4 loads, 8-byte / 16-byte bf16 stores).  C % 4 == 0, R % 8 == 0 (the bf16
// path requires dim % 8 and n % 4).  Also zeroes the control words the fused loss reduction and row_bound_kernel count in.
__global__ void __launch_bounds__(256)
pack_cat_bf16_kernel(const float* __restrict__ s0, const float* __restrict__ s1, int split, int R, int C,
                     __nv_bfloat16* __restrict__ obf, __nv_bfloat16* __restrict__ obf_t, unsigned int* __restrict__ zero_word) {
  __shared__ float tile[64][65];
  if (blockIdx.x == 0 && threadIdx.x < 2 && zero_word != nullptr) zero_word[threadIdx.x] = 0u;   // loss counter, row bound (kCtl...)
  const int tiles_c = (C + 63) / 64, tiles_r = (R + 63) / 64;
  const int tid = threadIdx.x;
  for (int t = blockIdx.x; t < tiles_c * tiles_r; t += gridDim.x) {
    const int r0 = (t / tiles_c) * 64, c0 = (t % tiles_c) * 64;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r0 + (tid >> 4) + 16 * i, c = c0 + (tid & 15) * 4;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < R && c < C) {
        x = *reinterpret_cast<const float4*>(r < split ? s0 + static_cast<size_t>(r) * C + c : s1 + static_cast<size_t>(r - split) * C + c);
        __nv_bfloat162 lo = __floats2bfloat162_rn(x.x, x.y), hi = __floats2bfloat162_rn(x.z, x.w);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&lo);
        pk.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(obf + static_cast<size_t>(r) * C + c) = pk;
      }
      float* dst = &tile[(tid >> 4) + 16 * i][(tid & 15) * 4];
      dst[0] = x.x; dst[1] = x.y; dst[2] = x.z; dst[3] = x.w;
    }
    __syncthreads();
    if (obf_t != nullptr) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int cl = (tid >> 3) + 32 * i, rl = (tid & 7) * 8;   // output row = column of the tile, 8 consecutive r per thread
        const int c = c0 + cl, r = r0 + rl;
        if (c < C && r < R) {
          uint32_t w[4];
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            __nv_bfloat162 v = __floats2bfloat162_rn(tile[rl + 2 * h][cl], tile[rl + 2 * h + 1][cl]);
            w[h] = *reinterpret_cast<uint32_t*>(&v);
          }
          *reinterpret_cast<uint4*>(obf_t + static_cast<size_t>(c) * R + r) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
    __syncthreads();
  }
}

// (m, l) pairs in the log2 domain: l_a 2^m_a + l_b 2^m_b
__device__ __forceinline__ void lse_combine(float& m, float& l, float m2, float l2) {
  const float mn = fmaxf(m, m2);
  if (mn == -INFINITY) return;  // both empty
  l = l * exp2f(m - mn) + l2 * exp2f(m2 - mn);
  m = mn;
}

// Fused loss reduction of the row-LSE kernels: every block leaves the sum of its rows' terms in `partials`, the block
// that finishes last adds them up in index order (deterministic) and resets `counter`.  256 threads, all call.
__device__ __forceinline__ void fused_loss_tail(float block_sum, float loss_scale, float* __restrict__ loss,
                                                float* __restrict__ partials, unsigned int* __restrict__ counter) {
  __shared__ bool last;
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = block_sum;
    __threadfence();
    last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  __shared__ float red[256];
  float acc = 0.f;                                     // fixed assignment of partials to threads, fixed tree
  for (unsigned int i = threadIdx.x; i < gridDim.x; i += 256) acc += __ldcg(partials + i);
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    loss[0] = loss_scale * red[0];
    *counter = 0u;
  }
}

// Row log-sum-exp from the slot partials: one WARP per row, lanes stride over the slots (coalesced;
// the old one-block kernel walked 74 partials per thread 592 bytes apart and took 540 us at 2N = 8192).
//   part   [rows][slots]          in-batch / prototype partials
//   part_q [rows_q][slots_q]      optional queue partials of row (i mod rows_q)  (the .repeat(2,1), :80)
//   include_pos: the positive logit pos[i] is an extra column (MoCo form, :30-36)
//   loss (optional): loss[0] = loss_scale * sum_i (lse_i - pos_i), fused in: every block leaves the sum of its 8 rows in
//   `partials`, the block that finishes last adds them up in index order (deterministic) and resets `counter` (which
//   the caller zeroed once, before the first call on this workspace -- the pack kernel does)
__global__ void __launch_bounds__(256)
lse_rows_kernel(const float2* __restrict__ part, int slots, const float2* __restrict__ part_q, int slots_q, int rows_q,
                const float* __restrict__ pos, int include_pos, int rows, float* __restrict__ lse,
                float* __restrict__ lse2, float loss_scale = 0.f, float* __restrict__ loss = nullptr,
                float* __restrict__ partials = nullptr, unsigned int* __restrict__ counter = nullptr) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  float term = 0.f;
  if (row < rows) {
    float m = -INFINITY, l = 0.f;
    if (include_pos && lane == 0) {
      m = pos[row] * kLog2e;
      l = 1.f;
    }
    for (int s = lane; s < slots; s += 32) {
      const float2 p = part[static_cast<size_t>(row) * slots + s];
      lse_combine(m, l, p.x, p.y);
    }
    if (part_q != nullptr) {
      const int r = row % rows_q;
      for (int s = lane; s < slots_q; s += 32) {
        const float2 p = part_q[static_cast<size_t>(r) * slots_q + s];
        lse_combine(m, l, p.x, p.y);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
      const float l2 = __shfl_xor_sync(0xffffffffu, l, o);
      lse_combine(m, l, m2, l2);
    }
    const float v = m + log2f(l);
    if (lane == 0) {
      lse2[row] = v;
      lse[row] = v * kLn2;
    }
    if (loss != nullptr) term = v * kLn2 - pos[row];
  }
  if (loss == nullptr) return;
  __shared__ float warp_term[8];
  if (lane == 0) warp_term[threadIdx.x >> 5] = term;
  __syncthreads();
  float sum = 0.f;
  if (threadIdx.x == 0)
    for (int w = 0; w < 8; ++w) sum += warp_term[w];
  fused_loss_tail(sum, loss_scale, loss, partials, counter);
}

// The same for the symmetric forward (SymLseEpilogue): plain sums against the common reference M.  A block takes 32
// consecutive rows (one 256-row tile holds them all): the column sums of the tiles above the rows' own tile are read
// slot by slot as whole 128-byte lines (lane = row), the rows' own slots row by row; then the queue partials, the LSE
// and the loss term as above.  Falls back to the (max, sum) partials of the full-matrix launch when that one ran.
__global__ void __launch_bounds__(256)
lse_rows_sym_kernel(const float* __restrict__ rowpart, const float* __restrict__ colpart, int col_pitch, int row_slots,
                    const unsigned int* __restrict__ ctl, float scale_log2, const float2* __restrict__ part, int slots,
                    const float2* __restrict__ part_q, int slots_q, int rows_q, const float* __restrict__ pos, int rows,
                    float* __restrict__ lse, float* __restrict__ lse2, float loss_scale, float* __restrict__ loss,
                    float* __restrict__ partials, unsigned int* __restrict__ counter) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * 32;
  const unsigned int bits = __ldg(ctl + kCtlBoundBits);
  const float mref = bounded_reference(scale_log2, bits);
  const bool sym = 2.f * mref <= kBoundedSpan;
  __shared__ float colsum[8][33];
  __shared__ float warp_term[8];
  const int tile = row0 >> 8;
  float own[4] = {0.f, 0.f, 0.f, 0.f};   // this lane's share of the four rows' own slots
  if (sym) {
    // Every load of a thread is issued before the first one is used (one L2 round trip instead of one per slot: the
    // kernel is a latency chain, not a bandwidth problem -- 10 MB of partials); the sums run in a fixed order.
    const float* cp = colpart + row0 + lane;
    float x[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int s = warp + 8 * i;
      x[i] = s < 8 * tile ? __ldg(cp + static_cast<size_t>(s) * col_pitch) : 0.f;
    }
    float beyond = 0.f;   // (more than 256 column slots: 2N > 8192)
    for (int s = 256 + warp; s < 8 * tile; s += 8) beyond += __ldg(cp + static_cast<size_t>(s) * col_pitch);
    float y[4][2];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float* rp = rowpart + static_cast<size_t>(row0 + warp * 4 + r) * row_slots;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int t = 2 * tile + lane + 32 * i;
        y[r][i] = t < row_slots ? rp[t] : 0.f;
      }
    }
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int i = 0; i < 32; i += 4) { a0 += x[i]; a1 += x[i + 1]; a2 += x[i + 2]; a3 += x[i + 3]; }
    colsum[warp][lane] = ((a0 + a1) + (a2 + a3)) + beyond;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      own[r] = y[r][0] + y[r][1];
      for (int t = 2 * tile + lane + 64; t < row_slots; t += 32)   // (more than 64 slots: 2N > 8192)
        own[r] += rowpart[static_cast<size_t>(row0 + warp * 4 + r) * row_slots + t];
    }
  }
  __syncthreads();
  float term = 0.f;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int ro = warp * 4 + r, row = row0 + ro;
    float m = -INFINITY, l = 0.f, total = 0.f;
    if (sym) {
      float sum = own[r];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
#pragma unroll
      for (int w = 0; w < 8; ++w) sum += colsum[w][ro];   // (broadcast reads: every lane ends up with the row's total)
      total = sum;
      if (lane == 0 && sum > 0.f) { m = mref; l = sum; }
    } else {
      for (int s = lane; s < slots; s += 32) {
        const float2 p = part[static_cast<size_t>(row) * slots + s];
        lse_combine(m, l, p.x, p.y);
      }
    }
    float v;
    if (sym && part_q == nullptr) {
      v = mref + log2f(total);   // no queue: nothing to combine with
    } else {
      if (part_q != nullptr) {
        const int rq = row % rows_q;
        for (int s = lane; s < slots_q; s += 32) {
          const float2 p = part_q[static_cast<size_t>(rq) * slots_q + s];
          lse_combine(m, l, p.x, p.y);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
        const float l2 = __shfl_xor_sync(0xffffffffu, l, o);
        lse_combine(m, l, m2, l2);
      }
      v = m + log2f(l);
    }
    if (lane == 0) {
      lse2[row] = v;
      lse[row] = v * kLn2;
    }
    term += v * kLn2 - pos[row];
  }
  if (lane == 0) warp_term[warp] = term;
  __syncthreads();
  float sum = 0.f;
  if (threadIdx.x == 0)
    for (int w = 0; w < 8; ++w) sum += warp_term[w];
  fused_loss_tail(sum, loss_scale, loss, partials, counter);
}

// loss = scale * sum_i (lse_i - pos_i)   (contrastive_loss.py:92 sum/2; :24,:42 mean).  One block, fixed
// reduction tree: deterministic.
__global__ void __launch_bounds__(1024)
loss_sum_kernel(const float* __restrict__ lse, const float* __restrict__ pos, int rows, float scale,
                float* __restrict__ loss) {
  __shared__ float red[32];
  float local = 0.f;
  for (int i = threadIdx.x; i < rows; i += blockDim.x) local += lse[i] - pos[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) loss[0] = scale * v;
  }
}

// MoCo positives (contrastive_loss.py:30): pos[i] = (q_i . k_i) * inv_T, fp32.  One warp per row.
__global__ void rowdot_kernel(const float* __restrict__ q, const float* __restrict__ k, int n, int dim, float inv_t,
                              float* __restrict__ pos) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  float acc = 0.f;
  for (int d = threadIdx.x & 31; d < dim; d += 32)
    acc = fmaf(q[static_cast<long long>(row) * dim + d], k[static_cast<long long>(row) * dim + d], acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) pos[row] = acc * inv_t;
}

// MoCo backward, positive column (contrastive_loss.py:30,:42): a_i = g (p0_i - 1) inv_T / N with
// p0_i = exp(pos_i - lse_i);  dq_i = a_i k_i,  dk_i = a_i q_i.  One warp per row.
__global__ void moco_pos_grad_kernel(const float* __restrict__ q, const float* __restrict__ k, int n, int dim,
                                     float inv_t, const float* __restrict__ lse, const float* __restrict__ grad,
                                     float* __restrict__ dq, float* __restrict__ dk) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const float* qr = q + static_cast<long long>(row) * dim;
  const float* kr = k + static_cast<long long>(row) * dim;
  float acc = 0.f;
  for (int d = threadIdx.x & 31; d < dim; d += 32) acc = fmaf(qr[d], kr[d], acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  const float a = __ldg(grad) * (expf(acc * inv_t - lse[row]) - 1.f) * inv_t / static_cast<float>(n);
  for (int d = threadIdx.x & 31; d < dim; d += 32) {
    dq[static_cast<long long>(row) * dim + d] = a * kr[d];
    dk[static_cast<long long>(row) * dim + d] = a * qr[d];
  }
}

__global__ void scale_to_log2_kernel(const float* __restrict__ lse, int n, float* __restrict__ lse2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) lse2[i] = lse[i] * kLog2e;
}

}  // namespace drs
