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

#include "topk.cuh"

namespace drs {

static constexpr float kLog2e = 1.4426950408889634f;
static constexpr float kLn2 = 0.6931471805599453f;

// ------------------------------------------------------------------------- forward: row LSE partials
// Per (row, split): running max m and l = sum 2^(y - m) over the split's columns (log2 domain).
struct LseEpilogue {
  struct Params {
    float2* part;     // [rows_a][num_slots]  (m, l)
    float* pos;       // [rows_a] positive logit (natural units, S * inv_T), or nullptr
    int rows_a;
    int rows_b;
    int num_slots;
    int half;         // N: pos(i) = (i + N) mod 2N; 0 = no positive / no diagonal mask (queue operand)
    float scale_log2; // inv_T * log2(e)
    float inv_t;
  };
  float m, l;

  __device__ __forceinline__ void begin_unit(const Params&, int, int, int) {
    m = -INFINITY;
    l = 0.f;
  }
  __device__ __forceinline__ void chunk(const Params& p, int row, int col0, const uint32_t (&v)[32]) {
    const int valid = p.rows_b - col0;
    const int diag = p.half ? row - col0 : -1;                       // column offset to skip
    int posj = -1;
    if (p.half) {
      const int pc = row < p.half ? row + p.half : row - p.half;
      posj = pc - col0;
    }
    float y[32];
    float cmax = -INFINITY;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float s = __uint_as_float(v[j]);
      const bool ok = (j < valid) && (j != diag);
      y[j] = ok ? s * p.scale_log2 : -INFINITY;
      cmax = fmaxf(cmax, y[j]);
      if (j == posj && row < p.rows_a && p.pos) p.pos[row] = s * p.inv_t;
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

// ------------------------------------------------------------------------- backward: dL/dS, symmetrised
// H[i][j] = c * (2^(y_ij - L_i) + 2^(y_ij - L_j) - 2 [j == pos(i)]),  H[i][i] = 0,  c = g * inv_T / 2
// so that dF = H F  (contrastive_loss.py:62 uses F on both sides of the product).
// Queue operand (half == 0): W[i][c] = c * (2^(y - L_i) + 2^(y - L_{i+N}))  (the .repeat(2,1) at :80).
template <typename OutT>
struct GradLogitEpilogue {
  struct Params {
    OutT* out;          // [rows_a][ld_out]
    const float* lse2;  // [2N] row LSE in log2 domain
    const float* grad;  // device scalar: upstream dL
    long long ld_out;
    int rows_a;
    int rows_b;
    int half;           // N (in-batch) ; 0 = queue operand
    int n_rows_q;       // queue operand: N (second LSE is lse2[row + N])
    float scale_log2;
    float inv_t;
  };
  float li, li2, c;

  __device__ __forceinline__ void begin_unit(const Params& p, int row, int, int) {
    const int r = min(row, p.rows_a - 1);
    li = p.lse2[r];
    li2 = p.half ? 0.f : p.lse2[r + p.n_rows_q];
    c = 0.5f * p.inv_t * __ldg(p.grad);
  }
  __device__ __forceinline__ void chunk(const Params& p, int row, int col0, const uint32_t (&v)[32]) {
    if (row >= p.rows_a) return;
    const int valid = min(32, p.rows_b - col0);
    if (valid <= 0) return;
    const int diag = p.half ? row - col0 : -1;
    int posj = -1;
    if (p.half) {
      const int pc = row < p.half ? row + p.half : row - p.half;
      posj = pc - col0;
    }
    float h[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float y = __uint_as_float(v[j]) * p.scale_log2;
      float val;
      if (p.half) {
        const float lj = (j < valid) ? __ldg(p.lse2 + col0 + j) : 0.f;
        val = exp2f(y - li) + exp2f(y - lj) - (j == posj ? 2.f : 0.f);
        if (j == diag) val = 0.f;
      } else {
        val = exp2f(y - li) + exp2f(y - li2);
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
  };
  __device__ __forceinline__ void begin_unit(const Params&, int, int, int) {}
  __device__ __forceinline__ void chunk(const Params& p, int row, int col0, const uint32_t (&v)[32]) {
    if (row >= p.rows_a) return;
    const int valid = min(32, p.rows_b - col0);
    if (valid <= 0) return;
    float* dst = (row < p.split_row ? p.out0 + static_cast<long long>(row) * p.ld_out
                                    : p.out1 + static_cast<long long>(row - p.split_row) * p.ld_out) + col0;
    if (valid == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                               __uint_as_float(v[j + 3]));
        if (p.accumulate) {
          const float4 old = *reinterpret_cast<const float4*>(dst + j);
          o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
        }
        *reinterpret_cast<float4*>(dst + j) = o;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < valid) dst[j] = __uint_as_float(v[j]) + (p.accumulate ? dst[j] : 0.f);
    }
  }
  __device__ __forceinline__ void end_unit(const Params&, int, int, int) {}
};

// ------------------------------------------------------------------------- helper kernels
// F = cat(q, k) in fp32 and/or bf16, plus bf16 F^T ([dim][2N]) for the K-major B operand of dF = H F.
__global__ void infonce_pack_kernel(const float* __restrict__ q, const float* __restrict__ k, int n, int dim,
                                    float* __restrict__ f32, __nv_bfloat16* __restrict__ bf, __nv_bfloat16* __restrict__ bf_t) {
  const long long total = 2ll * n * dim;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / dim), d = static_cast<int>(i - static_cast<long long>(r) * dim);
    const float x = r < n ? q[static_cast<long long>(r) * dim + d] : k[static_cast<long long>(r - n) * dim + d];
    if (f32) f32[i] = x;
    if (bf) bf[i] = __float2bfloat16_rn(x);
    if (bf_t) bf_t[static_cast<long long>(d) * (2ll * n) + r] = __float2bfloat16_rn(x);
  }
}
// queue [dim][K] fp32 -> transposed [K][dim] (fp32 and/or bf16) and a bf16 copy in the original layout
__global__ void infonce_queue_pack_kernel(const float* __restrict__ queue, int dim, long long klen,
                                          float* __restrict__ qt32, __nv_bfloat16* __restrict__ qt_bf,
                                          __nv_bfloat16* __restrict__ q_bf) {
  const long long total = klen * dim;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long c = i / dim;
    const int d = static_cast<int>(i - c * dim);
    const float x = queue[static_cast<long long>(d) * klen + c];
    if (qt32) qt32[i] = x;
    if (qt_bf) qt_bf[i] = __float2bfloat16_rn(x);
    if (q_bf) q_bf[static_cast<long long>(d) * klen + c] = __float2bfloat16_rn(x);
  }
}

// Combine the split partials (and the queue partials of row i mod N) into lse (natural log),
// lse2 (log2 domain) and the loss = sum_i (lse_i - pos_i) / 2   (contrastive_loss.py:92).
// One block; deterministic tree reduction.
__global__ void __launch_bounds__(1024)
infonce_finalize_kernel(const float2* __restrict__ part, int splits, const float2* __restrict__ part_q, int splits_q,
                        const float* __restrict__ pos, int two_n, int n, float* __restrict__ lse,
                        float* __restrict__ lse2, float* __restrict__ loss) {
  __shared__ float red[32];
  float local = 0.f;
  for (int i = threadIdx.x; i < two_n; i += blockDim.x) {
    float m = -INFINITY, l = 0.f;
    for (int s = 0; s < splits; ++s) {
      const float2 p = part[static_cast<size_t>(i) * splits + s];
      if (p.x == -INFINITY) continue;
      const float mn = fmaxf(m, p.x);
      l = l * exp2f(m - mn) + p.y * exp2f(p.x - mn);
      m = mn;
    }
    if (part_q) {
      const int r = i < n ? i : i - n;
      for (int s = 0; s < splits_q; ++s) {
        const float2 p = part_q[static_cast<size_t>(r) * splits_q + s];
        if (p.x == -INFINITY) continue;
        const float mn = fmaxf(m, p.x);
        l = l * exp2f(m - mn) + p.y * exp2f(p.x - mn);
        m = mn;
      }
    }
    const float l2 = m + log2f(l);
    lse2[i] = l2;
    lse[i] = l2 * kLn2;
    local += l2 * kLn2 - pos[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) loss[0] = 0.5f * v;
  }
}

__global__ void scale_to_log2_kernel(const float* __restrict__ lse, int n, float* __restrict__ lse2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) lse2[i] = lse[i] * kLog2e;
}

}  // namespace drs
