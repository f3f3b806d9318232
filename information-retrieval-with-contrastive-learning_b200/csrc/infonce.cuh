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
  };
  float li, li2, c;

  __device__ __forceinline__ void begin_unit(const Params& p, int row, int, int) {
    const int r = min(row, p.rows_a - 1);
    li = p.lse2[r];
    li2 = (p.mode == 1) ? p.lse2[r + p.n_rows_q] : 0.f;
    c = p.coef * __ldg(p.grad);
  }
  __device__ __forceinline__ void chunk(const Params& p, int row, int col0, const uint32_t (&v)[32]) {
    if (row >= p.rows_a) return;
    const int valid = min(32, p.rows_b - col0);
    if (valid <= 0) return;
    const int diag = row - col0;
    int posj = -1;
    if (p.mode == 0) {
      const int pc = row < p.half ? row + p.half : row - p.half;
      posj = pc - col0;
    }
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

// rows [R][dim] fp32 -> bf16 copy and/or bf16 transpose [dim][R]  (ProtoNCE prototypes, MoCo q)
__global__ void pack_rows_kernel(const float* __restrict__ src, long long rows, int dim,
                                 __nv_bfloat16* __restrict__ bf, __nv_bfloat16* __restrict__ bf_t) {
  const long long total = rows * dim;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / dim;
    const int d = static_cast<int>(i - r * dim);
    const __nv_bfloat16 x = __float2bfloat16_rn(src[i]);
    if (bf) bf[i] = x;
    if (bf_t) bf_t[static_cast<long long>(d) * rows + r] = x;
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

// Generic finalize: lse_i from the slot partials (plus the positive logit itself when it is not
// among the GEMM columns, MoCo form), loss = loss_scale * sum_i (lse_i - pos_i).  One block.
__global__ void __launch_bounds__(1024)
lse_finalize_kernel(const float2* __restrict__ part, int slots, const float* __restrict__ pos, int include_pos,
                    int rows, float loss_scale, float* __restrict__ lse, float* __restrict__ lse2,
                    float* __restrict__ loss) {
  __shared__ float red[32];
  float local = 0.f;
  for (int i = threadIdx.x; i < rows; i += blockDim.x) {
    float m = -INFINITY, l = 0.f;
    if (include_pos) {
      m = pos[i] * kLog2e;
      l = 1.f;
    }
    for (int s = 0; s < slots; ++s) {
      const float2 p = part[static_cast<size_t>(i) * slots + s];
      if (p.x == -INFINITY) continue;
      const float mn = fmaxf(m, p.x);
      l = l * exp2f(m - mn) + p.y * exp2f(p.x - mn);
      m = mn;
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
    if (threadIdx.x == 0) loss[0] = loss_scale * v;
  }
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
