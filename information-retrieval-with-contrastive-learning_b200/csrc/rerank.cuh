// Candidate-restricted scoring: the report's pipeline "TF-IDF top-100 -> dense re-rank -> top-15"
// (report.pdf section 3.2; intended call site src/evaluation.py:105-116) and the paired claim /
// evidence score of the commented-out evaluation, `(clm_vec * evdn_vec).sum(dim=-1)`
// (src/evaluation.py:112,115).
//
// Both are HBM-bound gathers: every (claim, candidate) pair reads one corpus row exactly once
// (D * sizeof(T) bytes) and does D FMAs -- 1 flop/byte, far below the tensor-core ridge, so there
// is no GEMM to form.  One warp per pair streams the row with 16-byte loads (a full 128-byte line
// per 8 lanes), the claim sits in shared memory as fp32, and the claim's candidates are selected
// in the same block from packed (score, ~id) keys -- the score list never reaches memory.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "merge.cuh"

namespace drs {

template <typename T> struct RowVec;
// (load = raw + cvt; the gather keeps many raw 16-byte loads in flight and converts afterwards)
template <> struct RowVec<float> {
  static constexpr int W = 4;  // elements per 16-byte load
  static __device__ __forceinline__ uint4 raw(const float* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
  static __device__ __forceinline__ void cvt(const uint4& t, float (&v)[4]) {
    v[0] = __uint_as_float(t.x); v[1] = __uint_as_float(t.y); v[2] = __uint_as_float(t.z); v[3] = __uint_as_float(t.w);
  }
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
};
template <> struct RowVec<__nv_bfloat16> {
  static constexpr int W = 8;
  static __device__ __forceinline__ uint4 raw(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
  static __device__ __forceinline__ void cvt(const uint4& t, float (&v)[8]) {
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // bf16 -> fp32 is a 16-bit shift
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
};

template <> struct RowVec<__half> {
  static constexpr int W = 8;
  static __device__ __forceinline__ uint4 raw(const __half* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
  static __device__ __forceinline__ void cvt(const uint4& t, float (&v)[8]) {
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      v[2 * i] = f.x;
      v[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void load(const __half* p, float (&v)[8]) {
    const uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      v[2 * i] = f.x;
      v[2 * i + 1] = f.y;
    }
  }
};

// dot(row, q) with q in shared memory (fp32); the whole warp cooperates, result on every lane.
template <typename T, bool VEC>
__device__ __forceinline__ float warp_row_dot(const T* __restrict__ row, const float* __restrict__ qs, int dim, int lane) {
  float acc = 0.f;
  if constexpr (VEC) {
    constexpr int W = RowVec<T>::W;
    for (int d = lane * W; d < dim; d += 32 * W) {
      float v[W];
      RowVec<T>::load(row + d, v);
#pragma unroll
      for (int i = 0; i < W; ++i) acc = fmaf(v[i], qs[d + i], acc);
    }
  } else {
    for (int d = lane; d < dim; d += 32) acc = fmaf(static_cast<float>(row[d]), qs[d], acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  return acc;
}

// One block per claim.  cand: [nq, m] corpus row ids (id < 0 or >= nc: padding, ignored).
// out: [nq, k] scores descending / ids; ties -> lower id; a duplicated id is reported once;
// fewer than k valid candidates pad with (-inf, -1).
template <typename T, bool VEC>
__global__ void __launch_bounds__(256)
rerank_kernel(const T* __restrict__ queries, const T* __restrict__ corpus, const long long* __restrict__ cand, int nq,
              long long nc, int dim, int m, int k, float* __restrict__ out_scores, long long* __restrict__ out_ids) {
  extern __shared__ __align__(16) unsigned char rr_smem[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(rr_smem);                        // [m]
  float* qs = reinterpret_cast<float*>(rr_smem + static_cast<size_t>(m) * 8);   // [dim]
  const int q = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int d = threadIdx.x; d < dim; d += blockDim.x) qs[d] = static_cast<float>(queries[static_cast<size_t>(q) * dim + d]);
  __syncthreads();
  const long long* my = cand + static_cast<size_t>(q) * m;
  if constexpr (VEC) {
    // Four candidates per warp at a time (8 lanes each: 8 x 16 bytes = one 128-byte line per step) and up to twelve
    // independent 16-byte loads in flight per lane (a whole 768-wide bf16 row): one candidate per warp with three dependent steps left the
    // gather latency-bound at 3.1 TB/s (47 % of the copy bandwidth) for 10 000 claims x 100 candidates x 768.
    constexpr int W = RowVec<T>::W;
    constexpr int U = 12;
    const int sub = lane >> 3, l8 = lane & 7;
    const int nvec = dim / W;
    for (int base = warp * 4; base < m; base += nwarps * 4) {
      const int c = base + sub;
      const long long id = c < m ? __ldg(my + c) : -1ll;
      const bool ok = id >= 0 && id < nc;
      const T* row = corpus + static_cast<size_t>(ok ? id : 0) * dim;
      float acc = 0.f;
      for (int v0 = l8; v0 < nvec; v0 += 8 * U) {
        uint4 raw[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int v = v0 + 8 * u;
          raw[u] = (ok && v < nvec) ? RowVec<T>::raw(row + static_cast<size_t>(v) * W) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int v = min(v0 + 8 * u, nvec - 1);
          float x[W];
          RowVec<T>::cvt(raw[u], x);
#pragma unroll
          for (int i = 0; i < W; ++i) acc = fmaf(x[i], qs[v * W + i], acc);
        }
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      // a NaN score is never a candidate, exactly as in the search epilogue (`s > thr` is false for NaN) and as
      // numpy's argpartition / argsort order NaN last in closest_docs (tfidf_doc_ranker.py:70-71)
      if (l8 == 0 && c < m) keys[c] = (ok && acc == acc) ? make_key(acc, static_cast<uint32_t>(id)) : 0ull;
    }
  } else {
    for (int c = warp; c < m; c += nwarps) {
      const long long id = __ldg(my + c);
      uint64_t key = 0ull;
      if (id >= 0 && id < nc) {
        const float s = warp_row_dot<T, VEC>(corpus + static_cast<size_t>(id) * dim, qs, dim, lane);
        if (s == s) key = make_key(s, static_cast<uint32_t>(id));
      }
      if (lane == 0) keys[c] = key;
    }
  }
  __syncthreads();
  if (warp != 0) return;
  uint64_t prev = ~0ull;
  for (int r = 0; r < k; ++r) {
    uint64_t best = 0ull;
    for (int c = lane; c < m; c += 32) {
      const uint64_t key = keys[c];
      if (key < prev && key > best) best = key;
    }
    best = warp_max_u64(best);
    if (lane == 0) {
      out_scores[static_cast<size_t>(q) * k + r] = best ? key_score(best) : -INFINITY;
      out_ids[static_cast<size_t>(q) * k + r] = best ? static_cast<long long>(key_index(best)) : -1ll;
    }
    prev = best;
  }
}

// out[i] = a[i] . b[i]   -- `(clm_vec * evdn_vec).sum(dim=-1)`, src/evaluation.py:112.  One warp per row.
template <typename T, bool VEC>
__global__ void __launch_bounds__(256)
pair_scores_kernel(const T* __restrict__ a, const T* __restrict__ b, long long n, int dim, float* __restrict__ out) {
  const long long row = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const T* pa = a + row * dim;
  const T* pb = b + row * dim;
  float acc = 0.f;
  if constexpr (VEC) {
    constexpr int W = RowVec<T>::W;
    for (int d = lane * W; d < dim; d += 32 * W) {
      float va[W], vb[W];
      RowVec<T>::load(pa + d, va);
      RowVec<T>::load(pb + d, vb);
#pragma unroll
      for (int i = 0; i < W; ++i) acc = fmaf(va[i], vb[i], acc);
    }
  } else {
    for (int d = lane; d < dim; d += 32) acc = fmaf(static_cast<float>(pa[d]), static_cast<float>(pb[d]), acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[row] = acc;
}

}  // namespace drs
