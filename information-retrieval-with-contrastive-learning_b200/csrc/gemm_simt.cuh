// fp32 "rows x rows" GEMM on the FFMA pipe with the same per-row epilogue interface as gemm_tc.cuh.
// This is the exact-comparison path of the north star (scores within 1e-5 relative of the
// reference's fp32 torch.matmul, src/contrastor/contrastive_loss.py:62): plain fp32 FMAs, fp32
// accumulation, no tensor-core rounding of the inputs.
//
//   S[i, j] = sum_d A[i, d] * B[j, d]          (B_KN == false: B is [rows_b, K] row-major)
//   S[i, j] = sum_d A[i, d] * B[d, j]          (B_KN == true : B is [K, rows_b] row-major)
//
// 128 x 128 tile per CTA, K slices of 16, 8 x 8 outputs per thread, register-prefetched global
// loads.  After the K loop the tile is parked in shared memory and threads 0..127 each walk one
// row through the epilogue functor, exactly like the TMEM epilogue of the tensor-core kernel.
#pragma once
#include "gemm_tc.cuh"  // GemmShape

namespace drs {

struct SimtCfg {
  static constexpr int BM = 128, BN = 128, BK = 16;
  static constexpr int LDS = BM + 4;      // padded k-major operand rows
  static constexpr int LDT = BN + 1;      // padded score tile rows
  static constexpr int SMEM_BYTES = (2 * BK * LDS + BM * LDT) * 4;
  static constexpr int THREADS = 256;
};

template <class Epi, bool B_KN>
__global__ void __launch_bounds__(256)
gemm_simt_f32_kernel(const float* __restrict__ A, const float* __restrict__ B, long long lda, long long ldb, int K,
                     const GemmShape shp, const typename Epi::Params ep) {
  using C = SimtCfg;
  if (!launch_gate_open(shp)) return;
  extern __shared__ float smem_f[];
  float* As = smem_f;                    // [BK][LDS]
  float* Bs = smem_f + C::BK * C::LDS;   // [BK][LDS]
  float* Ts = Bs + C::BK * C::LDS;       // [BM][LDT]

  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int base_units = shp.num_m_tiles * shp.num_splits;
  const int num_units = base_units * max(1, shp.k_splits);   // split-K: kb_per_split counts BK = 16 wide slices here
  const bool vec_ok = ((lda & 3) == 0) && ((ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
  Epi epi;

  for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
    const int ks = shp.k_splits > 1 ? u / base_units : 0;
    const int v = u - ks * base_units;
    const int m = v % shp.num_m_tiles, s = v / shp.num_m_tiles;
    const int t0 = s * shp.tiles_per_split;
    const int t1 = min(t0 + shp.tiles_per_split, shp.total_b_tiles);
    const int row0 = m * C::BM;
    const int k_lo = shp.k_splits > 1 ? ks * shp.kb_per_split * C::BK : 0;
    const int k_hi = shp.k_splits > 1 ? min(K, k_lo + shp.kb_per_split * C::BK) : K;
    if (tid < C::BM) epi.begin_unit(ep, row0 + tid, m, s + shp.num_splits * ks);

    for (int t = t0; t < t1; ++t) {
      const int col0 = t * C::BN;
      float acc[8][8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

      // global -> register staging: 2 float4 per operand per thread per K slice
      float4 ra[2], rb[2];
      auto load_a = [&](int k0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int r = (tid >> 2) + h * 64, kq = (tid & 3) * 4;
          const int gr = row0 + r, gk = k0 + kq;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (gr < shp.rows_a) {
            const float* p = A + static_cast<long long>(gr) * lda + gk;
            if (vec_ok && gk + 3 < K) v = *reinterpret_cast<const float4*>(p);
            else {
              if (gk + 0 < K) v.x = p[0];
              if (gk + 1 < K) v.y = p[1];
              if (gk + 2 < K) v.z = p[2];
              if (gk + 3 < K) v.w = p[3];
            }
          }
          ra[h] = v;
        }
      };
      auto load_b = [&](int k0) {
        if constexpr (!B_KN) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int r = (tid >> 2) + h * 64, kq = (tid & 3) * 4;
            const int gr = col0 + r, gk = k0 + kq;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gr < shp.rows_b) {
              const float* p = B + static_cast<long long>(gr) * ldb + gk;
              if (vec_ok && gk + 3 < K) v = *reinterpret_cast<const float4*>(p);
              else {
                if (gk + 0 < K) v.x = p[0];
                if (gk + 1 < K) v.y = p[1];
                if (gk + 2 < K) v.z = p[2];
                if (gk + 3 < K) v.w = p[3];
              }
            }
            rb[h] = v;
          }
        } else {
          // B[k][n], n contiguous: thread loads 4 consecutive n of one k row
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int kk = (tid >> 5) + h * 8, nq = (tid & 31) * 4;
            const int gk = k0 + kk, gn = col0 + nq;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gk < K) {
              const float* p = B + static_cast<long long>(gk) * ldb + gn;
              if (vec_ok && gn + 3 < shp.rows_b) v = *reinterpret_cast<const float4*>(p);
              else {
                if (gn + 0 < shp.rows_b) v.x = p[0];
                if (gn + 1 < shp.rows_b) v.y = p[1];
                if (gn + 2 < shp.rows_b) v.z = p[2];
                if (gn + 3 < shp.rows_b) v.w = p[3];
              }
            }
            rb[h] = v;
          }
        }
      };
      auto stage_to_smem = [&]() {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int r = (tid >> 2) + h * 64, kq = (tid & 3) * 4;
          As[(kq + 0) * C::LDS + r] = ra[h].x;
          As[(kq + 1) * C::LDS + r] = ra[h].y;
          As[(kq + 2) * C::LDS + r] = ra[h].z;
          As[(kq + 3) * C::LDS + r] = ra[h].w;
          if constexpr (!B_KN) {
            Bs[(kq + 0) * C::LDS + r] = rb[h].x;
            Bs[(kq + 1) * C::LDS + r] = rb[h].y;
            Bs[(kq + 2) * C::LDS + r] = rb[h].z;
            Bs[(kq + 3) * C::LDS + r] = rb[h].w;
          } else {
            const int kk = (tid >> 5) + h * 8, nq = (tid & 31) * 4;
            *reinterpret_cast<float4*>(&Bs[kk * C::LDS + nq]) = rb[h];
          }
        }
      };

      load_a(k_lo);
      load_b(k_lo);
      for (int k0 = k_lo; k0 < k_hi; k0 += C::BK) {
        __syncthreads();  // previous slice fully consumed (and previous tile's scan finished)
        stage_to_smem();
        __syncthreads();
        if (k0 + C::BK < k_hi) {
          load_a(k0 + C::BK);
          load_b(k0 + C::BK);
        }
#pragma unroll
        for (int kk = 0; kk < C::BK; ++kk) {
          const float4 a0 = *reinterpret_cast<const float4*>(&As[kk * C::LDS + ty * 8]);
          const float4 a1 = *reinterpret_cast<const float4*>(&As[kk * C::LDS + ty * 8 + 4]);
          const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk * C::LDS + tx * 8]);
          const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk * C::LDS + tx * 8 + 4]);
          const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
          const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
      }
      // park the tile, then one thread per row walks it through the epilogue
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) Ts[(ty * 8 + i) * C::LDT + tx * 8 + j] = acc[i][j];
      __syncthreads();
      if (tid < C::BM) {
#pragma unroll 1
        for (int c = 0; c < C::BN; c += 32) {
          uint32_t v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(Ts[tid * C::LDT + c + j]);
          epi.chunk(ep, row0 + tid, col0 + c, v);
        }
      }
      // the next tile's first __syncthreads (top of its K loop) orders this scan before Ts is rewritten
    }
    if (tid < C::BM) epi.end_unit(ep, row0 + tid, m, s);  // one column group: slot == split
  }
}

}  // namespace drs
