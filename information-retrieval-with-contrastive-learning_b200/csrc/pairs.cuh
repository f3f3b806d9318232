// Per-document sentence-pair similarity: the body of get_docs_sents_similarity
// (preprocessing/build_docs_sentence_similarity.py:48-66) for a whole batch of documents.
//
// Input is what `vectorizer.transform(doc)` (:49) yields, stacked over documents: one CSR matrix of
// sentence rows (float64 TF-IDF weights, sorted column indices) plus the sentence range of every
// document.  The reference calls sklearn's cosine_similarity (:50) on each document and then walks
// the strict upper triangle in Python (:59-63) and sorts by score, descending, stably (:65).
//
// Everything is float64 and reproduces the reference BIT FOR BIT, because the order of every
// floating-point operation is the reference's own:
//   * normalize(X): per row, sum of squares accumulated left to right, sqrt, each value divided by
//     the norm (sklearn.utils.sparsefuncs_fast.inplace_csr_row_normalize_l2);
//   * X_n @ X_n.T (scipy csr_matmat): entry (i, j) = products x_n[i,c] * x_n[j,c] added in ascending
//     column order c, starting from 0.0; no fused multiply-add (explicit _rn intrinsics below);
//   * list.sort(key=score, reverse=True) is stable: equal scores keep (i, j) lexicographic order.
// This is irregular integer/index work (sorted-list intersections), bound by HBM/L2 latency, not a
// GEMM: one block per document, one thread per pair, rank sort with the keys in shared memory.
#pragma once
#include <stdint.h>

namespace drs {

// x_n = x / ||row||   (rows with zero norm are left as they are)
__global__ void csr_row_normalize_kernel(const long long* __restrict__ indptr, const double* __restrict__ data,
                                         long long num_rows, double* __restrict__ out) {
  const long long r = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (r >= num_rows) return;
  const long long p0 = indptr[r], p1 = indptr[r + 1];
  double sum = 0.0;
  for (long long p = p0; p < p1; ++p) sum = __dadd_rn(sum, __dmul_rn(data[p], data[p]));
  if (sum == 0.0) {
    for (long long p = p0; p < p1; ++p) out[p] = data[p];
    return;
  }
  const double nrm = __dsqrt_rn(sum);
  for (long long p = p0; p < p1; ++p) out[p] = __ddiv_rn(data[p], nrm);
}

// double -> uint64 that orders like the double (-0.0 folded into +0.0: Python compares them equal)
__device__ __forceinline__ unsigned long long f64_ordered(double v) {
  unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(v));
  if (b == 0x8000000000000000ull) b = 0ull;
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

// linear index p of the strict upper triangle (row-major: (0,1),(0,2),...,(0,n-1),(1,2),...) -> (i, j)
__device__ __forceinline__ void decode_pair(long long p, int n, int* pi, int* pj) {
  const double b = 2.0 * n - 1.0;
  int i = static_cast<int>((b - sqrt(b * b - 8.0 * static_cast<double>(p))) * 0.5);
  if (i < 0) i = 0;
  if (i > n - 2) i = n - 2;
  auto start = [n](int r) { return static_cast<long long>(r) * (2LL * n - r - 1) / 2; };
  while (i > 0 && start(i) > p) --i;
  while (i < n - 2 && start(i + 1) <= p) ++i;
  *pi = i;
  *pj = static_cast<int>(p - start(i)) + i + 1;
}

__device__ __forceinline__ double csr_rows_dot(const int* __restrict__ indices, const double* __restrict__ xn,
                                               long long a0, long long a1, long long b0, long long b1) {
  double acc = 0.0;
  while (a0 < a1 && b0 < b1) {
    const int ca = indices[a0], cb = indices[b0];
    if (ca == cb) {
      acc = __dadd_rn(acc, __dmul_rn(xn[a0], xn[b0]));
      ++a0;
      ++b0;
    } else if (ca < cb) {
      ++a0;
    } else {
      ++b0;
    }
  }
  return acc;
}

// One block per document.  pair_offsets[d] = first output slot of document d (a document with n
// sentences owns n(n-1)/2 slots, 1 slot when n == 1 -- the ((0,0), s00) entry of :54-57 -- 0 when empty).
// tmp: [total_pairs] float64 scratch (unsorted scores).  Dynamic shared memory: smem_keys * 8 bytes.
__global__ void __launch_bounds__(256)
doc_sentence_pairs_kernel(const long long* __restrict__ indptr, const int* __restrict__ indices,
                          const double* __restrict__ xn, const long long* __restrict__ doc_offsets, long long num_docs,
                          const long long* __restrict__ pair_offsets, int smem_keys, double* __restrict__ tmp,
                          int* __restrict__ out_i, int* __restrict__ out_j, double* __restrict__ out_score) {
  extern __shared__ unsigned long long pk_smem[];
  for (long long d = blockIdx.x; d < num_docs; d += gridDim.x) {
    const long long s0 = doc_offsets[d];
    const int n = static_cast<int>(doc_offsets[d + 1] - s0);
    if (n <= 0) continue;
    const long long o0 = pair_offsets[d];
    const long long np = n == 1 ? 1 : static_cast<long long>(n) * (n - 1) / 2;
    const bool in_smem = np <= smem_keys;
    // 1. scores, in the reference's pair order
    for (long long p = threadIdx.x; p < np; p += blockDim.x) {
      int i = 0, j = 0;
      if (n > 1) decode_pair(p, n, &i, &j);
      const double s = csr_rows_dot(indices, xn, indptr[s0 + i], indptr[s0 + i + 1], indptr[s0 + j], indptr[s0 + j + 1]);
      tmp[o0 + p] = s;
      if (in_smem) pk_smem[p] = f64_ordered(s);
    }
    __syncthreads();
    // 2. stable descending rank sort: slot = #pairs that come before this one
    for (long long p = threadIdx.x; p < np; p += blockDim.x) {
      const double s = tmp[o0 + p];
      const unsigned long long key = f64_ordered(s);
      long long rank = 0;
      if (in_smem) {
        for (long long t = 0; t < np; ++t) {
          const unsigned long long kt = pk_smem[t];
          rank += (kt > key) || (kt == key && t < p);
        }
      } else {
        for (long long t = 0; t < np; ++t) {
          const unsigned long long kt = f64_ordered(tmp[o0 + t]);
          rank += (kt > key) || (kt == key && t < p);
        }
      }
      int i = 0, j = 0;
      if (n > 1) decode_pair(p, n, &i, &j);
      out_i[o0 + rank] = i;
      out_j[o0 + rank] = j;
      out_score[o0 + rank] = s;
    }
    __syncthreads();
  }
}

}  // namespace drs
