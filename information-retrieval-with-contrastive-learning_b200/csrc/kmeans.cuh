// Prototype clustering around the flat-L2 search (src/contrastor/utils.py:50-105, `run_kmeans`): the centroid update of
// a Lloyd iteration and the per-cluster distance statistics of the concentration estimate (:73-83).
//
// The assignment (nearest centroid of every sample) is drs_search_l2; its output is sorted by cluster on the host side
// (a stable sort: members stay in ascending sample order), so a cluster's members are a contiguous run of `order`.
// One block per cluster walks its run: thread t owns coordinates t, t + 128, ... and adds the members' values in run
// order in fp64 -- the result does not depend on the grid or on timing (atomics would), and it equals a sequential
// float64 sum over the samples in index order, which is what the oracle computes.  HBM-bound: every sample row is read
// once (coalesced 4 * dim bytes), 4 rows in flight per thread.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace drs {

constexpr int kClusterThreads = 128;

//   x        [n][dim] fp32 samples
//   order    [n]      sample indices sorted by assigned cluster (stable)
//   offsets  [k + 1]  run boundaries in `order`
//   dist     [n]      squared distance of every sample to its centroid, or nullptr
//   centroids [k][dim] in: previous centroids; out: mean of the members (an EMPTY cluster keeps its previous centroid)
//   sum_sqrt [k]      out (when dist != nullptr): sum of sqrt(dist) over the members, fp64 accumulation in run order
__global__ void __launch_bounds__(kClusterThreads)
cluster_update_kernel(const float* __restrict__ x, int dim, const long long* __restrict__ order,
                      const long long* __restrict__ offsets, long long k, const float* __restrict__ dist,
                      float* __restrict__ centroids, float* __restrict__ sum_sqrt) {
  __shared__ double red[kClusterThreads];
  for (long long c = blockIdx.x; c < k; c += gridDim.x) {
    const long long m0 = offsets[c], m1 = offsets[c + 1];
    const long long cnt = m1 - m0;
    if (centroids != nullptr && cnt > 0) {
      for (int d = threadIdx.x; d < dim; d += kClusterThreads) {
        double acc = 0.0;
        long long m = m0;
        for (; m + 4 <= m1; m += 4) {       // four independent loads in flight, added in run order
          const float v0 = x[order[m] * dim + d], v1 = x[order[m + 1] * dim + d];
          const float v2 = x[order[m + 2] * dim + d], v3 = x[order[m + 3] * dim + d];
          acc += v0; acc += v1; acc += v2; acc += v3;
        }
        for (; m < m1; ++m) acc += x[order[m] * dim + d];
        centroids[c * dim + d] = static_cast<float>(acc / static_cast<double>(cnt));
      }
    }
    if (dist != nullptr && sum_sqrt != nullptr) {
      double acc = 0.0;
      for (long long m = m0 + threadIdx.x; m < m1; m += kClusterThreads) acc += static_cast<double>(sqrtf(fmaxf(dist[order[m]], 0.f)));
      red[threadIdx.x] = acc;
      __syncthreads();
      for (int s = kClusterThreads / 2; s > 0; s >>= 1) {   // fixed tree: the same sum on every run
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
      }
      if (threadIdx.x == 0) sum_sqrt[c] = static_cast<float>(red[0]);
      __syncthreads();
    }
  }
}

}  // namespace drs
