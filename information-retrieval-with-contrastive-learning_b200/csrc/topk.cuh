// Ordered (score, index) keys and the per-thread running top-K list used by every
// select epilogue.  Order everywhere: score descending, then index ascending
// (BASELINE.json north star: "ties broken by lower index").
#pragma once
#include <stdint.h>

namespace drs {

// fp32 -> uint32 that sorts like the float (-0.0 folded into +0.0 first).
__device__ __forceinline__ uint32_t float_to_ordered(float v) {
  const uint32_t b = __float_as_uint(v + 0.0f);
  return b ^ (static_cast<uint32_t>(static_cast<int32_t>(b) >> 31) | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t o) {
  const uint32_t b = (o & 0x80000000u) ? (o ^ 0x80000000u) : ~o;
  return __uint_as_float(b);
}
// One 64-bit key per candidate: larger key == better candidate.  Key 0 == empty slot
// (it would need score bits 0xFFFFFFFF, a NaN, which is never inserted).
__device__ __forceinline__ uint64_t make_key(float score, uint32_t idx) {
  return (static_cast<uint64_t>(float_to_ordered(score)) << 32) | static_cast<uint64_t>(~idx);
}
__device__ __forceinline__ float key_score(uint64_t key) { return ordered_to_float(static_cast<uint32_t>(key >> 32)); }
__device__ __forceinline__ uint32_t key_index(uint64_t key) { return ~static_cast<uint32_t>(key); }

// Sorted (descending) list of the KCAP best (score, index) pairs seen so far by ONE thread, kept
// in registers: every loop is fully unrolled (no dynamic register indexing) and the insert is a
// set of INDEPENDENT compare/selects (position j takes the old j-1, the new value, or stays), so
// it issues at full rate instead of running a serial bubble chain.
// The scan feeds candidates in increasing index order, so a new candidate that ties an existing
// score has the larger index and belongs AFTER it: a strict `>` on the score alone implements
// the (score desc, index asc) order.  `thr` caches the k-th best score (k <= KCAP, runtime): the
// single compare on the fast path.  Empty slots are (-inf, 0xFFFFFFFF); NaN never enters.
template <int KCAP>
struct TopKList {
  float sc[KCAP];
  uint32_t ix[KCAP];
  float thr;    // a candidate must beat this: max(k-th best of the list, floor)
  float kth;    // the list's own k-th best score (-inf while it holds fewer than k)
  float floor;  // externally known lower bound: scores <= floor cannot be among the final k

  __device__ __forceinline__ void reset(float floor_ = -INFINITY) {
#pragma unroll
    for (int j = 0; j < KCAP; ++j) {
      sc[j] = -INFINITY;
      ix[j] = 0xFFFFFFFFu;
    }
    kth = -INFINITY;
    floor = floor_;
    thr = floor_;
  }
  // `s` = -inf makes this a no-op, which is how lanes without a candidate ride along
  __device__ __forceinline__ void insert(float s, uint32_t idx, int k) {
    bool gt[KCAP];
#pragma unroll
    for (int j = 0; j < KCAP; ++j) gt[j] = s > sc[j];
#pragma unroll
    for (int j = KCAP - 1; j >= 1; --j) {
      sc[j] = gt[j] ? (gt[j - 1] ? sc[j - 1] : s) : sc[j];
      ix[j] = gt[j] ? (gt[j - 1] ? ix[j - 1] : idx) : ix[j];
    }
    sc[0] = gt[0] ? s : sc[0];
    ix[0] = gt[0] ? idx : ix[0];
    float t = sc[0];
#pragma unroll
    for (int j = 1; j < KCAP; ++j) t = (j < k) ? sc[j] : t;
    kth = t;  // sc[k-1]
    thr = fmaxf(t, floor);
  }
  __device__ __forceinline__ uint64_t key(int j) const {
    return sc[j] == -INFINITY ? 0ull : make_key(sc[j], ix[j]);
  }
};

// v[j] for a warp-uniform runtime j without dynamic register indexing: a 5-level select tree.
__device__ __forceinline__ uint32_t pick32(const uint32_t (&v)[32], int j) {
  uint32_t a[16], b[8], c[4], d[2];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = (j & 1) ? v[2 * i + 1] : v[2 * i];
#pragma unroll
  for (int i = 0; i < 8; ++i) b[i] = (j & 2) ? a[2 * i + 1] : a[2 * i];
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i] = (j & 4) ? b[2 * i + 1] : b[2 * i];
#pragma unroll
  for (int i = 0; i < 2; ++i) d[i] = (j & 8) ? c[2 * i + 1] : c[2 * i];
  return (j & 16) ? d[1] : d[0];
}

}  // namespace drs
