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

// Sorted (descending) list of the KCAP best keys seen so far, kept in registers: every loop is
// fully unrolled so there is no dynamic register indexing.  `thr` caches the score of the worst
// kept key: a candidate whose index is larger than every index seen so far (true for a scan in
// increasing index order) can only enter if score > thr, which is the one compare on the fast path.
template <int KCAP>
struct TopKList {
  uint64_t keys[KCAP];
  float thr;

  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int j = 0; j < KCAP; ++j) keys[j] = 0ull;
    thr = -INFINITY;
  }
  // bubble the key down from the top; a zero/smaller key passes through without change
  __device__ __forceinline__ void insert_key(uint64_t key) {
#pragma unroll
    for (int j = 0; j < KCAP; ++j) {
      const uint64_t cur = keys[j];
      const bool gt = key > cur;
      keys[j] = gt ? key : cur;
      key = gt ? cur : key;
    }
    thr = keys[KCAP - 1] == 0ull ? -INFINITY : key_score(keys[KCAP - 1]);
  }
  __device__ __forceinline__ void insert(float score, uint32_t idx) { insert_key(make_key(score, idx)); }
};

}  // namespace drs
