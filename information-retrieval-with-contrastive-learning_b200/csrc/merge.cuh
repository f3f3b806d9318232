// Final selects: reduce the per-split candidate keys of one claim to its k best, and merge the
// per-shard (score, id) lists gathered from the other GPUs.  Order: score desc, id asc.
// One warp per claim.  Keys/pairs are unique per claim, so "the best candidate strictly worse
// than the previous pick" walks the order without mutating the candidate set.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "topk.cuh"

namespace drs {

__device__ __forceinline__ uint64_t warp_max_u64(uint64_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const uint64_t w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w > v ? w : v;
  }
  return v;
}

// ws: [nq][ncand] packed keys (0 = empty).  out_scores / out_ids: row pitch ld_out, k columns written
// (local row + id_base; -1 and -inf pad when fewer than k candidates exist).
// LCAP: candidates cached per lane in registers (ncand <= 32 * LCAP), else re-read from L2.
template <int LCAP>
__global__ void __launch_bounds__(128)
merge_keys_kernel(const uint64_t* __restrict__ ws, int nq, int ncand, int k, long long id_base,
                  float* __restrict__ out_scores, long long* __restrict__ out_ids, int ld_out,
                  uint64_t* __restrict__ bound_out, const float* __restrict__ row_term) {
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= nq) return;
  const uint64_t* src = ws + static_cast<size_t>(q) * ncand;
  uint64_t mine[LCAP > 0 ? LCAP : 1];
  if constexpr (LCAP > 0) {
#pragma unroll
    for (int i = 0; i < LCAP; ++i) {
      const int c = lane + 32 * i;
      mine[i] = c < ncand ? src[c] : 0ull;
    }
  }
  uint64_t prev = ~0ull;
  for (int r = 0; r < k; ++r) {
    uint64_t best = 0ull;
    if constexpr (LCAP > 0) {
#pragma unroll
      for (int i = 0; i < LCAP; ++i) {
        const uint64_t key = mine[i];
        if (key < prev && key > best) best = key;
      }
    } else {
      for (int c = lane; c < ncand; c += 32) {
        const uint64_t key = __ldg(src + c);
        if (key < prev && key > best) best = key;
      }
    }
    best = warp_max_u64(best);
    if (lane == 0) {
      // row_term: squared-L2 search reports |x|^2 - (2 x.c - |c|^2), ascending = ranked value descending
      float sc = best ? key_score(best) : -INFINITY;
      if (row_term != nullptr) sc = best ? fmaxf(row_term[q] - sc, 0.f) : INFINITY;
      out_scores[static_cast<size_t>(q) * ld_out + r] = sc;
      out_ids[static_cast<size_t>(q) * ld_out + r] = best ? static_cast<long long>(key_index(best)) + id_base : -1ll;
    }
    prev = best;  // best == 0 -> nothing is < 0: the remaining picks are all empty
  }
  // the next pass (k > list capacity) continues strictly below this pass's last pick
  if (bound_out != nullptr && lane == 0) bound_out[q] = prev;
}

// The candidates of one claim are `nslots` runs of `kcap` keys, each sorted descending (the register
// lists of the scan epilogue), so the k best come out of a k-way merge: each lane owns the heads of up
// to SL runs, one warp max per pick, the winning lane advances its run.  ~SL compares per pick instead
// of ncand / 32.  Returns pick r in lane r (0 = fewer than r + 1 candidates).  k <= 32.
template <int SL>
__device__ __forceinline__ uint64_t warp_select_runs(const uint64_t* __restrict__ src, int nslots, int kcap, int k, int lane) {
  uint64_t head[SL];
  int pos[SL];
#pragma unroll
  for (int i = 0; i < SL; ++i) {
    const int slot = lane + 32 * i;
    pos[i] = 0;
    head[i] = slot < nslots ? __ldg(src + static_cast<size_t>(slot) * kcap) : 0ull;
  }
  uint64_t mine = 0ull;
  for (int r = 0; r < k; ++r) {
    uint64_t m = 0ull;
#pragma unroll
    for (int i = 0; i < SL; ++i) m = head[i] > m ? head[i] : m;
    const uint64_t best = warp_max_u64(m);
    if (best == 0ull) break;
    if (lane == r) mine = best;
    if (m == best) {  // keys are unique per claim: exactly one lane, one run
#pragma unroll
      for (int i = 0; i < SL; ++i) {
        if (head[i] == best) {
          ++pos[i];
          head[i] = pos[i] < kcap ? __ldg(src + static_cast<size_t>(lane + 32 * i) * kcap + pos[i]) : 0ull;
        }
      }
    }
  }
  return mine;
}

// Final select of a single-pass search (k <= 32) by run merge: one warp per claim.
template <int SL>
__global__ void __launch_bounds__(128)
select_runs_kernel(const uint64_t* __restrict__ ws, int nq, int nslots, int kcap, int k, long long id_base,
                   float* __restrict__ out_scores, long long* __restrict__ out_ids, const float* __restrict__ row_term) {
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= nq) return;
  const uint64_t mine = warp_select_runs<SL>(ws + static_cast<size_t>(q) * nslots * kcap, nslots, kcap, k, lane);
  if (lane < k) {
    float sc = mine ? key_score(mine) : -INFINITY;
    if (row_term != nullptr) sc = mine ? fmaxf(row_term[q] - sc, 0.f) : INFINITY;
    out_scores[static_cast<size_t>(q) * k + lane] = sc;
    out_ids[static_cast<size_t>(q) * k + lane] = mine ? static_cast<long long>(key_index(mine)) + id_base : -1ll;
  }
}

// k > list capacity (32): ADAPTIVE passes.  Every (claim, slot) list holds the slot's KCAP best
// eligible keys, sorted descending, so the claim's candidates form `nslots` sorted runs and a
// k-way merge walks them best-first: each lane owns the heads of up to SL slots, one warp max
// per pick, the winning lane advances its run.  A full run (KCAP entries) may hide more of its
// slot below its last key; B = the largest such last key.  Every pick >= B is certainly the next
// best of the whole corpus (anything hidden is < its run's last key <= B), so picks are emitted
// while they stay >= B -- at least kcap per pass, usually all k on the first.  A claim that stops
// early leaves `bound` = its last pick and `done` = its count; the next pass rescans only keys
// below the bound and continues.  `remaining[0]` counts the unfinished claims of this pass: the
// next scan and merge return at once when it is zero.  Claims already complete are skipped.
template <int SL>
__global__ void __launch_bounds__(128)
merge_runs_kernel(const uint64_t* __restrict__ ws, int nq, int nslots, int kcap, int k, long long id_base,
                  float* __restrict__ out_scores, long long* __restrict__ out_ids, int ld_out,
                  uint64_t* __restrict__ bound, int* __restrict__ done, const unsigned int* __restrict__ active_in,
                  unsigned int* __restrict__ remaining_out, const float* __restrict__ row_term) {
  if (active_in != nullptr && *active_in == 0u) return;
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= nq) return;
  int r = done[q];
  if (r >= k) return;
  uint64_t keep = 0ull;
  const int r_begin = r;        // picks before this one were emitted by earlier passes
  const uint64_t* src = ws + static_cast<size_t>(q) * nslots * kcap;
  uint64_t head[SL], nxt[SL];   // nxt: the run's following key, loaded when the head is taken -- the pick loop
  int pos[SL];                  // never waits on memory (a dependent L2 load per pick made this kernel 1.1 ms)
  uint64_t hidden = 0ull;  // B
#pragma unroll
  for (int i = 0; i < SL; ++i) {
    const int slot = lane + 32 * i;
    pos[i] = 0;
    head[i] = 0ull;
    nxt[i] = 0ull;
    if (slot < nslots) {
      head[i] = src[static_cast<size_t>(slot) * kcap];
      nxt[i] = kcap > 1 ? src[static_cast<size_t>(slot) * kcap + 1] : 0ull;
      const uint64_t last = src[static_cast<size_t>(slot) * kcap + kcap - 1];
      hidden = last > hidden ? last : hidden;
    }
  }
  hidden = warp_max_u64(hidden);
  uint64_t prev = bound[q];
  bool exhausted = false;
  for (; r < k; ++r) {
    uint64_t mine = 0ull;
#pragma unroll
    for (int i = 0; i < SL; ++i) mine = head[i] > mine ? head[i] : mine;
    const uint64_t best = warp_max_u64(mine);
    if (best == 0ull) { exhausted = hidden == 0ull; break; }
    if (best < hidden) break;  // a hidden key could outrank it: the next pass continues below `prev`
    if (mine == best) {        // keys are unique: exactly one lane, one run
#pragma unroll
      for (int i = 0; i < SL; ++i) {
        if (head[i] == best) {
          ++pos[i];
          head[i] = nxt[i];
          nxt[i] = pos[i] + 1 < kcap ? src[static_cast<size_t>(lane + 32 * i) * kcap + pos[i] + 1] : 0ull;
        }
      }
    }
    if (lane == (r & 31)) {    // pick r waits in lane r % 32 for a coalesced store
      keep = best;
    }
    if ((r & 31) == 31 || r == k - 1) {
      const int r0 = r & ~31;
      if (r0 + lane <= r && r0 + lane >= r_begin) {
        float sc = key_score(keep);
        if (row_term != nullptr) sc = fmaxf(row_term[q] - sc, 0.f);
        out_scores[static_cast<size_t>(q) * ld_out + r0 + lane] = sc;
        out_ids[static_cast<size_t>(q) * ld_out + r0 + lane] = static_cast<long long>(key_index(keep)) + id_base;
      }
    }
    prev = best;
  }
  // picks of an incomplete group of 32 (the loop stopped early)
  if (r < k && (r & 31) != 0) {
    const int r0 = r & ~31;
    if (r0 + lane < r && r0 + lane >= r_begin) {
      float sc = key_score(keep);
      if (row_term != nullptr) sc = fmaxf(row_term[q] - sc, 0.f);
      out_scores[static_cast<size_t>(q) * ld_out + r0 + lane] = sc;
      out_ids[static_cast<size_t>(q) * ld_out + r0 + lane] = static_cast<long long>(key_index(keep)) + id_base;
    }
  }
  if (exhausted) {  // fewer than k rows exist: pad like the single-pass select
    for (int rr = r + lane; rr < k; rr += 32) {
      out_scores[static_cast<size_t>(q) * ld_out + rr] = row_term != nullptr ? INFINITY : -INFINITY;
      out_ids[static_cast<size_t>(q) * ld_out + rr] = -1ll;
    }
    r = k;
  }
  if (lane == 0) {
    done[q] = r;
    bound[q] = r >= k ? 0ull : prev;  // 0: nothing is eligible any more
    if (r < k) atomicAdd(remaining_out, 1u);
  }
}

// (score, id) pair order: a is better than b
__device__ __forceinline__ bool pair_better(float sa, long long ia, float sb, long long ib) {
  return (sa > sb) || (sa == sb && ia < ib);
}

// Merge g lists per claim: scores [g][nq][k], ids [g][nq][k] (id < 0 = empty) -> [nq][k].
__global__ void __launch_bounds__(128)
merge_pairs_kernel(const float* __restrict__ in_scores, const long long* __restrict__ in_ids, int g, int nq, int k,
                   float* __restrict__ out_scores, long long* __restrict__ out_ids) {
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= nq) return;
  const int ncand = g * k;
  float ps = INFINITY;
  long long pi = -1;  // previous pick; (inf, -1) is better than every real candidate
  for (int r = 0; r < k; ++r) {
    float bs = -INFINITY;
    long long bi = -1;  // -1 = none yet
    for (int c = lane; c < ncand; c += 32) {
      const int shard = c / k, j = c - shard * k;
      const size_t off = (static_cast<size_t>(shard) * nq + q) * k + j;
      const long long id = in_ids[off];
      if (id < 0) continue;
      const float sc = in_scores[off];
      if (!pair_better(ps, pi, sc, id)) continue;  // not strictly worse than the previous pick
      if (bi < 0 || pair_better(sc, id, bs, bi)) { bs = sc; bi = id; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, bs, o);
      const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oi >= 0 && (bi < 0 || pair_better(os, oi, bs, bi))) { bs = os; bi = oi; }
    }
    if (lane == 0) {
      out_scores[static_cast<size_t>(q) * k + r] = bi >= 0 ? bs : -INFINITY;
      out_ids[static_cast<size_t>(q) * k + r] = bi;
    }
    if (bi < 0) {  // exhausted: pad the rest
      for (int rr = r + 1 + lane; rr < k; rr += 32) {
        out_scores[static_cast<size_t>(q) * k + rr] = -INFINITY;
        out_ids[static_cast<size_t>(q) * k + rr] = -1;
      }
      break;
    }
    ps = bs;
    pi = bi;
  }
}

// Everything a scan needs staged, in ONE launch instead of four stream operations (each costs a few
// microseconds of launch latency, which is what a small-batch search over a shard is made of): zero the
// round-barrier counter and the threshold seeds, and build the zero-padded copy of the claims.
// All regions are 16-byte aligned and a multiple of 16 bytes long (256-byte aligned workspace carve-up).
__global__ void __launch_bounds__(256)
scan_prep_kernel(uint4* __restrict__ round_counter, uint4* __restrict__ seeds, size_t seed_vec,
                 const uint4* __restrict__ claims, uint4* __restrict__ pad, size_t live_vec, size_t pad_vec) {
  const size_t tid = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  if (round_counter != nullptr && tid < 16) round_counter[tid] = zero;  // 256 bytes
  for (size_t i = tid; i < seed_vec; i += stride) seeds[i] = zero;
  for (size_t i = tid; i < pad_vec; i += stride) pad[i] = i < live_vec ? claims[i] : zero;
}

// fp32 -> (hi, lo) for the 3 x TF32 GEMM (gemm_tc.cuh, PREC = 1): hi = x with the 13 low mantissa bits cleared -- the
// part of the word a kind::tf32 MMA reads -- and lo = x - hi, exact in fp32.  Vectors [live, total) are zero (row
// padding of the claims).  hi may be null: the caller then feeds the original words as the hi operand.
__global__ void __launch_bounds__(256)
split_f32_kernel(const float4* __restrict__ src, size_t live_vec, size_t total_vec, float4* __restrict__ hi, float4* __restrict__ lo) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total_vec; i += stride) {
    const float4 x = i < live_vec ? src[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 h;
    h.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
    h.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
    h.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u);
    h.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
    if (hi != nullptr) hi[i] = h;
    lo[i] = make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w);
  }
}

// |row|^2 in fp32 for fp32 or bf16 rows; out[r] = sign * |row|^2.  One warp per row.
template <typename T>
__global__ void row_sqnorm_kernel(const T* __restrict__ x, long long rows, int dim, float sign, float* __restrict__ out) {
  const long long row = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float acc = 0.f;
  for (int d = threadIdx.x & 31; d < dim; d += 32) {
    const float v = static_cast<float>(x[row * dim + d]);
    acc = fmaf(v, v, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) out[row] = sign * acc;
}

}  // namespace drs
