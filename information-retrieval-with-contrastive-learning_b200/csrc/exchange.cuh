// Fused select + exchange + merge for a row-sharded corpus: ONE kernel after the scan, over NVLink
// peer memory ("all-gather of per-shard top-k lists followed by an on-GPU merge").
//
// The unfused form is four launches and two collectives per search: select (per-split candidate
// keys -> this shard's top-k), ncclAllGather of the scores, ncclAllGather of the ids, merge.  The
// payload is tiny (10 000 claims x top-10 x 12 B = 1.2 MB per rank), so the step is pure launch and
// collective latency -- negligible next to a 40 ms tensor-bound scan, but 10 % of a step in the
// small-batch, HBM-bound regime where a shard is streamed in under a millisecond.
//
// Here every rank runs the same persistent kernel.  For each block of 32 claims:
//   1. select: one warp per claim reduces the candidate keys to the shard's top-k (pick r lands in
//      lane r) and stores the list straight into slot [rank] of EVERY peer's gather buffer -- plain
//      st.global through the NVLink peer mapping, 128-byte lines;
//   2. publish: __syncthreads, fence.sys, then st.release.sys of the call's epoch into flag
//      [block][rank] of every peer;
//   3. wait: ld.acquire.sys until all `world` flags of this claim block show the epoch -- the lists
//      of this block have arrived from every shard (no global barrier: blocks proceed independently);
//   4. merge: the warp orders the world * k candidates by (score desc, id asc) and writes the result.
// Flags carry a monotonically increasing epoch (a per-rank device counter of completed calls, so a captured
// CUDA graph replays correctly) and are never reset, so there is no cleanup pass and a change of batch size
// between calls is harmless.  The gather buffers are double-buffered by
// epoch parity: a rank can be at most one call ahead of its slowest peer (it cannot finish call n+1
// before every peer has published call n+1, i.e. finished reading call n).
// All CTAs are co-resident (grid <= resident capacity) and every block publishes before it waits,
// so the spin cannot deadlock; a wait that exceeds ~1.5 minutes traps with a hang report.
#pragma once
#include "merge.cuh"
#include "ptx.cuh"

namespace drs {

constexpr int kMaxPeers = 8;
enum : uint32_t { kTagExchangeWait = 6 };
// A peer's kernel starts whenever ITS host thread gets to the launch, so the wait must tolerate host-side skew
// between ranks (seconds), not just device latency; it still ends a genuinely stuck exchange (~1.5 min).
static constexpr long long kExchangeTimeoutCycles = 180000000000LL;

struct ExchangePeers {
  float* scores[kMaxPeers];     // rank p's gather buffers, parity 0: [world][nq][k] fp32   (peer-mapped pointers)
  long long* ids[kMaxPeers];    //                                    [world][nq][k] int64
  uint32_t* flags[kMaxPeers];   // rank p's flags:                    [claim blocks][world]
  size_t parity_stride;         // bytes from a parity-0 buffer to its parity-1 twin (same for scores and ids)
  const uint32_t* calls;        // this rank's device counter of completed exchange calls: epoch = *calls + 1
};

__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

constexpr int kExchangeRowsPerBlock = 8;  // 8 warps, one claim each

// ws: [nq][nslots][kcap] candidate keys of this shard (the scan's output: sorted runs).  k <= 32.
// SL: runs per lane of the k-way select (nslots <= 32 * SL); 0 = unsorted re-scan of all candidates per pick.
template <int SL>
__global__ void __launch_bounds__(256)
select_exchange_merge_kernel(const uint64_t* __restrict__ ws, int nq, int nslots, int kcap, int k, long long id_base,
                             ExchangePeers peers, int rank, int world,
                             float* __restrict__ out_scores, long long* __restrict__ out_ids) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // The epoch lives on the device (bumped by exchange_done_kernel after this kernel), so the call sequence can be
  // captured in a CUDA graph and replayed: no launch argument changes from call to call.
  const uint32_t epoch = *peers.calls + 1u;
  if (epoch & 1u) {
#pragma unroll
    for (int p = 0; p < kMaxPeers; ++p) {
      peers.scores[p] = reinterpret_cast<float*>(reinterpret_cast<char*>(peers.scores[p]) + peers.parity_stride);
      peers.ids[p] = reinterpret_cast<long long*>(reinterpret_cast<char*>(peers.ids[p]) + peers.parity_stride);
    }
  }
  const int num_blocks = (nq + kExchangeRowsPerBlock - 1) / kExchangeRowsPerBlock;
  const size_t slot = static_cast<size_t>(nq) * k;  // one rank's [nq][k] list
  for (int rb = blockIdx.x; rb < num_blocks; rb += gridDim.x) {
    const int q = rb * kExchangeRowsPerBlock + warp;
    // ---- 1. select this shard's top-k and scatter it to every peer
    if (q < nq) {
      const uint64_t* src = ws + static_cast<size_t>(q) * nslots * kcap;
      uint64_t mine = 0ull;
      if constexpr (SL > 0) {
        mine = warp_select_runs<SL>(src, nslots, kcap, k, lane);
      } else {
        uint64_t prev = ~0ull;
        for (int r = 0; r < k; ++r) {
          uint64_t best = 0ull;
          for (int c = lane; c < nslots * kcap; c += 32) {
            const uint64_t key = __ldg(src + c);
            if (key < prev && key > best) best = key;
          }
          best = warp_max_u64(best);
          if (lane == r) mine = best;
          prev = best;
        }
      }
      if (lane < k) {
        const float sc = mine ? key_score(mine) : -INFINITY;
        const long long id = mine ? static_cast<long long>(key_index(mine)) + id_base : -1ll;
        const size_t off = static_cast<size_t>(rank) * slot + static_cast<size_t>(q) * k + lane;
        for (int p = 0; p < world; ++p) {
          peers.scores[p][off] = sc;
          peers.ids[p][off] = id;
        }
      }
    }
    __syncthreads();
    if (warp == 0) {
      // ---- 2. publish (the barrier above ordered the block's stores before this fence)
      __threadfence_system();
      if (lane < world) st_release_sys_u32(peers.flags[lane] + static_cast<size_t>(rb) * world + rank, epoch);
      // ---- 3. wait for this claim block's lists from every shard
      const uint32_t* fl = peers.flags[rank] + static_cast<size_t>(rb) * world;
      const long long t0 = clock64();
      for (;;) {
        const uint32_t v = lane < world ? ld_acquire_sys_u32(fl + lane) : epoch;
        if (__all_sync(0xffffffffu, static_cast<int32_t>(v - epoch) >= 0)) break;
        __nanosleep(200);
        if (clock64() - t0 > kExchangeTimeoutCycles) mbar_hang(kTagExchangeWait, epoch, static_cast<uint32_t>(rb));
      }
    }
    __syncthreads();
    // ---- 4. merge world * k candidates (L2 reads: peer stores land in this GPU's L2, never in its L1)
    const float* in_s = peers.scores[rank];
    const long long* in_i = peers.ids[rank];
    const int total = world * k;
    if (q < nq) {
      float ps = INFINITY;
      long long pi = -1;
      // world * k <= 8 * 32 candidates: 8 per lane, read once (id < 0 = empty slot)
      float cs[kMaxPeers];
      long long ci[kMaxPeers];
#pragma unroll
      for (int u = 0; u < kMaxPeers; ++u) {
        const int c = lane + 32 * u;
        ci[u] = -1;
        cs[u] = -INFINITY;
        if (c < total) {
          const int shard = c / k, j = c - shard * k;
          const size_t off = static_cast<size_t>(shard) * slot + static_cast<size_t>(q) * k + j;
          ci[u] = __ldcg(in_i + off);
          cs[u] = __ldcg(in_s + off);
        }
      }
      for (int r = 0; r < k; ++r) {
        float bs = -INFINITY;
        long long bi = -1;
#pragma unroll
        for (int u = 0; u < kMaxPeers; ++u) {
          if (ci[u] < 0) continue;
          if (!pair_better(ps, pi, cs[u], ci[u])) continue;  // not strictly worse than the previous pick
          if (bi < 0 || pair_better(cs[u], ci[u], bs, bi)) { bs = cs[u]; bi = ci[u]; }
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
        if (bi < 0) {
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
  }
}

// One more exchange call is complete on this rank (stream-ordered after select_exchange_merge_kernel).
__global__ void exchange_done_kernel(uint32_t* calls) { *calls += 1u; }

// =====================================================================================================
// Query-sliced exchange for any k ("for cfg 5 prefer query-sliced all-to-all + local merge +
// all-gather of final (Q/g, k)").  With k = 100 and 65 536 claims a rank's lists are 78 MB; gathering every
// rank's lists everywhere moves world x 78 MB into each GPU and makes every GPU merge all 65 536 claims.  Sliced:
// rank s owns claims [s * per, (s + 1) * per), per = ceil(nq / world).
//   phase 1  scatter: each rank stores the lists of slice s straight into rank s's gather buffer (NVLink peer
//            stores, full lines) and publishes per-(claim block, source) flags on rank s -- no wait anywhere;
//   phase 2  merge:   rank s waits (per block of its slice) for the flags of all sources, merges the `world` sorted
//            runs of each claim (k-way merge, one warp per claim, runs staged in shared memory), and stores the
//            final list into EVERY rank's result buffer, then publishes per-block flags on every rank;
//   phase 3  collect: each rank waits for the result flags of all blocks of all slices and copies the rows to the
//            caller's tensors.
// Bytes over NVLink per rank: (world-1)/world x nq x k x 12 out in phase 1 (8x less than the all-gather at 8
// ranks) + the final (nq / world) x k x 12 to each peer.  Epochs, parity double-buffering and the call counter
// work as in the kernel above.  Phase 1 never waits, phase 2 waits only on phase-1 flags and phase 3 only on
// phase-2 flags, so with all CTAs resident (cooperative launch) the kernel cannot deadlock.
struct SlicedPeers {
  char* base[kMaxPeers];   // rank p's symmetric buffer (peer-mapped)
  size_t flags1_off;       // u32 [blocks per slice][world]   written by sources, read by the slice owner
  size_t flags2_off;       // u32 [world * blocks per slice]  written by slice owners, read by everyone
  size_t data_off;         // parity 0 data; parity 1 at + parity_stride
  size_t parity_stride;
  size_t gather_i_off;     // within a parity block: gather scores at 0, gather ids here, ...
  size_t out_s_off;
  size_t out_i_off;
  const uint32_t* calls;
};

constexpr int kSlicedRowsPerBlock = 8;   // claims per block iteration: one per warp
enum : uint32_t { kTagSlicedWait1 = 7, kTagSlicedWait2 = 8 };

__device__ __forceinline__ void wait_flags_sys(const uint32_t* fl, int count, uint32_t epoch, uint32_t tag, uint32_t extra) {
  // one warp; lane l < count watches fl[l]
  const int lane = threadIdx.x & 31;
  const long long t0 = clock64();
  for (;;) {
    bool ok = true;
    for (int c = lane; c < count; c += 32) ok = ok && static_cast<int32_t>(ld_acquire_sys_u32(fl + c) - epoch) >= 0;
    if (__all_sync(0xffffffffu, ok)) break;
    __nanosleep(200);
    if (clock64() - t0 > kExchangeTimeoutCycles) mbar_hang(tag, epoch, extra);
  }
}

__global__ void __launch_bounds__(256)
exchange_sliced_kernel(const float* __restrict__ local_s, const long long* __restrict__ local_i, int nq, int k,
                       SlicedPeers peers, int rank, int world, float* __restrict__ out_scores,
                       long long* __restrict__ out_ids) {
  extern __shared__ __align__(16) unsigned char xs_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t epoch = *peers.calls + 1u;
  const size_t par = peers.data_off + ((epoch & 1u) ? peers.parity_stride : 0);
  const int per = (nq + world - 1) / world;                                   // claims per slice
  const int bps = (per + kSlicedRowsPerBlock - 1) / kSlicedRowsPerBlock;      // claim blocks per slice
  const size_t run = static_cast<size_t>(per) * k;                            // one source's lists of one slice
  const bool vec_in = (k & 3) == 0 && (reinterpret_cast<uintptr_t>(local_s) & 15) == 0 && (reinterpret_cast<uintptr_t>(local_i) & 15) == 0;

  // ---- phase 1: scatter my lists, slice by slice, into the owners' gather buffers
  for (int gb = blockIdx.x; gb < world * bps; gb += gridDim.x) {
    const int dest = gb / bps, b = gb - dest * bps;
    const int ql = b * kSlicedRowsPerBlock + warp;                            // claim within the slice
    const int q = dest * per + ql;
    if (ql < per && q < nq) {
      float* gs = reinterpret_cast<float*>(peers.base[dest] + par);
      long long* gi = reinterpret_cast<long long*>(peers.base[dest] + par + peers.gather_i_off);
      const size_t dst = static_cast<size_t>(rank) * run + static_cast<size_t>(ql) * k;
      const size_t src = static_cast<size_t>(q) * k;
      if (vec_in) {         // rows are 16-byte multiples: 16-byte peer stores (4 scores / 2 ids per lane), 512 bytes per instruction
        const uint4* ss = reinterpret_cast<const uint4*>(local_s + src);
        const uint4* si = reinterpret_cast<const uint4*>(local_i + src);
        uint4* ds = reinterpret_cast<uint4*>(gs + dst);
        uint4* di = reinterpret_cast<uint4*>(gi + dst);
        for (int j = lane; j < k / 4; j += 32) ds[j] = __ldcg(ss + j);
        for (int j = lane; j < k / 2; j += 32) di[j] = __ldcg(si + j);
      } else {
        for (int j = lane; j < k; j += 32) {
          gs[dst + j] = __ldcg(local_s + src + j);
          gi[dst + j] = __ldcg(local_i + src + j);
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence_system();
      st_release_sys_u32(reinterpret_cast<uint32_t*>(peers.base[dest] + peers.flags1_off) + static_cast<size_t>(b) * world + rank, epoch);
    }
  }

  // ---- phase 2: merge the slice I own
  float* ms = reinterpret_cast<float*>(xs_smem) + static_cast<size_t>(warp) * world * k;                       // [world][k] per warp
  long long* mi = reinterpret_cast<long long*>(xs_smem + static_cast<size_t>(8) * world * k * sizeof(float)) +
                  static_cast<size_t>(warp) * world * k;
  const float* my_gs = reinterpret_cast<const float*>(peers.base[rank] + par);
  const long long* my_gi = reinterpret_cast<const long long*>(peers.base[rank] + par + peers.gather_i_off);
  for (int b = blockIdx.x; b < bps; b += gridDim.x) {
    if (warp == 0)
      wait_flags_sys(reinterpret_cast<const uint32_t*>(peers.base[rank] + peers.flags1_off) + static_cast<size_t>(b) * world, world,
                     epoch, kTagSlicedWait1, static_cast<uint32_t>(b));
    __syncthreads();
    const int ql = b * kSlicedRowsPerBlock + warp;
    const int q = rank * per + ql;
    if (ql < per && q < nq) {
      for (int c = lane; c < world * k; c += 32) {                      // stage the runs (L2 reads: peers stored them)
        const int srcr = c / k, j = c - srcr * k;
        const size_t off = static_cast<size_t>(srcr) * run + static_cast<size_t>(ql) * k + j;
        ms[c] = __ldcg(my_gs + off);
        mi[c] = __ldcg(my_gi + off);
      }
      __syncwarp();
      // lane l < world owns run l: head at position pos
      int pos = 0;
      float hs = -INFINITY;
      long long hi = -1;
      if (lane < world) { hs = ms[lane * k]; hi = mi[lane * k]; }
      float keep_s = -INFINITY;
      long long keep_i = -1;
      for (int r0 = 0; r0 < k; r0 += 32) {
        const int nr = min(32, k - r0);
        for (int r = 0; r < nr; ++r) {
          float bs = hs;
          long long bi = hi;
#pragma unroll
          for (int o = 4; o > 0; o >>= 1) {                               // runs live in lanes 0..7
            const float os = __shfl_xor_sync(0xffffffffu, bs, o);
            const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (oi >= 0 && (bi < 0 || pair_better(os, oi, bs, bi))) { bs = os; bi = oi; }
          }
          bs = __shfl_sync(0xffffffffu, bs, 0);
          bi = __shfl_sync(0xffffffffu, bi, 0);
          if (lane == r) { keep_s = bi >= 0 ? bs : -INFINITY; keep_i = bi; }
          if (bi >= 0 && lane < world && hi == bi && hs == bs) {           // (score, id) pairs are unique: one lane advances
            ++pos;
            if (pos < k) { hs = ms[lane * k + pos]; hi = mi[lane * k + pos]; }
            else { hs = -INFINITY; hi = -1; }
          }
        }
        if (lane < nr) {                                                  // 32 picks: one coalesced store per peer
          const size_t off = static_cast<size_t>(q) * k + r0 + lane;
          for (int p = 0; p < world; ++p) {
            reinterpret_cast<float*>(peers.base[p] + par + peers.out_s_off)[off] = keep_s;
            reinterpret_cast<long long*>(peers.base[p] + par + peers.out_i_off)[off] = keep_i;
          }
        }
        keep_s = -INFINITY;
        keep_i = -1;
      }
    }
    __syncthreads();
    if (threadIdx.x < world) {
      __threadfence_system();
      st_release_sys_u32(reinterpret_cast<uint32_t*>(peers.base[threadIdx.x] + peers.flags2_off) + static_cast<size_t>(rank) * bps + b, epoch);
    }
  }

  // ---- phase 3: collect every slice's results from my result buffer into the caller's tensors
  const float* rs = reinterpret_cast<const float*>(peers.base[rank] + par + peers.out_s_off);
  const long long* ri = reinterpret_cast<const long long*>(peers.base[rank] + par + peers.out_i_off);
  for (int gb = blockIdx.x; gb < world * bps; gb += gridDim.x) {
    if (warp == 0)
      wait_flags_sys(reinterpret_cast<const uint32_t*>(peers.base[rank] + peers.flags2_off) + gb, 1, epoch, kTagSlicedWait2,
                     static_cast<uint32_t>(gb));
    __syncthreads();
    const int s = gb / bps, b = gb - s * bps;
    const int q0 = s * per + b * kSlicedRowsPerBlock;
    const int q1 = min(min(q0 + kSlicedRowsPerBlock, (s + 1) * per), nq);
    if (q1 > q0) {
      const size_t lo = static_cast<size_t>(q0) * k, n = static_cast<size_t>(q1 - q0) * k;
      if ((k & 3) == 0 && (reinterpret_cast<uintptr_t>(out_scores) & 15) == 0 && (reinterpret_cast<uintptr_t>(out_ids) & 15) == 0) {
        const uint4* ss = reinterpret_cast<const uint4*>(rs + lo);
        const uint4* si = reinterpret_cast<const uint4*>(ri + lo);
        uint4* ds = reinterpret_cast<uint4*>(out_scores + lo);
        uint4* di = reinterpret_cast<uint4*>(out_ids + lo);
        for (size_t e = threadIdx.x; e < n / 4; e += blockDim.x) ds[e] = __ldcg(ss + e);
        for (size_t e = threadIdx.x; e < n / 2; e += blockDim.x) di[e] = __ldcg(si + e);
      } else {
        for (size_t e = threadIdx.x; e < n; e += blockDim.x) {
          out_scores[lo + e] = __ldcg(rs + lo + e);
          out_ids[lo + e] = __ldcg(ri + lo + e);
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace drs
