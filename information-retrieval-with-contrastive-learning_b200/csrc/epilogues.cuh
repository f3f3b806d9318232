// Per-row epilogue functors plugged into the GEMM mainloops (gemm_tc.cuh, gemm_simt.cuh).
// Each functor instance lives in the registers of ONE thread that owns ONE row of A for a whole
// unit (a contiguous range of B rows), sees the scores of that row in increasing column order,
// 32 at a time, and flushes once at the end of the unit.
#pragma once
#include "topk.cuh"

namespace drs {

// ---------------------------------------------------------------------------------------------
// Running top-K of one claim over a range of corpus rows.  Replaces the select of
// TfidfDocRanker.closest_docs (preprocessing/drqa/retriever/tfidf_doc_ranker.py:67-73) for the
// dense scores of src/evaluation.py:110-115.  Output: KCAP packed keys per (claim, split) in the
// workspace; merge.cuh reduces the splits to the final k.
constexpr int kMaxSeedSlots = 96;   // >= ceil(DRS_MAX_K / 3): certificates kept per claim (plan_search)

template <int KCAP, bool BIASED = false>
struct TopKEpilogue {
  struct Params {
    uint64_t* ws;  // [rows_a][num_slots][KCAP]   slot = split * (column groups per tile) + group
    int rows_a;
    int rows_b;
    int num_slots;
    int k;
    const uint64_t* bound;  // optional [rows_a]: only candidates with key < bound[row] are eligible
                            // (k > KCAP is served in passes: each pass continues below the last pick)
    const float* col_bias;  // BIASED: ranked value = score_scale * dot + col_bias[col]
    float score_scale;      //   (squared-L2 search: 2 x.c - |c|^2, src/contrastor/utils.py:64-67)
    // Threshold seeding across units.  seeds: [rows_a][1 + seed_slots] ordered-uint32 scores, zeroed before the
    // scan: word 0 is the claim's current FLOOR, words 1.. are certificates.  A unit's sorted list is cut into
    // seed_chunks groups of seed_group entries; group c's smallest entry v_c = sc[c * seed_group - 1] certifies
    // seed_group rows of the unit's corpus range that score >= v_c, disjoint from every other group of every unit.
    // Each claim keeps the seed_slots largest certificates published so far (replace-the-minimum by
    // compare-and-swap); with seed_slots * seed_group >= (the k the caller wants) their minimum is a lower bound on
    // the final k-th best score, published with atomicMax into the floor word.  Later units of the claim start
    // from it instead of from -inf and stay on the fast path (a fresh list needs ~k ln(n/k) inserts to warm up, and
    // one lane's insert stalls its warp).  Every certified row is in a candidate list, so the select still sees >= k
    // rows at or above the bound; stale or missing seeds only weaken it -- the result is exact either way.
    // Cost per unit: ONE load at the start; at the end one batch of independent loads (the certificates), the
    // replace-the-minimum bookkeeping on a private copy, and up to seed_chunks fire-and-forget CAS + one atomicMax.
    // (The first version walked the slots with dependent L2 round trips -- 34 slots x 5 certificates for top-100 --
    // and that walk, at every unit end, was what held the top-100 scan at 73 % tensor-pipe activity.)
    uint32_t* seeds;
    int seed_slots;
    int seed_group;
    int seed_chunks;
  };
  static constexpr bool kUsesScratch = false;
  static constexpr bool kStagesColumns = false;
  TopKList<KCAP> list;
  uint64_t bnd;

  __device__ __forceinline__ void begin_unit(const Params& p, int row, int, int) {
    float floor = -INFINITY;
    bnd = (p.bound != nullptr && row < p.rows_a) ? p.bound[row] : ~0ull;
    if (bnd == 0ull) {
      floor = INFINITY;  // this claim is complete: nothing is eligible, stay on the fast path
    } else if (p.seeds != nullptr && row < p.rows_a) {
      const uint32_t lo = __ldcg(p.seeds + static_cast<size_t>(row) * (p.seed_slots + 1));
      // strictly below the bound, so that rows TYING it (with a lower index) still enter.  The predecessor of
      // +0.0 in the ordered domain decodes to -0.0, which `s > thr` cannot tell from +0.0: step once more
      // (to the largest negative float), or rows scoring exactly 0 -- zero-padded corpus rows, all-zero
      // claims -- would lose their tie to a later unit.
      if (lo != 0u) floor = ordered_to_float(lo == 0x80000000u ? lo - 2u : lo - 1u);
    }
    list.reset(floor);
  }

  __device__ __forceinline__ void chunk(const Params& p, int row, int col0, const uint32_t (&raw)[32]) {
    if constexpr (BIASED) {
      uint32_t t[32];
      const int nv = p.rows_b - col0;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float b = (j < nv) ? __ldg(p.col_bias + col0 + j) : 0.f;
        t[j] = __float_as_uint(fmaf(p.score_scale, __uint_as_float(raw[j]), b));
      }
      scan(p, row, col0, t);
    } else {
      scan(p, row, col0, raw);
    }
  }

  __device__ __forceinline__ void scan(const Params& p, int /*row*/, int col0, const uint32_t (&v)[32]) {
    const int valid = p.rows_b - col0;  // columns >= rows_b are TMA zero fill, not corpus rows
    float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
    if (valid >= 32) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        m0 = fmaxf(m0, __uint_as_float(v[j + 0]));
        m1 = fmaxf(m1, __uint_as_float(v[j + 1]));
        m2 = fmaxf(m2, __uint_as_float(v[j + 2]));
        m3 = fmaxf(m3, __uint_as_float(v[j + 3]));
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) m0 = (j < valid) ? fmaxf(m0, __uint_as_float(v[j])) : m0;
    }
    const float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
    // fast path: nothing in this chunk beats the current k-th best of any row of the warp
    if (!__any_sync(0xffffffffu, mx > list.thr)) return;
    // Slow path (rare once the lists have warmed up): each lane walks ITS OWN candidate columns; the warp
    // iterates max-over-lanes times (lanes that are done idle), not once per column of the union.  The insert
    // code exists once (a 32-way unrolled version thrashes the instruction cache); pick32 selects v[j] for a
    // per-lane j without dynamic register indexing.
    uint32_t mine = 0u;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if ((j < valid) && (__uint_as_float(v[j]) > list.thr)) mine |= 1u << j;
#pragma unroll 1
    while (mine) {
      const int j = __ffs(mine) - 1;
      mine &= mine - 1u;
      const float s = __uint_as_float(pick32(v, j));
      const bool hit = (s > list.thr) && (make_key(s, static_cast<uint32_t>(col0 + j)) < bnd);
      list.insert(hit ? s : -INFINITY, static_cast<uint32_t>(col0 + j), p.k);
    }
    __syncwarp();  // lanes leave the loop at different times; the caller's next tcgen05.ld is .sync.aligned
  }

  __device__ __forceinline__ void end_unit(const Params& p, int row, int, int slot) {
    if (row >= p.rows_a) return;
    uint64_t* dst = p.ws + (static_cast<size_t>(row) * p.num_slots + slot) * KCAP;
#pragma unroll
    for (int j = 0; j < KCAP; ++j) dst[j] = list.key(j);
    if (p.seeds != nullptr) {
      uint32_t* a = p.seeds + static_cast<size_t>(row) * (p.seed_slots + 1);
      uint32_t cert[5];                                   // this unit's certificates, best first (seed_chunks <= 5)
      int ncert = 0;
#pragma unroll
      for (int c = 1; c <= 5; ++c) {
        const int pos = c * p.seed_group - 1;
        float val = -INFINITY;
#pragma unroll
        for (int j = 0; j < KCAP; ++j) val = (j == pos) ? list.sc[j] : val;
        // empty, or not above what the claim already had when the unit started: nothing to add (values descend)
        const bool live = c <= p.seed_chunks && val > list.floor && ncert == c - 1;
        cert[c - 1] = live ? float_to_ordered(val) : 0u;
        ncert += live ? 1 : 0;
      }
      if (ncert == 0) return;
      uint32_t sl[kMaxSeedSlots];                         // private copy of the claim's certificates
      for (int j = 0; j < p.seed_slots; j += 8) {         // independent loads: one L2 round trip per 8 slots
        uint32_t x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = (j + i < p.seed_slots) ? __ldcg(a + 1 + j + i) : 0xFFFFFFFFu;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (j + i < p.seed_slots) sl[j + i] = x[i];
      }
      for (int c = 0; c < ncert; ++c) {
        uint32_t mn = 0xFFFFFFFFu;
        int mi = 0;
        for (int j = 0; j < p.seed_slots; ++j)
          if (sl[j] < mn) { mn = sl[j]; mi = j; }
        if (cert[c] <= mn) break;                         // not among the seed_slots largest (nor are the smaller ones)
        // fire and forget: a lost race only drops this certificate from the shared array (a weaker bound for
        // others); the private copy stays a set of valid, disjoint certificates either way
        atomicCAS(a + 1 + mi, mn, cert[c]);
        sl[mi] = cert[c];
      }
      uint32_t mn = 0xFFFFFFFFu;
      for (int j = 0; j < p.seed_slots; ++j) mn = min(mn, sl[j]);
      if (mn != 0u) atomicMax(a, mn);                     // all slots filled: their minimum bounds the k-th best
    }
  }
};

}  // namespace drs
