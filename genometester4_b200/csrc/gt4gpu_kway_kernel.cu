// gt4gpu_kway_kernel.cu -- single-pass N-list union / intersection (N <= 8 per pass).
//
// Replaces the loops of union_multi / intersect_multi (/root/reference/src/glistcompare.c:545-591, :647-705) and of
// gt4_write_union (src/set-operations.c:77-116): every list is read ONCE and the result written once, instead of a
// tree of two-list passes with materialised intermediates.
//
// Decomposition (all by KEY, so equal words of different lists always meet in one place):
//
//   kway_sample_kernel    every KW_SAMPLE-th word of every list -> one array; it is radix-sorted with the list-building
//                         sort (gt4gpu_sort_kernel.cu)
//   kway_cuts_kernel      every m-th sorted sample is a tile boundary key; one lower-bound search per (boundary, list)
//                         gives the tile's slice of every list.  A tile holds at most (m + 2 N) * KW_SAMPLE records.
//   kway_tile_kernel      persistent, warp-specialised, mbarrier pipeline like setop2_stream_kernel:
//       producer warp     claims tiles by ticket, stages the N key slices + N count slices with 1-D TMA bulk copies
//                         (one lane per slice; slices are packed so that keys and counts share ONE index space)
//       consumers         (1) cut the tile's key range [lo, hi] into one bucket per consumer thread -- bucket =
//                         (key - lo) * NB / (hi - lo + 1), one multiply per record, no search -- and record in a small
//                         table where every bucket starts in every slice; (2) every thread merges its bucket with the
//                         N slice heads in registers: the smallest head is the next distinct word, all lists holding it
//                         advance together and their counts are folded by the rule (add / max / min with the
//                         reference's "!freq ||" guard / number); survivors of the cut-off go to a scratch buffer at
//                         the thread's input offset; (3) block scan of the survivor counts, hand the tile total to the
//                         look-back warp, compact the survivors to the front of the stage buffer
//       look-back warps   one per stage: decoupled look-back for the tile's global output offset
//       store warps       coalesced copy of the compacted tile to the output arrays, stage back to the producer
//
// The bucket split is exact in keys and approximate in load: it assumes the words of a tile (a key range only a few
// thousand words wide) are spread about evenly; a skewed tile is still merged correctly, by fewer threads.
//
// HBM-bound integer work: no tensor cores.  Algorithmic traffic 12 B per input record + 12 B per output record.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "gt4gpu_device.cuh"
#include "gt4gpu_internal.h"

#ifndef GT4_KWAY_CTAS
#define GT4_KWAY_CTAS 1            // CTAs per SM (2 needs GT4_KWAY_CAP <= 2048)
#endif

namespace gt4gpu {

namespace {

using namespace dev;

constexpr int KW_STORE_WARPS = 2;
constexpr uint64_t KW_TILE_END = ~0ull;
constexpr int KW_LB_W = 4;

template <int NL, int NC, int S>
struct KwayCfg {
  static constexpr int NWARPS = NC / 32;
  static constexpr int PRODUCER_WARP = NWARPS;
  static constexpr int LOOKBACK_WARP0 = NWARPS + 1;
  static constexpr int STORE_WARP0 = LOOKBACK_WARP0 + S;
  static constexpr int NTHREADS = NC + 32 + 32 * S + 32 * KW_STORE_WARPS;
  static constexpr int NB = NC;                                  // one bucket per consumer thread
  // every slice starts on a 4-slot boundary plus its count phase (0..3) and is padded to a multiple of 4 slots
  static constexpr int SLOTS = KWAY_TILE_CAP + 8 * NL + 8;
  static constexpr size_t STAGE_BYTES = (size_t) SLOTS * 12;
  static constexpr size_t SPARSE_BYTES = (size_t) KWAY_TILE_CAP * 12;
  static constexpr size_t HIST_BYTES = (size_t) (NL * NB + 32) * 4;     // records of every slice in every bucket (+ one dummy bucket per lane)
  static constexpr size_t SMEM_BYTES = S * STAGE_BYTES + SPARSE_BYTES + HIST_BYTES;
  static constexpr int LOG2_NB = NB == 128 ? 7 : NB == 256 ? 8 : 9;
  static constexpr int BUCKET_SHIFT = 30 - LOG2_NB;               // bucket = umulhi (u, mult) >> BUCKET_SHIFT, mult < 2^32
  static_assert (NB == 128 || NB == 256 || NB == 512, "bucket count");
  static_assert (SLOTS % 4 == 0, "the count region must start 16-byte aligned");
  static_assert (SLOTS < 65536, "slot indices are kept as u16");
};

template <int NL>
struct KwMeta {
  uint64_t tile;      // KW_TILE_END: no more work
  uint64_t lo;        // smallest key the tile can hold
  uint64_t hi;        // largest key the tile can hold
  uint32_t mult;      // bucket = umulhi (u, mult) >> BUCKET_SHIFT with u = high word of (key - lo) << lsh
  int lsh;            // leading zeros of hi - lo: the tile's key span is normalised to 32 significant bits
  int n_total;
  uint16_t idx0[NL];  // slot of the slice's first record (keys and counts share the index space)
  uint16_t n[NL];
};

struct KwMailbox {
  uint64_t tile;
  uint64_t base;
  int cnt;
};

// Bucket of a key: the tile's key span [lo, hi] is normalised to 32 significant bits (u) and scaled to NB buckets with
// one multiply; monotone in the key, < NB for lo <= key <= hi (the clamp only matters for lists that are not sorted).
template <int SHIFT>
__device__ __forceinline__ unsigned kw_bucket (uint64_t key, uint64_t lo, int lsh, uint32_t mult, unsigned nb)
{
  const uint64_t d = key - lo;
  const uint32_t dl = (uint32_t) d, dh = (uint32_t) (d >> 32);
  const uint32_t u = (lsh >= 32) ? (dl << (lsh & 31)) : __funnelshift_l (dl, dh, lsh);
  const unsigned b = __umulhi (u, mult) >> SHIFT;
  return b < nb ? b : nb - 1;
}

// fold of one more list's count into the running count of a word
template <int MODE>
__device__ __forceinline__ uint32_t kw_fold (uint32_t f, uint32_t c, bool first, int rule)
{
  if (MODE == KWAY_MODE_U_ADD) return f + c;
  if (MODE == KWAY_MODE_I_MIN) return (first || f == 0u || c < f) ? c : f;       // glistcompare.c:668-671 ("!freq ||")
  switch (rule) {
  case RULE_ADD: return f + c;
  case RULE_MAX: return (first || c > f) ? c : f;
  case RULE_MIN: return (first || f == 0u || c < f) ? c : f;
  default:       return f;          // RULE_NUMBER: the override is set after the fold
  }
}

template <int NL, int NC, int S, int MODE, bool COUNT_ONLY>
__global__ void __launch_bounds__ (KwayCfg<NL, NC, S>::NTHREADS, GT4_KWAY_CTAS)
kway_tile_kernel (const KwayArgs args)
{
  using Cfg = KwayCfg<NL, NC, S>;
  constexpr int NWARPS = Cfg::NWARPS;
  constexpr int NB = Cfg::NB;
  constexpr int SLOTS = Cfg::SLOTS;
  constexpr int BSH = Cfg::BUCKET_SHIFT;

  extern __shared__ __align__ (128) unsigned char smem_raw[];
  __shared__ __align__ (8) uint64_t bar_full[S];
  __shared__ __align__ (8) uint64_t bar_comp[S];
  __shared__ __align__ (8) uint64_t bar_agg[S];
  __shared__ __align__ (8) uint64_t bar_base[S];
  __shared__ __align__ (8) uint64_t bar_empty[S];
  __shared__ KwMeta<NL> s_meta[S];
  __shared__ KwMailbox s_mail[S];
  __shared__ int s_wcnt[2][NWARPS];
  __shared__ uint32_t s_wtot[NWARPS][NL / 2];
  __shared__ volatile unsigned int s_n_iter;
  __shared__ unsigned long long s_red[2][NWARPS];

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < S; s++) {
      mbar_init (&bar_full[s], 1);
      mbar_init (&bar_comp[s], NWARPS);
      mbar_init (&bar_agg[s], 1);
      mbar_init (&bar_base[s], 1);
      mbar_init (&bar_empty[s], COUNT_ONLY ? NWARPS : KW_STORE_WARPS);
    }
    s_n_iter = 0xffffffffu;
    fence_mbar_init ();
  }
  __syncthreads ();

  auto stage_keys = [&] (int s) { return reinterpret_cast<uint64_t *> (smem_raw + (size_t) s * Cfg::STAGE_BYTES); };
  auto stage_cnts = [&] (int s) { return reinterpret_cast<uint32_t *> (smem_raw + (size_t) s * Cfg::STAGE_BYTES + (size_t) SLOTS * 8); };
  uint64_t *const sparse_k = reinterpret_cast<uint64_t *> (smem_raw + (size_t) S * Cfg::STAGE_BYTES);
  uint32_t *const sparse_c = reinterpret_cast<uint32_t *> (smem_raw + (size_t) S * Cfg::STAGE_BYTES + (size_t) KWAY_TILE_CAP * 8);
  uint32_t *const hist = reinterpret_cast<uint32_t *> (smem_raw + (size_t) S * Cfg::STAGE_BYTES + Cfg::SPARSE_BYTES);
  for (int i = tid; i < NL * NB; i += Cfg::NTHREADS) hist[i] = 0;
  __syncthreads ();

  // ============================================================================ producer (all 32 lanes)
  if (warp == Cfg::PRODUCER_WARP) {
    const uint64_t n_tiles = args.n_tiles;
    uint64_t tile = 0;
    if (lane == 0) tile = atomicAdd (&args.hdr->ticket, 1u);
    tile = __shfl_sync (0xffffffffu, tile, 0);
    int s = 0;
    uint32_t ph = 0;
    while (true) {
      // this tile's slice of every list (lane j < NL) and its key range (lane 0); loaded before the wait
      uint64_t c_lo = 0, c_hi = 0, k_lo = 0, k_hi = 0, p_lo = 0, p_hi = 0;
      const uint64_t pf_tile = tile + gridDim.x;       // about what this CTA will claim next: its slices are prefetched into L2
      if (tile < n_tiles) {
        if (lane < NL) {
          c_lo = args.cuts[tile * NL + lane];
          c_hi = args.cuts[(tile + 1) * NL + lane];
          if (!COUNT_ONLY && pf_tile < n_tiles) {
            p_lo = args.cuts[pf_tile * NL + lane];
            p_hi = args.cuts[(pf_tile + 1) * NL + lane];
          }
        }
        if (lane == 0) {
          k_lo = args.bounds[tile];
          k_hi = args.bounds[tile + 1];
        }
      }
      mbar_wait_sleep (&bar_empty[s], ph ^ 1u, 50);
#if !GT4_STORE_FENCE
      fence_proxy_async ();      // the stage's last generic-proxy accesses (observed through bar_empty) before the TMA writes
#endif
      if (tile >= n_tiles) {
        // END marker through every stage under the normal protocol (see setop2_stream_kernel)
        if (lane == 0) {
          for (int q = 0; q < (COUNT_ONLY ? 1 : S); q++) {
            if (q > 0) mbar_wait_sleep (&bar_empty[s], ph ^ 1u, 200);
            s_meta[s].tile = KW_TILE_END;
            mbar_arrive (&bar_full[s]);
            if (++s == S) { s = 0; ph ^= 1u; }
          }
        }
        break;
      }
      const uint64_t n_list = (lane < NL) ? args.n[lane] : 0;
      const bool sane = c_hi >= c_lo && c_hi <= n_list && c_hi - c_lo <= (uint64_t) KWAY_TILE_CAP;
      int n = sane ? (int) (c_hi - c_lo) : 0;
      int total = n;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) total += __shfl_xor_sync (0xffffffffu, total, off);
      const bool any_bad = __any_sync (0xffffffffu, !sane);
      if (any_bad || total > KWAY_TILE_CAP) {
        if (lane == 0) args.hdr->overflow = any_bad ? 2u : 3u;     // unsorted input / a tile the sampling bound does not cover
        n = 0;
        total = 0;
      }
      const int pc = (int) (c_lo & 3u);                             // count phase inside its 16-byte group
      const int width = (lane < NL) ? ((pc + n + 1 + 3) & ~3) : 0;  // + 1: the slot behind the slice holds a sentinel key
      int incl = width;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync (0xffffffffu, incl, off);
        if (lane >= off) incl += t;
      }
      const int slot0 = incl - width + pc;                          // slot of the slice's first record
      KwMeta<NL> &m = s_meta[s];
      if (lane < NL) {
        m.idx0[lane] = (uint16_t) slot0;
        m.n[lane] = (uint16_t) n;
      }
      if (lane == 0) {
        m.tile = tile;
        m.n_total = total;
        const uint64_t span1 = k_hi >= k_lo ? k_hi - k_lo : 0;      // span - 1
        const int lsh = span1 ? __clzll ((long long) span1) : 63;
        const uint64_t top = (span1 << lsh) >> 32;                  // in [2^31, 2^32) unless the span is a single key
        m.lo = k_lo;
        m.hi = k_hi;
        m.lsh = lsh;
        m.mult = (uint32_t) (((uint64_t) NB << (32 + BSH)) / (top + 1ull));
      }
      // one block per lane: lanes 2j / 2j + 1 stage the keys / counts of list j
      const int j = lane >> 1, kind = lane & 1;
      const uint64_t b_lo = __shfl_sync (0xffffffffu, c_lo, j);
      const int b_n = __shfl_sync (0xffffffffu, n, j);
      const int b_slot = __shfl_sync (0xffffffffu, slot0, j);
      uint32_t tma_bytes = 0;
      const unsigned char *tma_src = nullptr;
      unsigned char *tma_dst = nullptr;
      if (lane < 2 * NL && b_n > 0) {
        const int es = kind ? 4 : 8;
        const uint64_t al = kind ? 4 : 2;                           // records per 16 bytes
        const unsigned char *arr = kind ? reinterpret_cast<const unsigned char *> (args.counts[j]) : reinterpret_cast<const unsigned char *> (args.words[j]);
        unsigned char *blk = kind ? reinterpret_cast<unsigned char *> (stage_cnts (s)) : reinterpret_cast<unsigned char *> (stage_keys (s));
        const uint64_t first = b_lo & ~(al - 1);
        const uint64_t last = (b_lo + (uint64_t) b_n + al - 1) & ~(al - 1);
        const uint64_t interior = args.n[j] & ~(al - 1);            // the array's 16-byte aligned interior (its base is aligned)
        const uint64_t t_last = last < interior ? last : interior;
        const long long dst0 = (long long) b_slot - (long long) (b_lo - first);
        if (t_last > first) {
          tma_bytes = (uint32_t) ((t_last - first) * es);
          tma_src = arr + first * es;
          tma_dst = blk + dst0 * es;
        }
        // records of the slice beyond the aligned interior (the unaligned tail of the array): plain loads
        for (uint64_t e = (t_last > b_lo ? t_last : b_lo); e < b_lo + (uint64_t) b_n; e++) {
          const long long d = dst0 + (long long) (e - first);
          if (kind) reinterpret_cast<uint32_t *> (blk)[d] = reinterpret_cast<const uint32_t *> (arr)[e];
          else reinterpret_cast<uint64_t *> (blk)[d] = reinterpret_cast<const uint64_t *> (arr)[e];
        }
      }
      uint32_t tx = tma_bytes;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) tx += __shfl_xor_sync (0xffffffffu, tx, off);
      __syncwarp ();                   // meta and hand-copied records of all lanes are ordered before lane 0's arrive (release)
      if (lane == 0) mbar_arrive_expect_tx (&bar_full[s], tx);
      __syncwarp ();
      if (tma_bytes) bulk_g2s (tma_dst, tma_src, tma_bytes, &bar_full[s]);
      // (measured: halves the time the consumers wait for a stage; the count-only pass is faster without it)
      if (!COUNT_ONLY && lane < NL && p_hi > p_lo && p_hi - p_lo <= (uint64_t) KWAY_TILE_CAP && p_hi <= args.n[lane]) {
        prefetch_l2 (args.words[lane] + p_lo, (p_hi - p_lo) * 8);
        prefetch_l2 (args.counts[lane] + p_lo, (p_hi - p_lo) * 4);
      }
      if (lane == 0) tile = atomicAdd (&args.hdr->ticket, 1u);
      tile = __shfl_sync (0xffffffffu, tile, 0);
      if (++s == S) { s = 0; ph ^= 1u; }
    }
    return;
  }

  // ============================================================================ look-back (one warp per stage)
  if (warp >= Cfg::LOOKBACK_WARP0 && warp < Cfg::STORE_WARP0) {
    if (COUNT_ONLY) return;
    const int s = warp - Cfg::LOOKBACK_WARP0;
    uint32_t ph = 0;
    for (uint32_t it = (uint32_t) s;; it += S, ph ^= 1u) {
      mbar_wait_sleep (&bar_agg[s], ph, 200);
      if (it >= s_n_iter) break;
      const uint64_t tile = s_mail[s].tile;
      const uint64_t base = lookback_exclusive<KW_LB_W> (args.desc, tile, (uint64_t) s_mail[s].cnt, lane);
      if (lane == 0) {
        s_mail[s].base = base;
        mbar_arrive (&bar_base[s]);
      }
      __syncwarp ();
    }
    return;
  }

  // ============================================================================ store warps
  if (warp >= Cfg::STORE_WARP0) {
    if (COUNT_ONLY) return;
    const int st_tid = tid - Cfg::STORE_WARP0 * 32;
    constexpr int ST_THREADS = 32 * KW_STORE_WARPS;
    int s = 0;
    uint32_t ph = 0;
    while (true) {
      mbar_wait_sleep (&bar_comp[s], ph, 200);
      if (s_mail[s].tile == KW_TILE_END) break;
      mbar_wait_sleep (&bar_base[s], ph, 100);
      const uint64_t base = s_mail[s].base;
      const int cnt = s_mail[s].cnt;
      const uint64_t *sk = stage_keys (s);
      const uint32_t *sc = stage_cnts (s);
      if (base + (uint64_t) cnt <= args.out_capacity) {
        uint64_t *ow = args.out_words + base;
        uint32_t *oc = args.out_counts + base;
        int x = st_tid;
        for (; x + 7 * ST_THREADS < cnt; x += 8 * ST_THREADS) {
          uint64_t k[8];
#pragma unroll
          for (int r = 0; r < 8; r++) k[r] = sk[x + r * ST_THREADS];
#pragma unroll
          for (int r = 0; r < 8; r++) ow[x + r * ST_THREADS] = k[r];
        }
        for (; x < cnt; x += ST_THREADS) ow[x] = sk[x];
        x = st_tid;
        for (; x + 7 * ST_THREADS < cnt; x += 8 * ST_THREADS) {
          uint32_t c[8];
#pragma unroll
          for (int r = 0; r < 8; r++) c[r] = sc[x + r * ST_THREADS];
#pragma unroll
          for (int r = 0; r < 8; r++) oc[x + r * ST_THREADS] = c[r];
        }
        for (; x < cnt; x += ST_THREADS) oc[x] = sc[x];
      } else if (st_tid == 0) {
        args.hdr->overflow = 1u;
      }
#if GT4_STORE_FENCE
      fence_proxy_async ();
#endif
      __syncwarp ();
      if (lane == 0) mbar_arrive (&bar_empty[s]);
      if (++s == S) { s = 0; ph ^= 1u; }
    }
    return;
  }

  // ============================================================================ consumers
  unsigned long long acc_n = 0, acc_sum = 0;
  int s = 0, n_end = 0;
  uint32_t ph = 0;
  const bool is_isect = args.op != 0;
  const int n_real = args.n_real;
  const int rule = args.rule;
  const uint32_t cutoff = args.final_pass ? args.cutoff : 0u;
  const bool prof = (args.debug & 32) != 0;
  long long t_wait = 0, t_fill = 0, t_loop = 0, t_scan = 0, t_comp = 0, n_prof = 0;
  for (uint32_t it = 0;; it++) {
    const long long c0 = prof ? clock64 () : 0;
    mbar_wait (&bar_full[s], ph);
    const long long c1 = prof ? clock64 () : 0;
    const KwMeta<NL> &m = s_meta[s];
    if (m.tile == KW_TILE_END) {
      if (COUNT_ONLY) break;
      if (tid == 0) {
        if (it < s_n_iter) s_n_iter = it;
        s_mail[s].tile = KW_TILE_END;
        mbar_arrive (&bar_agg[s]);
      }
      __syncwarp ();
      if (lane == 0) mbar_arrive (&bar_comp[s]);
      if (++n_end == S) break;
      if (++s == S) { s = 0; ph ^= 1u; }
      continue;
    }
    uint64_t *sk = stage_keys (s);
    uint32_t *sc = stage_cnts (s);
    const uint64_t lo = m.lo;
    const uint32_t mult = m.mult;
    const int lsh = m.lsh;
    const uint64_t tile_id = m.tile;

    // ---- (1) how many records of every slice fall into every bucket?  (hist is all zero between tiles)
    int base[NL], nj[NL];
    int n_max = 0;
#pragma unroll
    for (int j = 0; j < NL; j++) {
      nj[j] = m.n[j];
      base[j] = m.idx0[j];
      n_max = nj[j] > n_max ? nj[j] : n_max;
    }
    // a key larger than any other behind every slice: a list that runs out inside the merge loop below then shows a head
    // that can never be the smallest (the word after a thread's share of a slice is a natural sentinel: it belongs to a
    // later bucket, so it is larger than every word of this bucket)
    if (tid < NL) sk[m.idx0[tid] + m.n[tid]] = ~0ull;
    // (all loads of a round first, then all atomics: the compiler must keep shared-memory loads behind earlier shared
    // atomics; lanes beyond the end of a slice count into a dummy bucket of their own instead of branching)
    for (int i = tid; i < n_max; i += NC) {
      uint64_t k[NL];
      unsigned b[NL];
#pragma unroll
      for (int j = 0; j < NL; j++) k[j] = sk[base[j] + (i < nj[j] ? i : 0)];
#pragma unroll
      for (int j = 0; j < NL; j++) b[j] = (i < nj[j]) ? j * NB + kw_bucket<BSH> (k[j], lo, lsh, mult, NB) : NL * NB + lane;
#pragma unroll
      for (int j = 0; j < NL; j++) atomicAdd (&hist[b[j]], 1u);
    }
    consumer_sync<NC> ();

    // ---- (2) this thread's bucket: exclusive prefix of the bucket counts over the threads, two lists per 32-bit word
    uint32_t hcnt[NL];
    uint32_t pk[NL / 2];
#pragma unroll
    for (int j = 0; j < NL; j++) {
      hcnt[j] = hist[j * NB + tid];
      hist[j * NB + tid] = 0;
    }
#pragma unroll
    for (int q = 0; q < NL / 2; q++) pk[q] = hcnt[2 * q] | (hcnt[2 * q + 1] << 16);
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
#pragma unroll
      for (int q = 0; q < NL / 2; q++) {
        const uint32_t t = __shfl_up_sync (0xffffffffu, pk[q], off);
        if (lane >= off) pk[q] += t;
      }
    }
    if (lane == 31) {
#pragma unroll
      for (int q = 0; q < NL / 2; q++) s_wtot[warp][q] = pk[q];
    }
    consumer_sync<NC> ();
#pragma unroll
    for (int q = 0; q < NL / 2; q++) {
      // (the packed 16-bit halves cannot carry into each other: a tile holds fewer than 2^16 records)
      const uint32_t w = (lane < warp) ? s_wtot[lane][q] : 0u;          // warp < NWARPS
      pk[q] += __reduce_add_sync (0xffffffffu, w);                      // one REDUX: total of the warps before this one
    }
    const long long c2 = prof ? clock64 () : 0;
    int idx[NL], end[NL];
    uint64_t head[NL];
    int remaining = 0, sparse_off = 0;
#pragma unroll
    for (int j = 0; j < NL; j++) {
      const int incl_j = (int) ((pk[j >> 1] >> (16 * (j & 1))) & 0xffffu);
      const int excl_j = incl_j - (int) hcnt[j];
      idx[j] = base[j] + excl_j;
      end[j] = idx[j] + (int) hcnt[j];
      sparse_off += excl_j;
      remaining += (int) hcnt[j];
      head[j] = sk[idx[j]];
    }
    int pos = sparse_off;
    while (remaining > 0) {
      uint64_t mn = head[0];
#pragma unroll
      for (int j = 1; j < NL; j++) mn = head[j] < mn ? head[j] : mn;
      uint32_t f = 0;
      int n_hit = 0;
      // straight-line code on purpose: every list looks at its count and reloads its head whether it holds the word or not
      // (the loads are cheap, eight divergent branches per word are not)
#pragma unroll
      for (int j = 0; j < NL; j++) {
        const bool hit = head[j] == mn && idx[j] < end[j];
        const uint32_t c = sc[idx[j]];
        f = hit ? kw_fold<MODE> (f, c, n_hit == 0, rule) : f;
        n_hit += hit ? 1 : 0;
        idx[j] += hit ? 1 : 0;
        head[j] = sk[idx[j]];
      }
      if (n_hit == 0) {               // only lists that are not sorted get here: report, never spin
        args.hdr->overflow = 2u;
        break;
      }
      remaining -= n_hit;
      if (MODE == KWAY_MODE_GENERIC && rule == RULE_NUMBER) f = args.count_override;
      bool keep = f >= cutoff;
      if (MODE == KWAY_MODE_I_MIN || (MODE == KWAY_MODE_GENERIC && is_isect)) keep = keep && n_hit == n_real;
      if (keep) {
        if (!COUNT_ONLY) {
          sparse_k[pos] = mn;
          sparse_c[pos] = f;
        }
        pos += 1;
        acc_sum += f;
      }
    }
    const int cnt = pos - sparse_off;
    acc_n += (unsigned) cnt;
    const long long c3 = prof ? clock64 () : 0;
    if (prof) { t_wait += c1 - c0; t_fill += c2 - c1; t_loop += c3 - c2; n_prof += 1; }

    if (COUNT_ONLY) {
      __syncwarp ();
      if (lane == 0) mbar_arrive (&bar_empty[s]);
      if (++s == S) { s = 0; ph ^= 1u; }
      continue;
    }

    // ---- (3) block scan of the survivor counts, tile total to the look-back warp, compaction
    int incl = cnt;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int t = __shfl_up_sync (0xffffffffu, incl, off);
      if (lane >= off) incl += t;
    }
    if (lane == 31) s_wcnt[it & 1][warp] = incl;
    consumer_sync<NC> ();             // also: every consumer is done reading this stage's slices
    static_assert (NWARPS <= 32, "one lane per consumer warp");
    const int wv = (lane < NWARPS) ? s_wcnt[it & 1][lane] : 0;
    const int tile_cnt = (int) __reduce_add_sync (0xffffffffu, (unsigned) wv);
    const int warp_prefix = (int) __reduce_add_sync (0xffffffffu, (unsigned) (lane < warp ? wv : 0));
    if (tid == 0) {
      s_mail[s].tile = tile_id;
      s_mail[s].cnt = tile_cnt;
      mbar_arrive (&bar_agg[s]);
    }
    const long long c4 = prof ? clock64 () : 0;
    const int dst = warp_prefix + incl - cnt;
    for (int q = 0; q < cnt; q += 4) {           // four independent copies in flight, the last round predicated
      uint64_t k4[4];
      uint32_t c4[4];
#pragma unroll
      for (int r = 0; r < 4; r++) {
        const int src = sparse_off + (q + r < cnt ? q + r : q);
        k4[r] = sparse_k[src];
        c4[r] = sparse_c[src];
      }
#pragma unroll
      for (int r = 0; r < 4; r++) {
        if (q + r < cnt) {
          sk[dst + q + r] = k4[r];
          sc[dst + q + r] = c4[r];
        }
      }
    }
    __syncwarp ();
    if (lane == 0) mbar_arrive (&bar_comp[s]);
    if (prof) { const long long c5 = clock64 (); t_scan += c4 - c3; t_comp += c5 - c4; }
    if (++s == S) { s = 0; ph ^= 1u; }
  }
  if (prof && lane == 0) {       // experiments: per-phase cycles of the consumer warps
    atomicAdd (&args.hdr->dbg[0], (unsigned long long) t_wait); atomicAdd (&args.hdr->dbg[1], (unsigned long long) t_fill);
    atomicAdd (&args.hdr->dbg[4], (unsigned long long) t_loop); atomicAdd (&args.hdr->dbg[5], (unsigned long long) t_scan);
    atomicAdd (&args.hdr->dbg[6], (unsigned long long) t_comp); atomicAdd (&args.hdr->dbg[7], (unsigned long long) n_prof);
  }

  acc_n = warp_sum_u64 (acc_n);
  acc_sum = warp_sum_u64 (acc_sum);
  if (lane == 0) {
    s_red[0][warp] = acc_n;
    s_red[1][warp] = acc_sum;
  }
  consumer_sync<NC> ();
  if (tid == 0) {
    unsigned long long n = 0, sum = 0;
#pragma unroll
    for (int w = 0; w < NWARPS; w++) {
      n += s_red[0][w];
      sum += s_red[1][w];
    }
    unsigned long long *slot = args.hdr->totals[0][blockIdx.x & (TOTAL_SLOTS - 1)];
    atomicAdd (slot, n);
    atomicAdd (slot + 1, sum);
  }
}

// ---- pre-passes --------------------------------------------------------------------------------

// samples[sample_off[j] + i] = words_j[(i + 1) * KWAY_SAMPLE - 1]
__global__ void __launch_bounds__ (256)
kway_sample_kernel (const KwayArgs args, uint64_t n_samples, uint64_t *__restrict__ samples)
{
  const uint64_t g = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_samples) return;
  int j = 0;
#pragma unroll
  for (int q = 1; q < KWAY_MAX_LISTS; q++)
    if (q < args.n_lists && g >= args.sample_off[q]) j = q;
  const uint64_t i = g - args.sample_off[j];
  samples[g] = args.words[j][(i + 1) * KWAY_SAMPLE - 1];
}

// cuts[t * NL + j] = number of words of list j below the t-th boundary key (sorted[t * m]); bounds[t] = that key.
// Thread order: neighbouring threads search neighbouring boundaries in the same list (shared search paths).
__global__ void __launch_bounds__ (256)
kway_cuts_kernel (const KwayArgs args, const uint64_t *__restrict__ sorted, uint64_t m, int nl, uint64_t *__restrict__ cuts, uint64_t *__restrict__ bounds)
{
  const uint64_t g = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t per_list = args.n_tiles + 1;
  if (g >= per_list * (uint64_t) nl) return;
  const int j = (int) (g / per_list);
  const uint64_t t = g - (uint64_t) j * per_list;
  const uint64_t n = args.n[j];
  uint64_t cut;
  if (t == 0) cut = 0;
  else if (t == args.n_tiles) cut = n;
  else {
    const uint64_t x = sorted[t * m];
    const uint64_t *w = args.words[j];
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
      const uint64_t mid = lo + ((hi - lo) >> 1);
      if (w[mid] < x) lo = mid + 1;
      else hi = mid;
    }
    cut = lo;
    if (j == 0) bounds[t] = x;
  }
  cuts[t * (uint64_t) nl + j] = cut;
  if (j == 0 && (t == 0 || t == args.n_tiles)) {
    // the outer bounds: smallest first word / largest last word over the non-empty lists
    uint64_t v = t == 0 ? ~0ull : 0ull;
    for (int q = 0; q < args.n_lists; q++) {
      if (!args.n[q]) continue;
      const uint64_t e = t == 0 ? args.words[q][0] : args.words[q][args.n[q] - 1];
      v = t == 0 ? (e < v ? e : v) : (e > v ? e : v);
    }
    bounds[t] = v;
  }
}

template <int NL, int NC, int S, int MODE, bool CO>
cudaError_t launch_kway_one (const KwayArgs &args, int sm_count, cudaStream_t st)
{
  using Cfg = KwayCfg<NL, NC, S>;
  auto kernel = kway_tile_kernel<NL, NC, S, MODE, CO>;
  static bool configured = false;      // benign race: idempotent
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute (kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  uint64_t grid = (uint64_t) sm_count * GT4_KWAY_CTAS;
  if (grid > args.n_tiles) grid = args.n_tiles;
  kernel<<<(unsigned) grid, Cfg::NTHREADS, Cfg::SMEM_BYTES, st>>> (args);
  return cudaGetLastError ();
}

template <int NL, int NC, int S>
cudaError_t launch_kway_mode (const KwayArgs &args, int mode, bool count_only, int sm_count, cudaStream_t st)
{
#define GT4_KW(MODE)                                                                                      \
  return count_only ? launch_kway_one<NL, NC, S, MODE, true> (args, sm_count, st) : launch_kway_one<NL, NC, S, MODE, false> (args, sm_count, st)
  switch (mode) {
  case KWAY_MODE_U_ADD: GT4_KW (KWAY_MODE_U_ADD);
  case KWAY_MODE_I_MIN: GT4_KW (KWAY_MODE_I_MIN);
  default:              GT4_KW (KWAY_MODE_GENERIC);
  }
#undef GT4_KW
}

}  // namespace

int kway_select_mode (int op, int rule)
{
  if (op == 0 && rule == RULE_ADD) return KWAY_MODE_U_ADD;
  if (op != 0 && rule == RULE_MIN) return KWAY_MODE_I_MIN;
  return KWAY_MODE_GENERIC;
}

cudaError_t launch_kway_samples (const KwayArgs &args, uint64_t n_samples, uint64_t *samples, cudaStream_t st)
{
  if (n_samples == 0) return cudaSuccess;
  kway_sample_kernel<<<(unsigned) ((n_samples + 255) / 256), 256, 0, st>>> (args, n_samples, samples);
  return cudaGetLastError ();
}

cudaError_t launch_kway_cuts (const KwayArgs &args, const uint64_t *sorted_samples, uint64_t every, int nl, uint64_t *cuts, uint64_t *bounds, cudaStream_t st)
{
  const uint64_t n = (args.n_tiles + 1) * (uint64_t) nl;
  kway_cuts_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, st>>> (args, sorted_samples, every, nl, cuts, bounds);
  return cudaGetLastError ();
}

cudaError_t launch_kway_tiles (const KwayArgs &args, int nl, bool count_only, int sm_count, cudaStream_t st)
{
  if (args.n_tiles == 0) return cudaSuccess;
  const int mode = kway_select_mode (args.op, args.rule);
  static int consumers = 0;      // benign race: idempotent
  if (consumers == 0) {
    const char *env = getenv ("GT4GPU_KWAY_CONSUMERS");          // experiments: 256 / 512 consumer threads (8-list variant)
    const int v = env ? atoi (env) : KWAY_CONSUMERS;
    consumers = (v == 512) ? v : KWAY_CONSUMERS;
  }
  if (nl <= 4) return launch_kway_mode<4, KWAY_CONSUMERS, KWAY_STAGES> (args, mode, count_only, sm_count, st);
  if (nl <= 8) {
    if (consumers == 512) return launch_kway_mode<8, 512, KWAY_STAGES> (args, mode, count_only, sm_count, st);
    return launch_kway_mode<8, KWAY_CONSUMERS, KWAY_STAGES> (args, mode, count_only, sm_count, st);
  }
  return cudaErrorInvalidValue;
}

}  // namespace gt4gpu
