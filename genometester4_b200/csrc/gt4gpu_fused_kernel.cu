// gt4gpu_fused_kernel.cu -- the multi-output merge: ONE read of the two lists, up to four outputs.
//
// compare_wordmaps emits union / intersection / diff1 / diff2 from a single merge loop
// (/root/reference/src/glistcompare.c:843-905).  setop2_fused_kernel does the same on the pipeline of
// setop2_stream_kernel (producer warp + TMA staging, splitter warp, consumer warps, one look-back warp per stage,
// store warps); what differs:
//
//   * every merged slot is evaluated for all requested outputs (eval_stream, gt4gpu_core.cuh);
//   * without -du a word lands in at most ONE of intersection / diff1 / diff2 (their cut-off tests exclude each other),
//     so those three are compacted into one "rest" region, one sub-range each; the union -- when it is requested with
//     them -- is compacted in place at the front of the stage and the rest goes to an auxiliary buffer of the stage
//     (AUX = true: tiles of 512 x 5 slots, three stages); without the union the rest is compacted in place (AUX = false:
//     512 x 9, four stages);
//   * the four survivor counts of a tile travel in ONE look-back chain: a descriptor is 32 bytes -- word 0 carries the
//     status and the four tile counts (15 bits each), words 1..3 the inclusive prefixes -- so resolving four output
//     offsets costs the L2 round trips of one.
//
// Algorithmic traffic: 12 B per input record ONCE + 12 B per output record of every output.
#include <cuda_runtime.h>
#include <stdint.h>

#include "gt4gpu_device.cuh"
#include "gt4gpu_internal.h"

#ifndef GT4_FUSED_VT_AUX
#define GT4_FUSED_VT_AUX 5          // merged slots per consumer thread / stages when the union is one of the outputs
#endif
#ifndef GT4_FUSED_S_AUX
#define GT4_FUSED_S_AUX 3
#endif
#ifndef GT4_FUSED_VT_REST
#define GT4_FUSED_VT_REST 7         // ... and when it is not (everything is compacted in place)
#endif
#ifndef GT4_FUSED_S_REST
#define GT4_FUSED_S_REST 5
#endif
#ifndef GT4_FUSED_STORE_WARPS
#define GT4_FUSED_STORE_WARPS 4     // up to four outputs leave a tile
#endif
#ifndef GT4_FUSED_STORE_FENCE
#define GT4_FUSED_STORE_FENCE 0     // proxy fence before the TMA refills a stage: 1 = in the store warps (waits for their global stores), 0 = in the producer
#endif

namespace gt4gpu {

namespace {

constexpr int STORE_WARPS = GT4_FUSED_STORE_WARPS;
constexpr uint64_t TILE_END = ~0ull;
constexpr int FLB_W = 4;

using namespace dev;

template <int NC, int VT, int S, bool AUX>
struct FusedCfg {
  static constexpr int CONSUMERS = NC;
  static constexpr int PRODUCER_WARP = NC / 32;
  static constexpr int SPLITTER_WARP = NC / 32 + 1;
  static constexpr int LOOKBACK_WARP0 = NC / 32 + 2;
  static constexpr int STORE_WARP0 = LOOKBACK_WARP0 + S;
  static constexpr int NTHREADS = NC + 64 + 32 * S + 32 * STORE_WARPS;
  static constexpr int GROUP = NC / 32;
  static constexpr int NSPLIT = NC / GROUP + 1;
  static constexpr int TILE = CONSUMERS * VT;
  static constexpr int KSLOTS = (TILE + VT + 16 + 1) & ~1;
  static constexpr int CSLOTS = (TILE + VT + 28 + 3) & ~3;
  static constexpr int ASLOTS = AUX ? ((TILE + 4) & ~3) : 0;          // the rest region of a stage when the union sits in place
  static constexpr size_t IN_BYTES = (size_t) KSLOTS * 8 + (size_t) CSLOTS * 4;
  static constexpr size_t STAGE_BYTES = IN_BYTES + (size_t) ASLOTS * 12;
  static constexpr size_t SMEM_BYTES = S * STAGE_BYTES;
};

struct StageMeta {
  uint64_t tile;
  int na, nb;
  int ka, kb;
  int ca, cb;
  int flags;
};

struct Mailbox4 {
  uint64_t tile;
  uint64_t base[4];
  int cnt[4];
};

// ---- look-back over 32-byte descriptors: word 0 = status (2 bits) | four tile counts (15 bits each, PARTIAL and
// INCLUSIVE alike), words 1..3 hold the four inclusive prefixes.
__device__ __forceinline__ void unpack4 (uint64_t w, uint32_t c[4])
{
#pragma unroll
  for (int q = 0; q < 4; q++) c[q] = (uint32_t) (w >> (15 * q)) & 0x7fffu;
}

__device__ __forceinline__ uint64_t pack4 (const int c[4])
{
  return (uint64_t) c[0] | ((uint64_t) c[1] << 15) | ((uint64_t) c[2] << 30) | ((uint64_t) c[3] << 45);
}

// The inclusive prefixes (47 bits each) travel in words 1..3, 63 payload bits per word; bit 63 of every word says "written"
// (the descriptors start out zero), so no fence is needed between them and the status word: a reader that finds the
// status INCLUSIVE simply re-reads a prefix word until its bit 63 is set.
constexpr uint64_t W_VALID = 1ull << 63;

__device__ __forceinline__ void st_inclusive4 (uint64_t *d, const uint64_t incl[4], const int cnt[4])
{
  const uint64_t m47 = (1ull << 47) - 1;
  const uint64_t p0 = incl[0] & m47, p1 = incl[1] & m47, p2 = incl[2] & m47, p3 = incl[3] & m47;
  st_relaxed (d + 1, W_VALID | p0 | ((p1 & 0xffffull) << 47));
  st_relaxed (d + 2, W_VALID | (p1 >> 16) | ((p2 & 0xffffffffull) << 31));
  st_relaxed (d + 3, W_VALID | (p2 >> 32) | (p3 << 15));
  st_relaxed (d, DESC_INCLUSIVE | pack4 (cnt));
}

__device__ __forceinline__ void ld_inclusive4 (const uint64_t *d, uint64_t incl[4])
{
  uint64_t w1, w2, w3;
  do { w1 = ld_relaxed (d + 1); } while (!(w1 & W_VALID));
  do { w2 = ld_relaxed (d + 2); } while (!(w2 & W_VALID));
  do { w3 = ld_relaxed (d + 3); } while (!(w3 & W_VALID));
  const uint64_t m47 = (1ull << 47) - 1;
  incl[0] = w1 & m47;
  incl[1] = ((w1 >> 47) & 0xffffull) | ((w2 & 0x7fffffffull) << 16);
  incl[2] = ((w2 >> 31) & 0xffffffffull) | ((w3 & 0x7fffull) << 32);
  incl[3] = (w3 >> 15) & m47;
}

// All 32 lanes; publishes the tile's four counts and returns their exclusive prefixes in excl[4] (same value in all lanes).
__device__ __forceinline__ void lookback_exclusive4 (uint64_t *desc, uint64_t tile, const int cnt[4], int lane, uint64_t excl[4])
{
#pragma unroll
  for (int q = 0; q < 4; q++) excl[q] = 0;
  if (tile == 0) {
    if (lane == 0) {
      uint64_t incl[4];
#pragma unroll
      for (int q = 0; q < 4; q++) incl[q] = (uint64_t) cnt[q];
      st_inclusive4 (desc, incl, cnt);
    }
    return;
  }
  if (lane == 0) st_relaxed (desc + 4 * tile, DESC_PARTIAL | pack4 (cnt));
  uint64_t lane_sum[4] = {0, 0, 0, 0};
  int64_t pred = (int64_t) tile - 1 - lane;
  int64_t incl_tile = -1;                    // the tile whose inclusive prefixes close the sum (known to the lane that saw it)
  bool done = false;
  while (!done) {
    uint64_t d[FLB_W];
#pragma unroll
    for (int k = 0; k < FLB_W; k++) d[k] = (pred - 32 * k >= 0) ? ld_relaxed (desc + 4 * (pred - 32 * k)) : (DESC_INCLUSIVE | (1ull << 61));
#pragma unroll
    for (int k = 0; k < FLB_W; k++) {
      if (done) break;
      while (true) {
        const uint32_t st = (uint32_t) (d[k] >> 62);
        const uint32_t m_wait = __ballot_sync (0xffffffffu, st == 0);
        const uint32_t m_incl = __ballot_sync (0xffffffffu, st == 2);
        const uint32_t m_stop = m_wait | m_incl;
        uint32_t c[4];
        unpack4 (d[k], c);
        if (m_stop == 0) {
#pragma unroll
          for (int q = 0; q < 4; q++) lane_sum[q] += c[q];
          break;
        }
        const int first = __ffs (m_stop) - 1;
        if ((m_wait >> first) & 1u) {
          d[k] = (pred - 32 * k >= 0) ? ld_relaxed (desc + 4 * (pred - 32 * k)) : (DESC_INCLUSIVE | (1ull << 61));
          continue;
        }
        if (lane < first) {
#pragma unroll
          for (int q = 0; q < 4; q++) lane_sum[q] += c[q];
        }
        if (lane == first && !((d[k] >> 61) & 1ull)) incl_tile = pred - 32 * k;      // (bit 61: "before the first tile": prefix 0)
        done = true;
        break;
      }
    }
    pred -= 32 * FLB_W;
  }
  if (incl_tile >= 0) {
    uint64_t incl[4];
    ld_inclusive4 (desc + 4 * incl_tile, incl);
#pragma unroll
    for (int q = 0; q < 4; q++) lane_sum[q] += incl[q];
  }
#pragma unroll
  for (int q = 0; q < 4; q++) excl[q] = warp_sum_u64 (lane_sum[q]);
  if (lane == 0) {
    uint64_t incl[4];
#pragma unroll
    for (int q = 0; q < 4; q++) incl[q] = excl[q] + (uint64_t) cnt[q];
    st_inclusive4 (desc + 4 * tile, incl, cnt);
  }
}

// ---- the kernel ------------------------------------------------------------------------------
template <int NC, int VT, int S, bool AUX, bool DEFAULT_RULES>
__global__ void __launch_bounds__ (FusedCfg<NC, VT, S, AUX>::NTHREADS, 1)
setop2_fused_kernel (const TileArgs args)
{
  using Cfg = FusedCfg<NC, VT, S, AUX>;
  constexpr int TILE = Cfg::TILE;
  constexpr int STAGES = S;
  constexpr int NWARPS = NC / 32;
  constexpr int PRODUCER_WARP = Cfg::PRODUCER_WARP;
  constexpr int SPLITTER_WARP = Cfg::SPLITTER_WARP;
  constexpr int LOOKBACK_WARP0 = Cfg::LOOKBACK_WARP0;
  constexpr int GROUP = Cfg::GROUP;
  constexpr int NSPLIT = Cfg::NSPLIT;
  constexpr int STORE_WARP0 = Cfg::STORE_WARP0;
  constexpr bool COUNT_ONLY = false;

  extern __shared__ __align__ (128) unsigned char smem_raw[];
  __shared__ __align__ (8) uint64_t bar_full[STAGES];
  __shared__ __align__ (8) uint64_t bar_split[STAGES];
  __shared__ __align__ (8) uint64_t bar_comp[STAGES];
  __shared__ __align__ (8) uint64_t bar_agg[STAGES];
  __shared__ __align__ (8) uint64_t bar_base[STAGES];
  __shared__ __align__ (8) uint64_t bar_empty[STAGES];
  __shared__ StageMeta s_meta[STAGES];
  __shared__ Mailbox4 s_mail[STAGES];
  __shared__ int s_split[STAGES][NSPLIT];
  __shared__ unsigned long long s_wcnt[2][NWARPS];
  __shared__ volatile unsigned int s_n_iter;
  __shared__ unsigned long long s_red[8][NWARPS];

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; s++) {
      mbar_init (&bar_full[s], 1);
      mbar_init (&bar_split[s], 1);
      mbar_init (&bar_comp[s], NWARPS);
      mbar_init (&bar_agg[s], 1);
      mbar_init (&bar_base[s], 1);
      mbar_init (&bar_empty[s], STORE_WARPS);
    }
    s_n_iter = 0xffffffffu;
    fence_mbar_init ();
  }
  __syncthreads ();

  auto stage_keys = [&] (int s) { return reinterpret_cast<uint64_t *> (smem_raw + (size_t) s * Cfg::STAGE_BYTES); };
  auto stage_cnts = [&] (int s) { return reinterpret_cast<uint32_t *> (smem_raw + (size_t) s * Cfg::STAGE_BYTES + (size_t) Cfg::KSLOTS * 8); };
  // the rest region: the auxiliary buffer of the stage, or (no union requested) the front of the stage itself
  auto rest_keys = [&] (int s) { return AUX ? reinterpret_cast<uint64_t *> (smem_raw + (size_t) s * Cfg::STAGE_BYTES + Cfg::IN_BYTES) : stage_keys (s); };
  auto rest_cnts = [&] (int s) { return AUX ? reinterpret_cast<uint32_t *> (smem_raw + (size_t) s * Cfg::STAGE_BYTES + Cfg::IN_BYTES + (size_t) Cfg::ASLOTS * 8) : stage_cnts (s); };

  // ============================================================================ producer
  if (warp == PRODUCER_WARP) {
    if (lane != 0) return;
    const uint64_t total = args.na + args.nb;
    const uint64_t n_tiles = args.n_tiles;
    // 16-byte aligned interior [lo, hi) of the four input arrays
    uintptr_t lim[8];
    {
      const uintptr_t lo[4] = {(uintptr_t) args.a_words, (uintptr_t) args.b_words, (uintptr_t) args.a_counts, (uintptr_t) args.b_counts};
      const uintptr_t hi[4] = {(uintptr_t) (args.a_words + args.na), (uintptr_t) (args.b_words + args.nb),
                               (uintptr_t) (args.a_counts + args.na), (uintptr_t) (args.b_counts + args.nb)};
#pragma unroll
      for (int q = 0; q < 4; q++) {
        lim[2 * q] = (lo[q] + 15) & ~(uintptr_t) 15;
        lim[2 * q + 1] = hi[q] & ~(uintptr_t) 15;
      }
    }
    uint64_t nxt = atomicAdd (&args.hdr->ticket, 1u);
    uint64_t nxt_lo = 0, nxt_hi = 0;
    if (nxt < n_tiles) { nxt_lo = args.part[nxt]; nxt_hi = args.part[nxt + 1]; }
    // L2 prefetch of the tiles the grid will claim about two rounds from now (co-ranks loaded one iteration early)
    const uint64_t pf_dist = gridDim.x;
    uint64_t pf_tile = nxt + pf_dist, pf_lo = 0, pf_hi = 0;
    if (pf_tile < n_tiles) { pf_lo = args.part[pf_tile]; pf_hi = args.part[pf_tile + 1]; }
    int s = 0;
    uint32_t ph = 0;
    while (true) {
      const uint64_t tile = nxt, a_lo = nxt_lo, a_hi = nxt_hi;
      const uint64_t cur_pf = pf_tile, cur_pf_lo = pf_lo, cur_pf_hi = pf_hi;
      if (tile < n_tiles) {     // claim the following tile now: its latency hides behind the wait below
        nxt = atomicAdd (&args.hdr->ticket, 1u);
        if (nxt < n_tiles) { nxt_lo = args.part[nxt]; nxt_hi = args.part[nxt + 1]; }
        pf_tile = nxt + pf_dist;
        if (pf_tile < n_tiles) { pf_lo = args.part[pf_tile]; pf_hi = args.part[pf_tile + 1]; }
      }
      if (!COUNT_ONLY && (args.debug & 4) == 0 && cur_pf < n_tiles && cur_pf_hi >= cur_pf_lo && cur_pf_hi - cur_pf_lo <= (uint64_t) TILE) {   // (the count-only pass is faster without it)
        const uint64_t pd_lo = cur_pf * TILE;
        const uint64_t pd_hi = (pd_lo + TILE < total) ? pd_lo + TILE : total;
        const uint64_t pb_lo = pd_lo - cur_pf_lo, pb_hi = pd_hi - cur_pf_hi;
        prefetch_l2 (args.a_words + cur_pf_lo, (cur_pf_hi - cur_pf_lo) * 8);
        prefetch_l2 (args.b_words + pb_lo, (pb_hi - pb_lo) * 8);
        prefetch_l2 (args.a_counts + cur_pf_lo, (cur_pf_hi - cur_pf_lo) * 4);
        prefetch_l2 (args.b_counts + pb_lo, (pb_hi - pb_lo) * 4);
      }
      mbar_wait_relaxed (&bar_empty[s], ph ^ 1u);
#if !GT4_FUSED_STORE_FENCE
      fence_proxy_async ();      // the stage's last generic-proxy accesses (observed through bar_empty) before the TMA writes
#endif
      if (tile >= n_tiles) {
        // out of work: send an END marker through EVERY stage, in order and under the normal stage protocol
        // (each look-back warp owns one stage and must see its own marker; a barrier may never be advanced
        // twice before its waiter has looked)
        for (int q = 0; q < (COUNT_ONLY ? 1 : STAGES); q++) {
          if (q > 0) mbar_wait_relaxed (&bar_empty[s], ph ^ 1u);
          s_meta[s].tile = TILE_END;
          mbar_arrive (&bar_full[s]);
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
        break;
      }
      const uint64_t d_lo = tile * TILE;
      const uint64_t d_hi = (d_lo + TILE < total) ? d_lo + TILE : total;
      // co-ranks of strictly ascending lists are monotone with 0 <= a_hi - a_lo <= d_hi - d_lo; anything else means the
      // inputs are not sorted: stage an empty tile and report it instead of copying out of bounds
      const bool sane = a_hi >= a_lo && a_hi - a_lo <= d_hi - d_lo;
      if (!sane) args.hdr->overflow = 2u;
      const uint64_t a_hi_ok = sane ? a_hi : a_lo;
      const uint64_t b_lo = d_lo - a_lo, b_hi = sane ? d_hi - a_hi : b_lo;
      const int na = (int) (a_hi_ok - a_lo), nb = (int) (b_hi - b_lo);
      const int halo = a_lo > 0 ? 1 : 0, peek = b_hi < args.nb ? 1 : 0;
      const int peek_a = (sane && a_hi < args.na) ? 1 : 0;       // the element of A after the tile: a natural sentinel (merge_slots_interior)
      uint64_t *sk = stage_keys (s);
      uint32_t *sc = stage_cnts (s);

      // Byte ranges to stage.  A block is laid out from the 16-byte boundary below its first byte to the one above
      // its last byte (TMA bulk copies need 16-byte aligned addresses and sizes).  Nothing outside the arrays is
      // ever read: see the edge case below.
      const uintptr_t ak0 = (uintptr_t) (args.a_words + a_lo - halo), ak1 = (uintptr_t) (args.a_words + a_hi_ok + peek_a);
      const uintptr_t bk0 = (uintptr_t) (args.b_words + b_lo), bk1 = (uintptr_t) (args.b_words + b_hi + peek);
      const uintptr_t ac0 = (uintptr_t) (args.a_counts + a_lo - halo), ac1 = (uintptr_t) (args.a_counts + a_hi_ok + peek_a);
      const uintptr_t bc0 = (uintptr_t) (args.b_counts + b_lo), bc1 = (uintptr_t) (args.b_counts + b_hi + peek);
      const uintptr_t ak0a = ak0 & ~(uintptr_t) 15, bk0a = bk0 & ~(uintptr_t) 15, ac0a = ac0 & ~(uintptr_t) 15, bc0a = bc0 & ~(uintptr_t) 15;
      const uint32_t ak_bytes = (ak1 > ak0) ? (uint32_t) (((ak1 + 15) & ~(uintptr_t) 15) - ak0a) : 0u;
      const uint32_t bk_bytes = (bk1 > bk0) ? (uint32_t) (((bk1 + 15) & ~(uintptr_t) 15) - bk0a) : 0u;
      const uint32_t ac_bytes = (ac1 > ac0) ? (uint32_t) (((ac1 + 15) & ~(uintptr_t) 15) - ac0a) : 0u;
      const uint32_t bc_bytes = (bc1 > bc0) ? (uint32_t) (((bc1 + 15) & ~(uintptr_t) 15) - bc0a) : 0u;

      StageMeta m;
      m.tile = tile;
      m.na = na;
      m.nb = nb;
      m.ka = (int) ((ak0 - ak0a) >> 3) + halo;                       // A keys start at element 0 of the key array
      m.kb = (int) (ak_bytes >> 3) + (int) ((bk0 - bk0a) >> 3);      // B keys follow the A block
      m.ca = (int) ((ac0 - ac0a) >> 2) + halo;
      m.cb = (int) (ac_bytes >> 2) + (int) ((bc0 - bc0a) >> 2);
      m.flags = halo | (peek << 1) | ((halo && peek && peek_a && d_hi - d_lo == (uint64_t) TILE) ? 4 : 0);   // bit 2: interior tile
      s_meta[s] = m;

      unsigned char *skb = reinterpret_cast<unsigned char *> (sk), *scb = reinterpret_cast<unsigned char *> (sc);
      const bool interior = ak0a >= lim[0] && ak0a + ak_bytes <= lim[1] && bk0a >= lim[2] && bk0a + bk_bytes <= lim[3] &&
                            ac0a >= lim[4] && ac0a + ac_bytes <= lim[5] && bc0a >= lim[6] && bc0a + bc_bytes <= lim[7];
      if (interior) {     // the common case: four aligned bulk copies
        mbar_arrive_expect_tx (&bar_full[s], ak_bytes + bk_bytes + ac_bytes + bc_bytes);
        if (ak_bytes) bulk_g2s (skb, (const void *) ak0a, ak_bytes, &bar_full[s]);
        if (bk_bytes) bulk_g2s (skb + ak_bytes, (const void *) bk0a, bk_bytes, &bar_full[s]);
        if (ac_bytes) bulk_g2s (scb, (const void *) ac0a, ac_bytes, &bar_full[s]);
        if (bc_bytes) bulk_g2s (scb + ac_bytes, (const void *) bc0a, bc_bytes, &bar_full[s]);
      } else {
        // a slice touches an unaligned head or tail of its array: clip the TMA part to the aligned interior and copy
        // the rest with plain loads
        struct Piece { uintptr_t src; uint32_t bytes; unsigned char *dst; };
        Piece tma[4];
        uint32_t tx = 0;
        auto plan_block = [&] (int q, uintptr_t x0, uintptr_t x1, uintptr_t x0a, uintptr_t in_lo, uintptr_t in_hi, unsigned char *block) {
          tma[q].bytes = 0;
          if (x1 <= x0) return;
          uintptr_t t_lo = x0a, t_hi = (x1 + 15) & ~(uintptr_t) 15;
          if (t_lo < in_lo) t_lo = in_lo;
          if (t_hi > in_hi) t_hi = in_hi;
          if (t_hi > t_lo) {
            tma[q].src = t_lo;
            tma[q].bytes = (uint32_t) (t_hi - t_lo);
            tma[q].dst = block + (t_lo - x0a);
            tx += tma[q].bytes;
          } else {
            t_lo = t_hi = x0;      // nothing for the TMA: copy everything by hand
          }
          for (uintptr_t p = x0; p < x1 && p < t_lo; p += 4)                       // unaligned head of the array
            *reinterpret_cast<uint32_t *> (block + (p - x0a)) = *reinterpret_cast<const uint32_t *> (p);
          for (uintptr_t p = (t_hi > x0 ? t_hi : x0); p < x1; p += 4)              // unaligned tail of the array
            *reinterpret_cast<uint32_t *> (block + (p - x0a)) = *reinterpret_cast<const uint32_t *> (p);
        };
        plan_block (0, ak0, ak1, ak0a, lim[0], lim[1], skb);
        plan_block (1, bk0, bk1, bk0a, lim[2], lim[3], skb + ak_bytes);
        plan_block (2, ac0, ac1, ac0a, lim[4], lim[5], scb);
        plan_block (3, bc0, bc1, bc0a, lim[6], lim[7], scb + ac_bytes);
        mbar_arrive_expect_tx (&bar_full[s], tx);      // (release: the plain copies above are visible to whoever sees the phase complete)
#pragma unroll
        for (int q = 0; q < 4; q++)
          if (tma[q].bytes) bulk_g2s (tma[q].dst, (const void *) tma[q].src, tma[q].bytes, &bar_full[s]);
      }
      if (++s == STAGES) { s = 0; ph ^= 1u; }
    }
    return;
  }

  // ============================================================================ splitter
  // Works one or two tiles ahead of the consumers on the stages the TMA has already filled: co-rank of every
  // GROUP-th consumer thread's diagonal over the whole tile (full-depth searches, 32 at a time), so that the
  // consumers' own searches only span GROUP * VT slots.
  if (warp == SPLITTER_WARP) {
    int s = 0, n_end = 0;
    uint32_t ph = 0;
    while (true) {
      mbar_wait (&bar_full[s], ph);
      const StageMeta m = s_meta[s];
      if (m.tile != TILE_END) {
        const uint64_t *sk = stage_keys (s);
        const uint64_t *ka = sk + m.ka, *kb = sk + m.kb;
        const int n_tile = m.na + m.nb;
#pragma unroll
        for (int r = 0; r < (NSPLIT - 1 + 31) / 32; r++) {
          const int g = lane + 32 * r;
          if (g < NSPLIT - 1) {
            const int dg = g * GROUP * VT;
            s_split[s][g] = (dg < n_tile) ? merge_path<int> (ka, m.na, kb, m.nb, dg) : m.na;
          }
        }
        if (lane == 0) s_split[s][NSPLIT - 1] = m.na;     // the diagonal at the end of the tile
      }
      __syncwarp ();
      if (lane == 0) mbar_arrive (&bar_split[s]);
      if (m.tile == TILE_END && ++n_end == (COUNT_ONLY ? 1 : STAGES)) break;
      if (++s == STAGES) { s = 0; ph ^= 1u; }
    }
    return;
  }


  // ============================================================================ look-back (one warp per stage)
  if (warp >= LOOKBACK_WARP0 && warp < STORE_WARP0) {
    const int s = warp - LOOKBACK_WARP0;
    uint32_t ph = 0;
    for (uint32_t it = (uint32_t) s;; it += STAGES, ph ^= 1u) {
      mbar_wait_relaxed (&bar_agg[s], ph);
      if (it >= s_n_iter) break;
      int cnt[4];
#pragma unroll
      for (int q = 0; q < 4; q++) cnt[q] = s_mail[s].cnt[q];
      uint64_t excl[4];
      lookback_exclusive4 (args.desc, s_mail[s].tile, cnt, lane, excl);
      if (lane == 0) {
#pragma unroll
        for (int q = 0; q < 4; q++) s_mail[s].base[q] = excl[q];
        mbar_arrive (&bar_base[s]);
      }
      __syncwarp ();
    }
    return;
  }

  // ============================================================================ store warps
  if (warp >= STORE_WARP0) {
    const int st_tid = tid - STORE_WARP0 * 32;
    constexpr int ST_THREADS = 32 * STORE_WARPS;
    int s = 0;
    uint32_t ph = 0;
    while (true) {
      mbar_wait_relaxed (&bar_comp[s], ph);
      if (s_mail[s].tile == TILE_END) break;
      mbar_wait_relaxed (&bar_base[s], ph);
      int off_rest = 0;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int cnt = s_mail[s].cnt[q];
        if (cnt == 0) continue;
        const uint64_t base = s_mail[s].base[q];
        const bool in_place = AUX && q == 0;
        const uint64_t *sk = in_place ? stage_keys (s) : rest_keys (s) + off_rest;
        const uint32_t *sc = in_place ? stage_cnts (s) : rest_cnts (s) + off_rest;
        if (!in_place) off_rest += cnt;
        if (base + (uint64_t) cnt <= args.out_capacity[q]) {
          uint64_t *ow = args.out_words[q] + base;
          uint32_t *oc = args.out_counts[q] + base;
          int x = st_tid;
          for (; x + 3 * ST_THREADS < cnt; x += 4 * ST_THREADS) {
            uint64_t k[4];
#pragma unroll
            for (int r = 0; r < 4; r++) k[r] = sk[x + r * ST_THREADS];
#pragma unroll
            for (int r = 0; r < 4; r++) ow[x + r * ST_THREADS] = k[r];
          }
          for (; x < cnt; x += ST_THREADS) ow[x] = sk[x];
          x = st_tid;
          for (; x + 3 * ST_THREADS < cnt; x += 4 * ST_THREADS) {
            uint32_t c[4];
#pragma unroll
            for (int r = 0; r < 4; r++) c[r] = sc[x + r * ST_THREADS];
#pragma unroll
            for (int r = 0; r < 4; r++) oc[x + r * ST_THREADS] = c[r];
          }
          for (; x < cnt; x += ST_THREADS) oc[x] = sc[x];
        } else if (st_tid == 0) {
          args.hdr->overflow = 1u;
        }
      }
#if GT4_FUSED_STORE_FENCE
      fence_proxy_async ();
#endif
      __syncwarp ();
      if (lane == 0) mbar_arrive (&bar_empty[s]);
      if (++s == STAGES) { s = 0; ph ^= 1u; }
    }
    return;
  }

  // ============================================================================ consumers
  unsigned long long acc_n[4] = {0, 0, 0, 0}, acc_sum[4] = {0, 0, 0, 0};
  const uint32_t ops = args.p.ops;
  int s = 0, n_end = 0;
  uint32_t ph = 0;
  for (uint32_t it = 0;; it++) {
    mbar_wait (&bar_split[s], ph);
    mbar_wait (&bar_full[s], ph);
    const StageMeta m = s_meta[s];
    if (m.tile == TILE_END) {
      if (tid == 0) {
        if (it < s_n_iter) s_n_iter = it;
        s_mail[s].tile = TILE_END;
        mbar_arrive (&bar_agg[s]);
      }
      __syncwarp ();
      if (lane == 0) mbar_arrive (&bar_comp[s]);
      if (++n_end == STAGES) break;
      if (++s == STAGES) { s = 0; ph ^= 1u; }
      continue;
    }
    uint64_t *sk = stage_keys (s);
    uint32_t *sc = stage_cnts (s);
    const uint64_t *ka = sk + m.ka;
    const uint32_t *ca = sc + m.ca;
    const uint64_t *kb = sk + m.kb;
    const uint32_t *cb = sc + m.cb;
    const int n_tile = m.na + m.nb;
    const int d0 = (tid * VT < n_tile) ? tid * VT : n_tile;
    const int i0 = merge_path_window<int> (ka, m.na, kb, m.nb, d0, s_split[s][tid / GROUP], s_split[s][tid / GROUP + 1]);

    uint64_t o_key[VT];
    uint32_t o_fu[VT], o_fx[VT];
    uint32_t mask_u = 0, kinds = 0;              // kinds: 2 bits per slot, 0 = no rest output, 1..3 = intersection / diff1 / diff2
    auto sink = [&] (int sl, uint64_t key, uint32_t c1, uint32_t c2, bool in_a, bool in_b, bool live) {
        uint32_t f = 0, fx = 0, kind = 0;
        bool keep_u = false;
        if (DEFAULT_RULES) {
          // union add / intersection min / difference subtract, no -du: eval_stream written out as straight-line code
          // (glistcompare.c:459-489 with calculate_freq :433-455)
          const uint32_t c = args.p.cutoff;
          const uint32_t f1 = in_a ? c1 : 0u, f2 = in_b ? c2 : 0u;
          const bool a_ok = f1 >= c, b_ok = f2 >= c;
          f = f1 + f2;
          keep_u = AUX && live && (a_ok || b_ok) && f != 0u;
          const uint32_t lo = f1 < f2 ? f1 : f2;
          const bool is_i = ((ops >> 1) & 1u) && live && in_a && in_b && a_ok && b_ok && lo != 0u;
          const bool is_d1 = ((ops >> 2) & 1u) && live && in_a && a_ok && !b_ok && f1 > f2;
          const bool is_d2 = ((ops >> 3) & 1u) && live && in_b && b_ok && !a_ok && f2 > f1;
          kind = is_i ? 1u : is_d1 ? 2u : is_d2 ? 3u : 0u;
          fx = is_i ? lo : is_d1 ? f1 - f2 : f2 - f1;
        } else {
          keep_u = AUX && live && eval_stream (args.p, 0, c1, c2, in_a, in_b, f);        // (AUX <=> the union is requested)
#pragma unroll
          for (int q = 3; q >= 1; q--) {          // at most one of them holds (exclusive cut-off tests)
            uint32_t g = 0;
            if (((ops >> q) & 1u) && live && eval_stream (args.p, q, c1, c2, in_a, in_b, g)) { kind = (uint32_t) q; fx = g; }
          }
        }
        o_key[sl] = key;
        if (AUX) o_fu[sl] = f;
        o_fx[sl] = fx;
        mask_u |= (keep_u ? 1u : 0u) << sl;
        kinds |= kind << (2 * sl);
      };
    if (m.flags & 4) merge_slots_interior<VT> (ka, ca, kb, cb, i0, d0, sink);     // full tile away from the ends of the lists: no cursor bounds
    else merge_slots<VT> (ka, ca, m.na, (m.flags & 1) != 0, kb, cb, m.nb, (m.flags & 2) != 0, i0, d0, sink);
    int cnt[4] = {__popc (mask_u), 0, 0, 0};
#pragma unroll
    for (int sl = 0; sl < VT; sl++) {
      const uint32_t kind = (kinds >> (2 * sl)) & 3u;
      cnt[1] += kind == 1u;
      cnt[2] += kind == 2u;
      cnt[3] += kind == 3u;
      if (AUX) acc_sum[0] += ((mask_u >> sl) & 1u) ? o_fu[sl] : 0u;
      acc_sum[1] += kind == 1u ? o_fx[sl] : 0u;
      acc_sum[2] += kind == 2u ? o_fx[sl] : 0u;
      acc_sum[3] += kind == 3u ? o_fx[sl] : 0u;
    }
#pragma unroll
    for (int q = 0; q < 4; q++) acc_n[q] += (unsigned) cnt[q];

    // block scan of the four survivor counts, 16 bits each in one 64-bit word (a tile holds < 2^16 slots)
    unsigned long long own = (unsigned long long) cnt[0] | ((unsigned long long) cnt[1] << 16) | ((unsigned long long) cnt[2] << 32) | ((unsigned long long) cnt[3] << 48);
    unsigned long long incl = own;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned long long t = __shfl_up_sync (0xffffffffu, incl, off);
      if (lane >= off) incl += t;
    }
    if (lane == 31) s_wcnt[it & 1][warp] = incl;
    consumer_sync<NC> ();           // also: every consumer is done reading this stage's inputs
    static_assert (NWARPS <= 32, "one lane per consumer warp");
    const unsigned long long wv = (lane < NWARPS) ? s_wcnt[it & 1][lane] : 0ull;
    // four warp reductions (REDUX) over the 32-bit halves instead of a shuffle scan of the 64-bit words: the 16-bit fields
    // cannot carry into each other (a tile holds fewer than 2^16 slots)
    const unsigned long long wb = (lane < warp) ? wv : 0ull;
    const unsigned long long tile_cnt = (unsigned long long) __reduce_add_sync (0xffffffffu, (unsigned) wv)
                                      | ((unsigned long long) __reduce_add_sync (0xffffffffu, (unsigned) (wv >> 32)) << 32);
    const unsigned long long before = ((unsigned long long) __reduce_add_sync (0xffffffffu, (unsigned) wb)
                                       | ((unsigned long long) __reduce_add_sync (0xffffffffu, (unsigned) (wb >> 32)) << 32)) + incl - own;      // exclusive prefixes of this thread
    int tc[4], pos[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      tc[q] = (int) ((tile_cnt >> (16 * q)) & 0xffffu);
      pos[q] = (int) ((before >> (16 * q)) & 0xffffu);
    }
    if (tid == 0) {
      s_mail[s].tile = m.tile;
#pragma unroll
      for (int q = 0; q < 4; q++) s_mail[s].cnt[q] = tc[q];
      mbar_arrive (&bar_agg[s]);
    }
    // rest region: [intersection | diff1 | diff2]; with AUX the union sits at the front of the stage itself, otherwise the
    // rest region IS the front of the stage (and there is no union)
    pos[2] += tc[1];
    pos[3] += tc[1] + tc[2];
    uint64_t *rk = rest_keys (s);
    uint32_t *rc = rest_cnts (s);
#pragma unroll
    for (int sl = 0; sl < VT; sl++) {
      if (AUX && ((mask_u >> sl) & 1u)) {
        sk[pos[0]] = o_key[sl];
        sc[pos[0]] = o_fu[sl];
        pos[0] += 1;
      }
      const uint32_t kind = (kinds >> (2 * sl)) & 3u;
      if (kind) {
        const int p = kind == 1u ? pos[1] : kind == 2u ? pos[2] : pos[3];
        rk[p] = o_key[sl];
        rc[p] = o_fx[sl];
        pos[1] += kind == 1u;
        pos[2] += kind == 2u;
        pos[3] += kind == 3u;
      }
    }
    __syncwarp ();
    if (lane == 0) mbar_arrive (&bar_comp[s]);
    if (++s == STAGES) { s = 0; ph ^= 1u; }
  }

  // header totals: one pair of atomics per CTA and output
#pragma unroll
  for (int q = 0; q < 4; q++) {
    acc_n[q] = warp_sum_u64 (acc_n[q]);
    acc_sum[q] = warp_sum_u64 (acc_sum[q]);
    if (lane == 0) {
      s_red[2 * q][warp] = acc_n[q];
      s_red[2 * q + 1][warp] = acc_sum[q];
    }
  }
  consumer_sync<NC> ();
  if (tid < 4 && ((ops >> tid) & 1u)) {
    unsigned long long n = 0, sum = 0;
#pragma unroll
    for (int w = 0; w < NWARPS; w++) {
      n += s_red[2 * tid][w];
      sum += s_red[2 * tid + 1][w];
    }
    unsigned long long *slot = args.hdr->totals[tid][blockIdx.x & (TOTAL_SLOTS - 1)];
    atomicAdd (slot, n);
    atomicAdd (slot + 1, sum);
  }
}

template <int NC, int VT, int S, bool AUX, bool DEFAULT_RULES>
cudaError_t launch_fused_one (const TileArgs &args, int sm_count, cudaStream_t st)
{
  using Cfg = FusedCfg<NC, VT, S, AUX>;
  auto kernel = setop2_fused_kernel<NC, VT, S, AUX, DEFAULT_RULES>;
  static bool configured = false;      // benign race: idempotent
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute (kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  uint64_t grid = (uint64_t) sm_count;
  if (grid > args.n_tiles) grid = args.n_tiles;
  kernel<<<(unsigned) grid, Cfg::NTHREADS, Cfg::SMEM_BYTES, st>>> (args);
  return cudaGetLastError ();
}

}  // namespace

// Several outputs of one two-list merge in one pass.  Applies when the outputs other than the union exclude each other
// (no -du).  The tile size differs by variant: ask fused_tile_slots first, it is what the partition must be cut for.
bool fused_applicable (const SetOpParams &p, uint32_t ops)
{
  if (p.sem != SEM_PAIR || p.subtract) return false;
  // measured at 1e9 + 1e9 (profiles/r02_multi_output.jsonl): three or four outputs, or two without the union, are faster
  // fused (-u -i -d 25.1 vs 29.4 ms, -u -i -dd 25.6 vs 39.0, -i -dd 19.5 vs 28.4); the union with ONE more output is not
  // (24.4 vs 20.0 ms: the union forces the small tiles of the auxiliary-buffer variant)
  const int n = __builtin_popcount (ops);
  return n >= 3 || (n == 2 && !(ops & OP_UNION));
}

uint32_t fused_tile_slots (uint32_t ops)
{
  return (ops & 1u) ? 512u * GT4_FUSED_VT_AUX : 512u * GT4_FUSED_VT_REST;
}

size_t fused_desc_bytes (uint64_t n_tiles) { return (size_t) n_tiles * 32; }

cudaError_t launch_setop2_fused (const TileArgs &args, int sm_count, cudaStream_t st)
{
  if (args.n_tiles == 0) return cudaSuccess;
  const SetOpParams &p = args.p;
  const bool dflt = (!(p.ops & 1u) || p.rule[0] == RULE_ADD) && (!(p.ops & 2u) || p.rule[1] == RULE_MIN) &&
                    (!(p.ops & 4u) || p.rule[2] == RULE_SUBTRACT) && (!(p.ops & 8u) || p.rule[3] == RULE_SUBTRACT);
  if (p.ops & 1u)
    return dflt ? launch_fused_one<512, GT4_FUSED_VT_AUX, GT4_FUSED_S_AUX, true, true> (args, sm_count, st)
                : launch_fused_one<512, GT4_FUSED_VT_AUX, GT4_FUSED_S_AUX, true, false> (args, sm_count, st);
  return dflt ? launch_fused_one<512, GT4_FUSED_VT_REST, GT4_FUSED_S_REST, false, true> (args, sm_count, st)
              : launch_fused_one<512, GT4_FUSED_VT_REST, GT4_FUSED_S_REST, false, false> (args, sm_count, st);
}

}  // namespace gt4gpu
