// gt4gpu_stream_kernel.cu -- the single-output merge kernel, second generation.
//
// setop2_stream_kernel is a persistent, warp-specialised sm_100a kernel.  One CTA:
//
//   warps 0..NC/32-1  consumers   per tile: wait for the stage, per-thread co-rank + serial merge
//                          of VT slots out of shared memory (gt4gpu_core.cuh), predicate + count
//                          rule (compile-time fast paths), block scan of the survivors, hand the
//                          tile count to the look-back warp, compact the survivors to the front of
//                          the stage buffer.  They never wait for a global offset.
//   producer warp     claims tiles in order from a global ticket, reads their co-ranks and stages
//                          the four slices (A keys, A counts, B keys, B counts, +1 halo / +1 peek)
//                          with 1-D TMA bulk copies (cp.async.bulk, 16-byte aligned over-fetch)
//                          that complete on the stage's "full" mbarrier
//   splitter warp     works ahead of the consumers on filled stages: full-depth co-rank searches for
//                          every 8th consumer thread, so the consumers only search 8 * VT slots
//   S look-back warps decoupled look-back over the per-tile descriptors, one warp per stage so that
//                          several tiles of the CTA resolve their offsets concurrently
//   2 store warps     once a tile is compacted AND its global offset is known, copy its records
//                          from the stage buffer to the output arrays (coalesced) and hand the
//                          stage back to the producer
//
// Stages cycle  fill (TMA) -> merge + compact -> await offset -> store -> free  through mbarriers
// (full / compacted / count posted / offset ready / empty).  With S stages the consumers can run
// S - 2 tiles ahead of the slowest look-back, so the L2 round trips of the prefix chain and the
// skew between CTAs stay off their critical path.  The only CTA-wide synchronisation of the
// consumers is one named barrier per tile (inside the scan).  Tiles are claimed through an atomic
// ticket, so a tile's predecessors are always owned by CTAs that are already running and the
// look-back cannot deadlock whatever the residency of the grid.
//
// Side-buffer variants (template parameter SIDE, chosen per call from a density sample, gt4gpu_api.cu): an intersection or
// difference keeps few of a tile's slots, so its survivors are compacted into a small side buffer of the tile's output
// slot and the stage returns to the producer right after the merge instead of waiting for the look-back and the store.
// SIDE == 1: three stages, a side buffer of a third of a tile per stage; SIDE == 2: two stages and three output slots
// (mailbox, barriers, look-back warp, side buffer of 61 % of a tile) that are not tied to a stage.  A tile that keeps
// more than its side buffer holds compacts in place and keeps its stage, exactly as in the plain kernel.
//
// HBM-bound integer work: no tensor cores.  Algorithmic traffic 12 B per input record + 12 B per
// output record.
#include <cuda_runtime.h>
#include <stdint.h>

#include "gt4gpu_device.cuh"
#include "gt4gpu_internal.h"

namespace gt4gpu {

namespace {

#ifndef GT4_STORE_WARPS
#define GT4_STORE_WARPS 2
#endif
constexpr int STORE_WARPS = GT4_STORE_WARPS;
#ifndef GT4_INTERIOR_MERGE
#define GT4_INTERIOR_MERGE 1      // full tiles away from the ends of the lists take the merge loop without cursor bounds
#endif
#ifndef GT4_REDUX_SCAN
#define GT4_REDUX_SCAN 1
#endif
#ifndef GT4_STATIC_TILES
#define GT4_STATIC_TILES 0
#endif
#ifndef GT4_CLAIM_MODE
#define GT4_CLAIM_MODE 0          // when the producer claims a tile: 0 = one stage ahead, 1 = when the stage is free, 2 = when its previous tile has its offset
#endif
#ifndef GT4_HELPER_SLEEP_NS
#define GT4_HELPER_SLEEP_NS 0      // > 0: the helper warps sleep between two looks at a barrier they wait for
#endif
__device__ __forceinline__ void helper_wait (uint64_t *bar, uint32_t parity)
{
#if GT4_HELPER_SLEEP_NS > 0
  dev::mbar_wait_sleep (bar, parity, GT4_HELPER_SLEEP_NS);
#else
  dev::mbar_wait_relaxed (bar, parity);
#endif
}
constexpr uint64_t TILE_END = ~0ull;

using namespace dev;

// NC consumer threads (warps 0 .. NC/32-1), then the producer warp, the look-back warp and the store warps
template <int NC, int VT, int S, int SIDE = 0>
struct StreamCfg {
  static constexpr int CONSUMERS = NC;
  static constexpr int STAGES = S;
  static constexpr int PRODUCER_WARP = NC / 32;
  static constexpr int SPLITTER_WARP = NC / 32 + 1;
  // output slots: mailbox + barriers (+ side buffer) of a tile between the end of its merge and the end of its store.  One
  // per stage, except SIDE == 2: two stages, three slots with bigger side buffers
  static constexpr int OUT_SLOTS = (SIDE == 2) ? 3 : S;
  static constexpr int LOOKBACK_WARP0 = NC / 32 + 2;          // one look-back warp per output slot
  static constexpr int STORE_WARP0 = LOOKBACK_WARP0 + OUT_SLOTS;
  static constexpr int NTHREADS = NC + 64 + 32 * OUT_SLOTS + 32 * STORE_WARPS;
#ifndef GT4_SPLIT_POINTS
#define GT4_SPLIT_POINTS 32
#endif
  static constexpr int GROUP = NC / GT4_SPLIT_POINTS;         // consumer threads per coarse co-rank (the splitter warp searches 32 co-ranks per round)
  static constexpr int NSPLIT = NC / GROUP + 1;
  static constexpr int MIN_CTAS = (NC <= 256) ? 2 : 1;
  static constexpr int TILE = CONSUMERS * VT;
  // A and B slices are over-fetched to 16-byte boundaries on both sides and carry +1 halo / +1 peek;
  // merge_slots may read VT + 1 elements past a slice
  static constexpr int KSLOTS = (TILE + VT + 16 + 1) & ~1;
  static constexpr int CSLOTS = (TILE + VT + 28 + 3) & ~3;
  static constexpr size_t STAGE_BYTES = (size_t) KSLOTS * 8 + (size_t) CSLOTS * 4;
  // SIDE: sparse outputs (intersections, differences) are compacted into a small side buffer per stage, so the stage goes
  // back to the producer right after the merge instead of waiting for the tile's offset and store
  // (SIDE == 1: three stages, a side buffer of a third of a tile per stage; SIDE == 2: two stages, three side buffers of 61 %
  // of a tile for outputs of medium density)
  static constexpr int SIDE_CAP = SIDE == 1 ? 1536 : SIDE == 2 ? 2816 : 0;          // survivors a side buffer holds; fuller tiles keep their stage
  static constexpr size_t SIDE_BYTES = (size_t) SIDE_CAP * 12;
  static constexpr size_t SMEM_BYTES = S * STAGE_BYTES + OUT_SLOTS * SIDE_BYTES;
};

struct StageMeta {
  uint64_t tile;     // TILE_END: no more work
  int na, nb;        // slice lengths
  int ka, kb;        // element offset of A[a_lo] / B[b_lo] inside the stage's key array
  int ca, cb;        // same inside the count array
  int flags;         // bit 0: halo present (A[a_lo - 1]), bit 1: peek present (B[b_hi]), bit 2: full tile with both and A[a_hi]
};

struct Mailbox {
  uint64_t tile;
  uint64_t base;     // exclusive prefix of the tile's output count (written by the look-back warp)
  int cnt;           // the tile's output count (written by consumer thread 0)
  int side;          // SIDE kernels: 1 = the survivors sit in the slot's side buffer, 0 = at the front of the stage itself
  int stage;         // the stage the tile was merged in
};

// All 32 lanes of the look-back warp; returns the exclusive prefix of `aggregate`.
//
// The chain of prefixes has to advance as fast as tiles are produced (~50-100 tiles/us at the HBM
// roofline) and every hop costs an L2 round trip, so one hop inspects LB_W rows of 32 descriptors
// (row k, lane l -> tile pred - 32 k - l: every row is one coalesced 256-byte request) instead of
// the textbook single row.  Rows are consumed nearest-first; only a descriptor NEARER than the
// nearest inclusive one can make the warp wait, and then only its row is polled again (the shared
// implementation is lookback_exclusive<W> in gt4gpu_device.cuh).
#ifndef GT4_LB_W
#define GT4_LB_W 4
#endif
constexpr int LB_W = GT4_LB_W;

struct LookbackStats { unsigned long long cycles, polls, hops, calls; };

__device__ __forceinline__ uint64_t lookback_with_stats (uint64_t *desc, uint64_t tile, uint64_t aggregate, int lane, LookbackStats *stats)
{
  if (!stats) return lookback_exclusive<LB_W> (desc, tile, aggregate, lane);
  const long long t_begin = clock64 ();
  unsigned polls = 0, hops = 0;
  const uint64_t exclusive = lookback_exclusive<LB_W> (desc, tile, aggregate, lane, &polls, &hops);
  if (tile != 0) {
    stats->cycles += (unsigned long long) (clock64 () - t_begin);
    stats->polls += polls;
    stats->hops += hops;
    stats->calls += 1;
  }
  return exclusive;
}

// ---- the kernel ------------------------------------------------------------------------------
template <int NC, int VT, int S, int FAST, bool COUNT_ONLY, int SIDE = 0>
__global__ void __launch_bounds__ (StreamCfg<NC, VT, S, SIDE>::NTHREADS, StreamCfg<NC, VT, S, SIDE>::MIN_CTAS)
setop2_stream_kernel (const TileArgs args)
{
  using Cfg = StreamCfg<NC, VT, S, SIDE>;
  static_assert (!(SIDE && COUNT_ONLY), "the side buffers are an output path");
  constexpr int QS = Cfg::OUT_SLOTS;
  constexpr bool ONE_END = COUNT_ONLY || QS != S;      // the producer sends one END marker (else one per stage)
  constexpr int TILE = Cfg::TILE;
  constexpr int STAGES = S;
  constexpr int NWARPS = NC / 32;
  constexpr int PRODUCER_WARP = Cfg::PRODUCER_WARP;
  constexpr int SPLITTER_WARP = Cfg::SPLITTER_WARP;
  constexpr int LOOKBACK_WARP0 = Cfg::LOOKBACK_WARP0;
  constexpr int GROUP = Cfg::GROUP;
  constexpr int NSPLIT = Cfg::NSPLIT;
  constexpr int STORE_WARP0 = Cfg::STORE_WARP0;

  extern __shared__ __align__ (128) unsigned char smem_raw[];
  __shared__ __align__ (8) uint64_t bar_full[STAGES];    // producer -> consumers: slices have landed (TMA tx)
  __shared__ __align__ (8) uint64_t bar_split[STAGES];   // splitter -> consumers: coarse co-ranks ready (implies full)
  __shared__ __align__ (8) uint64_t bar_comp[QS];        // consumers -> store warps: survivors compacted
  __shared__ __align__ (8) uint64_t bar_agg[QS];         // consumers -> look-back: tile count posted
  __shared__ __align__ (8) uint64_t bar_base[QS];        // look-back -> store warps: global offset ready
  __shared__ __align__ (8) uint64_t bar_empty[STAGES];   // store warps (count-only: consumers; SIDE, sparse tile: consumer thread 0) -> producer
  __shared__ __align__ (8) uint64_t bar_done[QS];        // SIDE: store warps -> consumers: the slot's tile has been stored
  __shared__ StageMeta s_meta[STAGES];
  __shared__ Mailbox s_mail[QS];
  __shared__ int s_split[STAGES][NSPLIT];
  __shared__ int s_wcnt[2][NWARPS];
  __shared__ volatile unsigned int s_n_iter;            // tiles this CTA ended up processing (set when the END marker arrives)
  __shared__ unsigned long long s_red[2][NWARPS];

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; s++) {
      mbar_init (&bar_full[s], 1);
      mbar_init (&bar_split[s], 1);
      mbar_init (&bar_empty[s], COUNT_ONLY ? NWARPS : STORE_WARPS);
    }
#pragma unroll
    for (int s = 0; s < QS; s++) {
      mbar_init (&bar_comp[s], NWARPS);
      mbar_init (&bar_agg[s], 1);
      mbar_init (&bar_base[s], 1);
      mbar_init (&bar_done[s], STORE_WARPS);
    }
    s_n_iter = 0xffffffffu;
    fence_mbar_init ();
  }
  __syncthreads ();

  auto stage_keys = [&] (int s) { return reinterpret_cast<uint64_t *> (smem_raw + (size_t) s * Cfg::STAGE_BYTES); };
  auto stage_cnts = [&] (int s) { return reinterpret_cast<uint32_t *> (smem_raw + (size_t) s * Cfg::STAGE_BYTES + (size_t) Cfg::KSLOTS * 8); };
  auto side_keys = [&] (int s) { return reinterpret_cast<uint64_t *> (smem_raw + (size_t) STAGES * Cfg::STAGE_BYTES + (size_t) s * Cfg::SIDE_BYTES); };
  auto side_cnts = [&] (int s) { return reinterpret_cast<uint32_t *> (smem_raw + (size_t) STAGES * Cfg::STAGE_BYTES + (size_t) s * Cfg::SIDE_BYTES + (size_t) Cfg::SIDE_CAP * 8); };

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
#if GT4_STATIC_TILES
    // experiment: tile = round * grid + CTA (needs every CTA of the grid resident; fewer look-back polls, no balancing)
    uint64_t static_round = 0;
#define GT4_CLAIM_TICKET() (((uint64_t) blockIdx.x + (static_round++) * (uint64_t) gridDim.x) * tile_stride)
#else
#define GT4_CLAIM_TICKET() ((uint64_t) atomicAdd (&args.hdr->ticket, 1u) * tile_stride)
#endif
    const uint64_t tile_stride = args.tile_stride ? args.tile_stride : 1;      // > 1: a count-only pass over a sample of the tiles
    uint64_t nxt = GT4_CLAIM_TICKET ();
    uint64_t nxt_lo = 0, nxt_hi = 0;
    if (nxt < n_tiles) { nxt_lo = args.part[nxt]; nxt_hi = args.part[nxt + 1]; }
    // L2 prefetch of the tiles the grid will claim about two rounds from now (co-ranks loaded one iteration early)
    const uint64_t pf_dist = gridDim.x * tile_stride;
    uint64_t pf_tile = nxt + pf_dist, pf_lo = 0, pf_hi = 0;
    if (pf_tile < n_tiles) { pf_lo = args.part[pf_tile]; pf_hi = args.part[pf_tile + 1]; }
    int s = 0;
    uint32_t ph = 0;
    bool first = true;
    while (true) {
#if GT4_CLAIM_MODE >= 1
      // A ticket fixes the tile's place in the output order, and every later tile's look-back waits until this tile has
      // been merged: claim as LATE as possible.  (Mode 0 claimed the following tile before waiting for a free stage; the
      // wait varies between 0 and a tile period from CTA to CTA, and exactly that spread is what the look-backs of the
      // other CTAs then wait for.)  Mode 2 claims when the stage's previous tile has got its output offset, i.e. while its
      // store runs; mode 1 when the stage is free.
      if (!first) {
#if GT4_CLAIM_MODE == 2
        if (!COUNT_ONLY) helper_wait (&bar_base[s], ph ^ 1u);
        else helper_wait (&bar_empty[s], ph ^ 1u);
#else
        helper_wait (&bar_empty[s], ph ^ 1u);
#endif
        nxt = GT4_CLAIM_TICKET ();
        if (nxt < n_tiles) { nxt_lo = args.part[nxt]; nxt_hi = args.part[nxt + 1]; }
        pf_tile = nxt + pf_dist;
        if (pf_tile < n_tiles) { pf_lo = args.part[pf_tile]; pf_hi = args.part[pf_tile + 1]; }
      }
      first = false;
      const uint64_t tile = nxt, a_lo = nxt_lo, a_hi = nxt_hi;
      const uint64_t cur_pf = pf_tile, cur_pf_lo = pf_lo, cur_pf_hi = pf_hi;
#else
      (void) first;
      const uint64_t tile = nxt, a_lo = nxt_lo, a_hi = nxt_hi;
      const uint64_t cur_pf = pf_tile, cur_pf_lo = pf_lo, cur_pf_hi = pf_hi;
      if (tile < n_tiles) {     // claim the following tile now: its latency hides behind the wait below
        nxt = GT4_CLAIM_TICKET ();
        if (nxt < n_tiles) { nxt_lo = args.part[nxt]; nxt_hi = args.part[nxt + 1]; }
        pf_tile = nxt + pf_dist;
        if (pf_tile < n_tiles) { pf_lo = args.part[pf_tile]; pf_hi = args.part[pf_tile + 1]; }
      }
#endif
      if (!COUNT_ONLY && (args.debug & 4) == 0 && cur_pf < n_tiles && cur_pf_hi >= cur_pf_lo && cur_pf_hi - cur_pf_lo <= (uint64_t) TILE) {   // (the count-only pass is faster without it)
        const uint64_t pd_lo = cur_pf * TILE;
        const uint64_t pd_hi = (pd_lo + TILE < total) ? pd_lo + TILE : total;
        const uint64_t pb_lo = pd_lo - cur_pf_lo, pb_hi = pd_hi - cur_pf_hi;
        prefetch_l2 (args.a_words + cur_pf_lo, (cur_pf_hi - cur_pf_lo) * 8);
        prefetch_l2 (args.b_words + pb_lo, (pb_hi - pb_lo) * 8);
        prefetch_l2 (args.a_counts + cur_pf_lo, (cur_pf_hi - cur_pf_lo) * 4);
        prefetch_l2 (args.b_counts + pb_lo, (pb_hi - pb_lo) * 4);
      }
      helper_wait (&bar_empty[s], ph ^ 1u);
#if !GT4_STORE_FENCE
      fence_proxy_async ();      // the stage's last generic-proxy accesses (observed through bar_empty) before the TMA writes
#endif
      if (tile >= n_tiles) {
        // out of work: send an END marker through EVERY stage, in order and under the normal stage protocol
        // (each look-back warp owns one stage and must see its own marker; a barrier may never be advanced
        // twice before its waiter has looked)
        for (int q = 0; q < (ONE_END ? 1 : STAGES); q++) {
          if (q > 0) helper_wait (&bar_empty[s], ph ^ 1u);
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
    const bool sprof = (args.debug & 32) != 0;
    long long t_sfull = 0, t_ssearch = 0;
    while (true) {
      const long long s0 = sprof ? clock64 () : 0;
      mbar_wait (&bar_full[s], ph);
      const long long s1 = sprof ? clock64 () : 0;
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
      if (sprof) { t_sfull += s1 - s0; t_ssearch += clock64 () - s1; }
      if (m.tile == TILE_END && ++n_end == (ONE_END ? 1 : STAGES)) break;
      if (++s == STAGES) { s = 0; ph ^= 1u; }
    }
    if (sprof && lane == 0) {        // experiments: how long the splitter waits for the TMA, how long its searches take
      atomicAdd (&args.hdr->dbg[2], (unsigned long long) t_sfull);
      atomicAdd (&args.hdr->dbg[3], (unsigned long long) t_ssearch);
    }
    return;
  }

  // ============================================================================ look-back (one warp per stage)
  // A look-back is a handful of dependent L2 round trips (microseconds under full memory load), longer
  // than a tile period, so the tiles of one CTA are resolved by S warps in parallel: warp s owns stage s.
  if (warp >= LOOKBACK_WARP0 && warp < STORE_WARP0) {
    if (COUNT_ONLY) return;
    const int s = warp - LOOKBACK_WARP0;
    uint32_t ph = 0;
    LookbackStats stats = {0, 0, 0, 0};
    for (uint32_t it = (uint32_t) s;; it += QS, ph ^= 1u) {
      helper_wait (&bar_agg[s], ph);
      if (it >= s_n_iter) break;                  // the END marker, not a tile
      const uint64_t tile = s_mail[s].tile;
      const uint64_t base = (args.debug & 1) ? tile * TILE
                          : lookback_with_stats (args.desc, tile, (uint64_t) s_mail[s].cnt, lane, (args.debug & 2) ? &stats : nullptr);
      if (lane == 0) {
        s_mail[s].base = base;
        mbar_arrive (&bar_base[s]);
      }
      __syncwarp ();
    }
    if ((args.debug & 2) && lane == 0) {          // experiments: look-back latency statistics
      atomicAdd (&args.hdr->dbg[0], stats.cycles);
      atomicAdd (&args.hdr->dbg[1], stats.polls);
      atomicAdd (&args.hdr->dbg[2], stats.hops);
      atomicAdd (&args.hdr->dbg[3], stats.calls);
    }
    return;
  }

  // ============================================================================ store warps
  if (warp >= STORE_WARP0) {
    if (COUNT_ONLY) return;
    const int st_tid = tid - STORE_WARP0 * 32;
    constexpr int ST_THREADS = 32 * STORE_WARPS;
    const int stream = args.stream0;
    int s = 0;
    uint32_t ph = 0;
    while (true) {
      helper_wait (&bar_comp[s], ph);
      if (s_mail[s].tile == TILE_END) break;      // every real tile precedes the first END marker
      helper_wait (&bar_base[s], ph);
      const uint64_t base = s_mail[s].base;
      const int cnt = s_mail[s].cnt;
      const bool from_side = SIDE && s_mail[s].side != 0;
      const int stage = s_mail[s].stage;
      const uint64_t *sk = from_side ? side_keys (s) : stage_keys (stage);
      const uint32_t *sc = from_side ? side_cnts (s) : stage_cnts (stage);
      if (args.debug & 8) {
        // experiment: no stores
      } else if (base + (uint64_t) cnt > args.out_capacity[stream]) {
        if (st_tid == 0) args.hdr->overflow = 1u;
      } else {
        uint64_t *ow = args.out_words[stream] + base;
        uint32_t *oc = args.out_counts[stream] + base;
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
      }
#if GT4_STORE_FENCE
      fence_proxy_async ();          // generic accesses to the stage before the async proxy (TMA) refills it
#endif
      __syncwarp ();
      if (lane == 0) {
        if (!from_side) mbar_arrive (&bar_empty[stage]);   // (a tile in a side buffer gave its stage back long ago)
        if (SIDE) mbar_arrive (&bar_done[s]);
      }
      if (++s == QS) { s = 0; ph ^= 1u; }
    }
    return;
  }

  // ============================================================================ consumers
  const int stream = args.stream0;
  unsigned long long acc_n = 0, acc_sum = 0;   // this thread's share of the header totals
  int s = 0, n_end = 0;
  uint32_t ph = 0;
  int q = 0;                  // the tile's output slot and its phase (the same as s / ph unless there are more slots than stages)
  uint32_t qph = 0;
  const bool prof = (args.debug & 32) != 0;
  long long t_wait = 0, t_search = 0, t_merge = 0, t_scan = 0, t_scatter = 0, n_tiles_done = 0;
  for (uint32_t it = 0;; it++) {
    const long long c0 = prof ? clock64 () : 0;
    mbar_wait (&bar_split[s], ph);
    mbar_wait (&bar_full[s], ph);        // already complete; observed directly for the TMA-written data
    const StageMeta m = s_meta[s];
    if (m.tile == TILE_END) {
      if (COUNT_ONLY) break;
      // END markers arrive on S consecutive stages; pass each one on to that stage's look-back warp
      // (and the first one to the store warps)
      // every output slot passes an END marker on (each look-back warp owns a slot, the store warps take the first): one
      // per END marker when slots and stages coincide, all at once when the producer sends a single marker
      for (int e = 0; e < (QS != S ? QS : 1); e++) {
        if (SIDE) mbar_wait (&bar_done[q], qph ^ 1u);        // the store warps may still be reading the slot's previous mailbox
        if (tid == 0) {
          if (it < s_n_iter) s_n_iter = it;
          s_mail[q].tile = TILE_END;
          mbar_arrive (&bar_agg[q]);
        }
        __syncwarp ();
        if (lane == 0) mbar_arrive (&bar_comp[q]);
        if (++q == QS) { q = 0; qph ^= 1u; }
      }
      if (QS != S || ++n_end == STAGES) break;
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
    const long long c1 = prof ? clock64 () : 0;
    const int i0 = merge_path_window<int> (ka, m.na, kb, m.nb, d0, s_split[s][tid / GROUP], s_split[s][tid / GROUP + 1]);
    const long long c2 = prof ? clock64 () : 0;

    uint64_t o_key[VT];
    uint32_t o_freq[VT];
    uint32_t mask = 0;
    auto sink = [&] (int sl, uint64_t key, uint32_t c1, uint32_t c2, bool in_a, bool in_b, bool live) {
        uint32_t f = 0;
        const bool keep = eval_fast<FAST> (args.p, stream, c1, c2, in_a, in_b, f) && live;
        o_key[sl] = key;
        o_freq[sl] = f;
        mask |= (keep ? 1u : 0u) << sl;
      };
#if GT4_INTERIOR_MERGE
    if (m.flags & 4) merge_slots_interior<VT> (ka, ca, kb, cb, i0, d0, sink);
    else
#endif
    merge_slots<VT> (ka, ca, m.na, (m.flags & 1) != 0, kb, cb, m.nb, (m.flags & 2) != 0, i0, d0, sink);
    const int cnt = __popc (mask);
    acc_n += (unsigned) cnt;
#pragma unroll
    for (int sl = 0; sl < VT; sl++) acc_sum += ((mask >> sl) & 1u) ? o_freq[sl] : 0u;

    const long long c3 = prof ? clock64 () : 0;
    if (prof) { t_wait += c1 - c0; t_search += c2 - c1; t_merge += c3 - c2; n_tiles_done += 1; }
    if (COUNT_ONLY) {
      __syncwarp ();
      if (lane == 0) mbar_arrive (&bar_empty[s]);    // only generic reads touched the stage
      if (++s == STAGES) { s = 0; ph ^= 1u; }
      continue;
    }

    // block scan of the survivors (one named barrier; scratch double-buffered by tile parity)
    int incl = cnt;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int t = __shfl_up_sync (0xffffffffu, incl, off);
      if (lane >= off) incl += t;
    }
    if (lane == 31) s_wcnt[it & 1][warp] = incl;
    consumer_sync<NC> ();           // also: every consumer is done reading this stage's inputs
    const long long c4 = prof ? clock64 () : 0;
    // prefix over the (at most 32) warp totals with one more shuffle scan instead of a loop per thread
    static_assert (NWARPS <= 32, "one lane per consumer warp");
    const int wv = (lane < NWARPS) ? s_wcnt[it & 1][lane] : 0;
#if GT4_REDUX_SCAN
    // two warp reductions (REDUX) instead of a shuffle scan over the warp totals: every thread only needs the tile total and
    // the total of the warps before its own
    const int tile_cnt = (int) __reduce_add_sync (0xffffffffu, (unsigned) wv);
    const int warp_prefix = (int) __reduce_add_sync (0xffffffffu, (unsigned) (lane < warp ? wv : 0));
#else
    int wincl = wv;
#pragma unroll
    for (int off = 1; off < NWARPS; off <<= 1) {
      const int t = __shfl_up_sync (0xffffffffu, wincl, off);
      if (lane >= off) wincl += t;
    }
    const int tile_cnt = __shfl_sync (0xffffffffu, wincl, NWARPS - 1);
    const int warp_prefix = __shfl_sync (0xffffffffu, wincl - wv, warp);
#endif
    const bool to_side = SIDE && tile_cnt <= Cfg::SIDE_CAP;
    if (SIDE) mbar_wait (&bar_done[q], qph ^ 1u);     // the slot's previous tile has been stored: its mailbox and side buffer are free
    if (tid == 0) {
      s_mail[q].tile = m.tile;
      s_mail[q].cnt = tile_cnt;
      s_mail[q].side = to_side ? 1 : 0;
      s_mail[q].stage = s;
      mbar_arrive (&bar_agg[q]);    // the look-back warp takes it from here
    }
    uint64_t *dk = sk;
    uint32_t *dc = sc;
    if (to_side) {
      // few survivors: they go to the stage's side buffer and the stage itself returns to the producer NOW (every consumer
      // read its inputs before the named barrier above), one look-back + store earlier than otherwise
      if (tid == 0) mbar_arrive_n (&bar_empty[s], STORE_WARPS);      // standing in for the store warps, which never touch this stage
      dk = side_keys (q);
      dc = side_cnts (q);
    }

    // compact this tile's survivors (to the front of its own stage buffer unless to_side), then hand them to the store warps
    int pos = warp_prefix + incl - cnt;
#pragma unroll
    for (int sl = 0; sl < VT; sl++) {
      if ((mask >> sl) & 1u) {
        dk[pos] = o_key[sl];
        dc[pos] = o_freq[sl];
        pos += 1;
      }
    }
    __syncwarp ();
    if (lane == 0) mbar_arrive (&bar_comp[q]);

    if (prof) { const long long c5 = clock64 (); t_scan += c4 - c3; t_scatter += c5 - c4; }
    if (++s == STAGES) { s = 0; ph ^= 1u; }
    if (++q == QS) { q = 0; qph ^= 1u; }
  }
  if (prof && lane == 0) {       // experiments: per-phase cycles of the consumer warps
    atomicAdd (&args.hdr->dbg[0], (unsigned long long) t_wait); atomicAdd (&args.hdr->dbg[1], (unsigned long long) t_search);
    atomicAdd (&args.hdr->dbg[4], (unsigned long long) t_merge); atomicAdd (&args.hdr->dbg[5], (unsigned long long) t_scan);
    atomicAdd (&args.hdr->dbg[6], (unsigned long long) t_scatter); atomicAdd (&args.hdr->dbg[7], (unsigned long long) n_tiles_done);
  }

  // header totals: one pair of atomics per CTA
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
    unsigned long long *slot = args.hdr->totals[stream][blockIdx.x & (TOTAL_SLOTS - 1)];
    atomicAdd (slot, n);
    atomicAdd (slot + 1, sum);
  }
}

// ---- launch ------------------------------------------------------------------------------------
template <int NC, int VT, int S, int FAST, bool CO, int SIDE = 0>
cudaError_t launch_stream_one (const TileArgs &args, int sm_count, cudaStream_t st)
{
  using Cfg = StreamCfg<NC, VT, S, SIDE>;
  constexpr int NTHREADS = Cfg::NTHREADS;
  static int ctas_per_sm = 0;      // benign race: idempotent
  auto kernel = setop2_stream_kernel<NC, VT, S, FAST, CO, SIDE>;
  if (ctas_per_sm == 0) {
    cudaError_t e = cudaFuncSetAttribute (kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor (&occ, kernel, NTHREADS, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    ctas_per_sm = occ;
  }
  uint64_t grid = (uint64_t) sm_count * ctas_per_sm;
  const uint64_t n_claims = args.tile_stride > 1 ? (args.n_tiles + args.tile_stride - 1) / args.tile_stride : args.n_tiles;
  if (grid > n_claims) grid = n_claims;
  kernel<<<(unsigned) grid, NTHREADS, Cfg::SMEM_BYTES, st>>> (args);
  return cudaGetLastError ();
}

template <int NC, int VT, int S, bool CO>
cudaError_t launch_stream_fast (const TileArgs &args, int fast, int sm_count, cudaStream_t st)
{
  switch (fast) {
  case FAST_U_ADD:  return launch_stream_one<NC, VT, S, FAST_U_ADD, CO> (args, sm_count, st);
  case FAST_I_MIN:  return launch_stream_one<NC, VT, S, FAST_I_MIN, CO> (args, sm_count, st);
  case FAST_D_SUB:  return launch_stream_one<NC, VT, S, FAST_D_SUB, CO> (args, sm_count, st);
  case FAST_NU_ADD: return launch_stream_one<NC, VT, S, FAST_NU_ADD, CO> (args, sm_count, st);
  case FAST_NI_MIN: return launch_stream_one<NC, VT, S, FAST_NI_MIN, CO> (args, sm_count, st);
  case FAST_D2_SUB: return launch_stream_one<NC, VT, S, FAST_D2_SUB, CO> (args, sm_count, st);
  default:          return launch_stream_one<NC, VT, S, FAST_GENERIC, CO> (args, sm_count, st);
  }
}

}  // namespace

int g_stream_side = 1;      // option "stream_side": 1 = sparse outputs (TileArgs::side_hint) take the side-buffer variant, 2 = everything it is built for (measurements), 0 = never

bool stream_side_capable (const SetOpParams &p, int stream, int consumers, int items)
{
  if (g_stream_side != 1 || consumers != 512 || items != 9) return false;
  const int fast = select_fast_path (p, stream);
  return fast == FAST_I_MIN || fast == FAST_D_SUB || fast == FAST_D2_SUB || fast == FAST_NI_MIN;
}

// supported (consumer threads, items per thread, stages) triples; the stage count is fixed per shape by shared memory
#define GT4GPU_STREAM_SHAPES(X) X (256, 7, 5) X (256, 9, 4) X (256, 11, 3) X (384, 9, 5) X (384, 11, 4) X (512, 7, 5) X (512, 9, 4) X (512, 11, 3)

bool stream_shape_supported (int consumers, int items)
{
#define X(NC, VT, S) if (consumers == NC && items == VT) return true;
  GT4GPU_STREAM_SHAPES (X)
#undef X
  return false;
}

cudaError_t launch_setop2_stream (const TileArgs &args, int consumers, int items, bool count_only, int sm_count, cudaStream_t st)
{
  if (args.n_tiles == 0) return cudaSuccess;
  const int fast = select_fast_path (args.p, args.stream0);
  // outputs that are sparse as a rule (intersections, differences) at the default tile: the side-buffer variant, 3 stages
  if (!count_only && consumers == 512 && items == 9 && (g_stream_side >= 2 || (g_stream_side == 1 && args.side_hint))) {
    if (g_stream_side == 2 && fast == FAST_U_ADD) return launch_stream_one<512, 9, 3, FAST_U_ADD, false, 1> (args, sm_count, st);      // (measurements)
    // side_hint 1: sparse output, three stages + side buffers of a third of a tile; 2: medium density, two stages + three
    // side buffers of 61 % of a tile (option "stream_side" 2 / 3 force the one / the other)
    if ((g_stream_side == 1 && args.side_hint == 2) || g_stream_side == 3) {
      switch (fast) {
      case FAST_I_MIN:  return launch_stream_one<512, 9, 2, FAST_I_MIN, false, 2> (args, sm_count, st);
      case FAST_D_SUB:  return launch_stream_one<512, 9, 2, FAST_D_SUB, false, 2> (args, sm_count, st);
      case FAST_D2_SUB: return launch_stream_one<512, 9, 2, FAST_D2_SUB, false, 2> (args, sm_count, st);
      case FAST_NI_MIN: return launch_stream_one<512, 9, 2, FAST_NI_MIN, false, 2> (args, sm_count, st);
      default: break;
      }
    } else {
      switch (fast) {
      case FAST_I_MIN:  return launch_stream_one<512, 9, 3, FAST_I_MIN, false, 1> (args, sm_count, st);
      case FAST_D_SUB:  return launch_stream_one<512, 9, 3, FAST_D_SUB, false, 1> (args, sm_count, st);
      case FAST_D2_SUB: return launch_stream_one<512, 9, 3, FAST_D2_SUB, false, 1> (args, sm_count, st);
      case FAST_NI_MIN: return launch_stream_one<512, 9, 3, FAST_NI_MIN, false, 1> (args, sm_count, st);
      default: break;
      }
    }
  }
#define X(NC, VT, S)                                                                                        \
  if (consumers == NC && items == VT) return count_only ? launch_stream_fast<NC, VT, S, true> (args, fast, sm_count, st) \
                                                        : launch_stream_fast<NC, VT, S, false> (args, fast, sm_count, st);
  GT4GPU_STREAM_SHAPES (X)
#undef X
  return cudaErrorInvalidValue;
}

}  // namespace gt4gpu
