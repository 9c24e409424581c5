// gt4gpu_stream_kernel.cu -- the single-output merge kernel, second generation.
//
// setop2_stream_kernel is a persistent, warp-specialised sm_100a kernel.  One CTA = 10 warps:
//
//   warps 0-7  consumers   per tile: wait for the stage, per-thread co-rank + serial merge of VT
//                          slots out of shared memory (gt4gpu_core.cuh), predicate + count rule
//                          (compile-time fast paths), block scan of the survivors, hand the tile
//                          count to the look-back warp, store the PREVIOUS tile (whose global
//                          offset has arrived meanwhile) with coalesced stores, then compact the
//                          current tile's survivors into its own stage buffer
//   warp 8     producer    claims tiles in order from a global ticket, reads their co-ranks and
//                          stages the four slices (A keys, A counts, B keys, B counts, +1 halo /
//                          +1 peek) with 1-D TMA bulk copies (cp.async.bulk, 16-byte aligned
//                          over-fetch) that complete on the stage's "full" mbarrier
//   warp 9     look-back   decoupled look-back over the per-tile descriptors; runs one tile
//                          behind the consumers, so its L2 round trips are never on their
//                          critical path
//
// Stages cycle  fill (TMA) -> merge -> hold the compacted output -> store -> free  through
// mbarriers (full / empty / count posted / offset ready); with 3 stages the loads of tile n+2 are
// in flight while tile n+1 is merged and tile n is stored.  The only CTA-wide synchronisation of
// the consumers is one named barrier per tile (inside the scan).  Tiles are claimed through an
// atomic ticket, so a tile's predecessors are always owned by CTAs that are already running and
// the look-back cannot deadlock whatever the residency of the grid.
//
// HBM-bound integer work: no tensor cores.  Algorithmic traffic 12 B per input record + 12 B per
// output record.
#include <cuda_runtime.h>
#include <stdint.h>

#include "gt4gpu_internal.h"

namespace gt4gpu {

namespace {

constexpr int CONSUMERS = 256;                 // 8 consumer warps
constexpr int PRODUCER_WARP = CONSUMERS / 32;  // warp 8
constexpr int LOOKBACK_WARP = PRODUCER_WARP + 1;
constexpr int NTHREADS = CONSUMERS + 64;
constexpr int STAGES = 3;
constexpr uint64_t TILE_END = ~0ull;

constexpr uint64_t DESC_PARTIAL = 1ull << 62;
constexpr uint64_t DESC_INCLUSIVE = 2ull << 62;
constexpr uint64_t DESC_VALUE_MASK = (1ull << 62) - 1;

template <int VT>
struct StreamCfg {
  static constexpr int TILE = CONSUMERS * VT;
  // A and B slices are over-fetched to 16-byte boundaries on both sides and carry +1 halo / +1 peek;
  // merge_slots may read VT + 1 elements past a slice
  static constexpr int KSLOTS = (TILE + VT + 16 + 1) & ~1;
  static constexpr int CSLOTS = (TILE + VT + 28 + 3) & ~3;
  static constexpr size_t STAGE_BYTES = (size_t) KSLOTS * 8 + (size_t) CSLOTS * 4;
  static constexpr size_t SMEM_BYTES = STAGES * STAGE_BYTES;
};

struct StageMeta {
  uint64_t tile;     // TILE_END: no more work
  int na, nb;        // slice lengths
  int ka, kb;        // element offset of A[a_lo] / B[b_lo] inside the stage's key array
  int ca, cb;        // same inside the count array
  int flags;         // bit 0: halo present (A[a_lo - 1]), bit 1: peek present (B[b_hi])
};

struct Mailbox {
  uint64_t tile;
  uint64_t base;     // exclusive prefix of the tile's output count (written by the look-back warp)
  int cnt;           // the tile's output count (written by consumer thread 0)
};

// ---- PTX wrappers ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32 (const void *p) { return (uint32_t) __cvta_generic_to_shared (p); }

__device__ __forceinline__ void mbar_init (uint64_t *bar, uint32_t count)
{
  asm volatile ("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32 (bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_arrive (uint64_t *bar)
{
  asm volatile ("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32 (bar)) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx (uint64_t *bar, uint32_t bytes)
{
  asm volatile ("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32 (bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait (uint64_t *bar, uint32_t parity)
{
  uint32_t ok;
  asm volatile ("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(ok) : "r"(smem_u32 (bar)), "r"(parity) : "memory");
  return ok != 0;
}

__device__ __forceinline__ void mbar_wait (uint64_t *bar, uint32_t parity)
{
  while (!mbar_try_wait (bar, parity)) { }
}

// 1-D TMA bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s (void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
  asm volatile ("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                :: "r"(smem_u32 (dst)), "l"(src), "r"(bytes), "r"(smem_u32 (bar)) : "memory");
}

__device__ __forceinline__ void fence_proxy_async () { asm volatile ("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init () { asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void consumer_sync () { asm volatile ("bar.sync 1, %0;" :: "n"(CONSUMERS) : "memory"); }

__device__ __forceinline__ uint64_t ld_relaxed (const uint64_t *p)
{
  uint64_t v;
  asm volatile ("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void st_relaxed (uint64_t *p, uint64_t v)
{
  asm volatile ("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ uint64_t warp_sum_u64 (uint64_t v)
{
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync (0xffffffffu, v, off);
  return v;
}

// all 32 lanes of the look-back warp; returns the exclusive prefix of `aggregate`
__device__ __forceinline__ uint64_t lookback_exclusive (uint64_t *desc, uint64_t tile, uint64_t aggregate, int lane)
{
  if (tile == 0) {
    if (lane == 0) st_relaxed (desc, DESC_INCLUSIVE | aggregate);
    return 0;
  }
  if (lane == 0) st_relaxed (desc + tile, DESC_PARTIAL | aggregate);
  uint64_t exclusive = 0;
  int64_t pred = (int64_t) tile - 1;
  while (true) {
    const int64_t idx = pred - lane;
    uint64_t d = (idx >= 0) ? ld_relaxed (desc + idx) : DESC_INCLUSIVE;
    while (__any_sync (0xffffffffu, (d >> 62) == 0)) {
      if ((d >> 62) == 0) d = ld_relaxed (desc + idx);
    }
    const uint32_t incl = __ballot_sync (0xffffffffu, (d >> 62) == 2);
    if (incl) {
      const int first = __ffs (incl) - 1;
      exclusive += warp_sum_u64 (lane <= first ? (d & DESC_VALUE_MASK) : 0ull);
      break;
    }
    exclusive += warp_sum_u64 (d & DESC_VALUE_MASK);
    pred -= 32;
  }
  if (lane == 0) st_relaxed (desc + tile, DESC_INCLUSIVE | (exclusive + aggregate));
  return exclusive;
}

// ---- the kernel ------------------------------------------------------------------------------
template <int VT, int FAST, bool COUNT_ONLY>
__global__ void __launch_bounds__ (NTHREADS, 2)
setop2_stream_kernel (const TileArgs args)
{
  using Cfg = StreamCfg<VT>;
  constexpr int TILE = Cfg::TILE;

  extern __shared__ __align__ (128) unsigned char smem_raw[];
  __shared__ __align__ (8) uint64_t bar_full[STAGES];    // producer -> consumers: slices have landed (TMA tx)
  __shared__ __align__ (8) uint64_t bar_empty[STAGES];   // consumers -> producer: stage may be refilled
  __shared__ __align__ (8) uint64_t bar_agg[STAGES];     // consumers -> look-back: tile count posted
  __shared__ __align__ (8) uint64_t bar_base[STAGES];    // look-back -> consumers: global offset ready
  __shared__ StageMeta s_meta[STAGES];
  __shared__ Mailbox s_mail[STAGES];
  __shared__ int s_wcnt[2][CONSUMERS / 32];
  __shared__ unsigned long long s_red[2][CONSUMERS / 32];

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; s++) {
      mbar_init (&bar_full[s], 1);
      mbar_init (&bar_empty[s], CONSUMERS);
      mbar_init (&bar_agg[s], 1);
      mbar_init (&bar_base[s], 1);
    }
    fence_mbar_init ();
  }
  __syncthreads ();

  auto stage_keys = [&] (int s) { return reinterpret_cast<uint64_t *> (smem_raw + (size_t) s * Cfg::STAGE_BYTES); };
  auto stage_cnts = [&] (int s) { return reinterpret_cast<uint32_t *> (smem_raw + (size_t) s * Cfg::STAGE_BYTES + (size_t) Cfg::KSLOTS * 8); };

  // ============================================================================ producer
  if (warp == PRODUCER_WARP) {
    if (lane != 0) return;
    const uint64_t total = args.na + args.nb;
    const uint64_t n_tiles = args.n_tiles;
    uint64_t nxt = atomicAdd (&args.hdr->ticket, 1u);
    uint64_t nxt_lo = 0, nxt_hi = 0;
    if (nxt < n_tiles) { nxt_lo = args.part[nxt]; nxt_hi = args.part[nxt + 1]; }
    for (uint32_t it = 0;; it++) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1u;
      const uint64_t tile = nxt, a_lo = nxt_lo, a_hi = nxt_hi;
      if (tile < n_tiles) {     // claim the following tile now: its latency hides behind the wait below
        nxt = atomicAdd (&args.hdr->ticket, 1u);
        if (nxt < n_tiles) { nxt_lo = args.part[nxt]; nxt_hi = args.part[nxt + 1]; }
      }
      mbar_wait (&bar_empty[s], ph ^ 1u);
      if (tile >= n_tiles) {
        s_meta[s].tile = TILE_END;
        mbar_arrive (&bar_full[s]);
        break;
      }
      const uint64_t d_lo = tile * TILE;
      const uint64_t d_hi = (d_lo + TILE < total) ? d_lo + TILE : total;
      const uint64_t b_lo = d_lo - a_lo, b_hi = d_hi - a_hi;
      const int na = (int) (a_hi - a_lo), nb = (int) (b_hi - b_lo);
      const int halo = a_lo > 0 ? 1 : 0, peek = b_hi < args.nb ? 1 : 0;
      uint64_t *sk = stage_keys (s);
      uint32_t *sc = stage_cnts (s);

      // byte ranges to fetch, widened to 16-byte boundaries (TMA bulk copies need aligned address and size)
      const uintptr_t ak0 = (uintptr_t) (args.a_words + a_lo - halo), ak1 = (uintptr_t) (args.a_words + a_hi);
      const uintptr_t bk0 = (uintptr_t) (args.b_words + b_lo), bk1 = (uintptr_t) (args.b_words + b_hi + peek);
      const uintptr_t ac0 = (uintptr_t) (args.a_counts + a_lo - halo), ac1 = (uintptr_t) (args.a_counts + a_hi);
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
      m.flags = halo | (peek << 1);
      s_meta[s] = m;

      mbar_arrive_expect_tx (&bar_full[s], ak_bytes + bk_bytes + ac_bytes + bc_bytes);
      if (ak_bytes) bulk_g2s (sk, (const void *) ak0a, ak_bytes, &bar_full[s]);
      if (bk_bytes) bulk_g2s (sk + (ak_bytes >> 3), (const void *) bk0a, bk_bytes, &bar_full[s]);
      if (ac_bytes) bulk_g2s (sc, (const void *) ac0a, ac_bytes, &bar_full[s]);
      if (bc_bytes) bulk_g2s (sc + (ac_bytes >> 2), (const void *) bc0a, bc_bytes, &bar_full[s]);
    }
    return;
  }

  // ============================================================================ look-back
  if (warp == LOOKBACK_WARP) {
    if (COUNT_ONLY) return;
    for (uint32_t it = 0;; it++) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1u;
      mbar_wait (&bar_agg[s], ph);
      const uint64_t tile = s_mail[s].tile;
      if (tile == TILE_END) break;
      const uint64_t base = lookback_exclusive (args.desc, tile, (uint64_t) s_mail[s].cnt, lane);
      if (lane == 0) {
        s_mail[s].base = base;
        mbar_arrive (&bar_base[s]);
      }
      __syncwarp ();
    }
    return;
  }

  // ============================================================================ consumers
  const int stream = args.stream0;
  unsigned long long acc_n = 0, acc_sum = 0;   // this thread's share of the header totals
  int prev = -1, prev_cnt = 0;
  uint32_t prev_ph = 0;

  // store of a finished tile: its compacted records sit at the front of its stage buffer
  auto store_tile = [&] (int s, int cnt, uint32_t ph) {
    mbar_wait (&bar_base[s], ph);
    const uint64_t base = s_mail[s].base;
    const uint64_t *sk = stage_keys (s);
    const uint32_t *sc = stage_cnts (s);
    if (base + (uint64_t) cnt <= args.out_capacity[stream]) {
      uint64_t *ow = args.out_words[stream] + base;
      uint32_t *oc = args.out_counts[stream] + base;
#pragma unroll
      for (int r = 0; r < VT; r++) {
        const int x = tid + r * CONSUMERS;
        if (x < cnt) {
          ow[x] = sk[x];
          oc[x] = sc[x];
        }
      }
    } else if (tid == 0) {
      args.hdr->overflow = 1u;
    }
    fence_proxy_async ();          // generic accesses to the stage before the async proxy (TMA) refills it
    mbar_arrive (&bar_empty[s]);
  };

  for (uint32_t it = 0;; it++) {
    const int s = it % STAGES;
    const uint32_t ph = (it / STAGES) & 1u;
    mbar_wait (&bar_full[s], ph);
    const StageMeta m = s_meta[s];
    if (m.tile == TILE_END) {
      if (!COUNT_ONLY && tid == 0) {
        s_mail[s].tile = TILE_END;
        mbar_arrive (&bar_agg[s]);
      }
      break;
    }
    uint64_t *sk = stage_keys (s);
    uint32_t *sc = stage_cnts (s);
    const uint64_t *ka = sk + m.ka;
    const uint32_t *ca = sc + m.ca;
    const uint64_t *kb = sk + m.kb;
    const uint32_t *cb = sc + m.cb;
    const int n_tile = m.na + m.nb;
    const int d0 = (tid * VT < n_tile) ? tid * VT : n_tile;
    const int i0 = merge_path<int> (ka, m.na, kb, m.nb, d0);

    uint64_t o_key[VT];
    uint32_t o_freq[VT];
    uint32_t mask = 0;
    merge_slots<VT> (ka, ca, m.na, (m.flags & 1) != 0, kb, cb, m.nb, (m.flags & 2) != 0, i0, d0,
      [&] (int sl, uint64_t key, uint32_t c1, uint32_t c2, bool in_a, bool in_b, bool live) {
        uint32_t f = 0;
        const bool keep = eval_fast<FAST> (args.p, stream, c1, c2, in_a, in_b, f) && live;
        o_key[sl] = key;
        o_freq[sl] = f;
        mask |= (keep ? 1u : 0u) << sl;
      });
    const int cnt = __popc (mask);
    acc_n += (unsigned) cnt;
#pragma unroll
    for (int sl = 0; sl < VT; sl++) acc_sum += ((mask >> sl) & 1u) ? o_freq[sl] : 0u;

    if (COUNT_ONLY) {
      mbar_arrive (&bar_empty[s]);    // only generic reads touched the stage
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
    consumer_sync ();               // also: every consumer is done reading this stage's inputs
    int warp_prefix = 0, tile_cnt = 0;
#pragma unroll
    for (int w = 0; w < CONSUMERS / 32; w++) {
      const int v = s_wcnt[it & 1][w];
      if (w < warp) warp_prefix += v;
      tile_cnt += v;
    }
    if (tid == 0) {
      s_mail[s].tile = m.tile;
      s_mail[s].cnt = tile_cnt;
      mbar_arrive (&bar_agg[s]);    // the look-back warp takes it from here
    }

    // the previous tile's offset has had a whole merge phase to arrive
    if (prev >= 0) store_tile (prev, prev_cnt, prev_ph);

    // compact this tile's survivors to the front of its own stage buffer
    int pos = warp_prefix + incl - cnt;
#pragma unroll
    for (int sl = 0; sl < VT; sl++) {
      if ((mask >> sl) & 1u) {
        sk[pos] = o_key[sl];
        sc[pos] = o_freq[sl];
        pos += 1;
      }
    }
    prev = s;
    prev_cnt = tile_cnt;
    prev_ph = ph;
  }

  if (!COUNT_ONLY) {
    consumer_sync ();               // the last tile's compaction is complete
    if (prev >= 0) store_tile (prev, prev_cnt, prev_ph);
  }

  // header totals: one pair of atomics per CTA
  acc_n = warp_sum_u64 (acc_n);
  acc_sum = warp_sum_u64 (acc_sum);
  if (lane == 0) {
    s_red[0][warp] = acc_n;
    s_red[1][warp] = acc_sum;
  }
  consumer_sync ();
  if (tid == 0) {
    unsigned long long n = 0, sum = 0;
#pragma unroll
    for (int w = 0; w < CONSUMERS / 32; w++) {
      n += s_red[0][w];
      sum += s_red[1][w];
    }
    unsigned long long *slot = args.hdr->totals[stream][blockIdx.x & (TOTAL_SLOTS - 1)];
    atomicAdd (slot, n);
    atomicAdd (slot + 1, sum);
  }
}

// ---- launch ------------------------------------------------------------------------------------
template <int VT, int FAST, bool CO>
cudaError_t launch_stream_one (const TileArgs &args, int sm_count, cudaStream_t st)
{
  using Cfg = StreamCfg<VT>;
  static int ctas_per_sm = 0;      // benign race: idempotent
  auto kernel = setop2_stream_kernel<VT, FAST, CO>;
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
  if (grid > args.n_tiles) grid = args.n_tiles;
  kernel<<<(unsigned) grid, NTHREADS, Cfg::SMEM_BYTES, st>>> (args);
  return cudaGetLastError ();
}

template <int VT, bool CO>
cudaError_t launch_stream_fast (const TileArgs &args, int fast, int sm_count, cudaStream_t st)
{
  switch (fast) {
  case FAST_U_ADD:  return launch_stream_one<VT, FAST_U_ADD, CO> (args, sm_count, st);
  case FAST_I_MIN:  return launch_stream_one<VT, FAST_I_MIN, CO> (args, sm_count, st);
  case FAST_D_SUB:  return launch_stream_one<VT, FAST_D_SUB, CO> (args, sm_count, st);
  case FAST_NU_ADD: return launch_stream_one<VT, FAST_NU_ADD, CO> (args, sm_count, st);
  case FAST_NI_MIN: return launch_stream_one<VT, FAST_NI_MIN, CO> (args, sm_count, st);
  default:          return launch_stream_one<VT, FAST_GENERIC, CO> (args, sm_count, st);
  }
}

}  // namespace

#define GT4GPU_STREAM_VTS(X) X (7) X (9) X (11) X (13)

bool stream_shape_supported (int items)
{
#define X(VT) if (items == VT) return true;
  GT4GPU_STREAM_VTS (X)
#undef X
  return false;
}

int stream_tile_size (int items) { return CONSUMERS * items; }

cudaError_t launch_setop2_stream (const TileArgs &args, int items, bool count_only, int sm_count, cudaStream_t st)
{
  if (args.n_tiles == 0) return cudaSuccess;
  const int fast = select_fast_path (args.p, args.stream0);
#define X(VT)                                                                            \
  if (items == VT) return count_only ? launch_stream_fast<VT, true> (args, fast, sm_count, st) \
                                     : launch_stream_fast<VT, false> (args, fast, sm_count, st);
  GT4GPU_STREAM_VTS (X)
#undef X
  return cudaErrorInvalidValue;
}

}  // namespace gt4gpu
