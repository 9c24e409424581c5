// gt4gpu_device.cuh -- device helpers shared by the kernels: relaxed 64-bit descriptor accesses and the decoupled
// look-back over per-tile descriptors.  Not installed; not part of the ABI.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

// The store warps read a stage with ordinary (generic-proxy) loads and hand it back to the producer through an mbarrier;
// the producer then lets the TMA (async proxy) overwrite it.  GT4_STORE_FENCE = 1 (default) puts the proxy fence into
// the store warps, after their last read of the stage; 0 puts it into the producer, right before the TMA copies.
// Measured: no difference (10.65 vs 10.69 ms on the headline merge, profiles/r02_stream_variants.txt).
#ifndef GT4_STORE_FENCE
#define GT4_STORE_FENCE 1
#endif

namespace gt4gpu {
namespace dev {

// look-back descriptor: one u64 per tile, status in the top 2 bits, value below.  A single relaxed 64-bit store
// publishes status and value together, so no fence is needed.
constexpr uint64_t DESC_PARTIAL = 1ull << 62;     // value = this tile's own count
constexpr uint64_t DESC_INCLUSIVE = 2ull << 62;   // value = inclusive prefix up to this tile
constexpr uint64_t DESC_VALUE_MASK = (1ull << 62) - 1;

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

// All 32 lanes of one warp; publishes `aggregate` for `tile` and returns its exclusive prefix over the tiles.  One hop
// inspects W rows of 32 descriptors (row k, lane l -> tile pred - 32 k - l: every row is one coalesced 256-byte
// request), nearest first: with hundreds of tiles in flight the nearest tile that already knows its prefix is often
// more than 32 tiles back, and every hop costs an L2 round trip.  Only a descriptor NEARER than the nearest inclusive
// one can make the warp wait, and then only its row is polled again.  polls / hops (optional) count for statistics.
template <int W>
__device__ __forceinline__ uint64_t lookback_exclusive (uint64_t *desc, uint64_t tile, uint64_t aggregate, int lane,
                                                        unsigned *polls = nullptr, unsigned *hops = nullptr)
{
  if (tile == 0) {
    if (lane == 0) st_relaxed (desc, DESC_INCLUSIVE | aggregate);
    return 0;
  }
  if (lane == 0) st_relaxed (desc + tile, DESC_PARTIAL | aggregate);
  uint64_t lane_sum = 0;
  int64_t pred = (int64_t) tile - 1 - lane;
  bool done = false;
  while (!done) {
    uint64_t d[W];
#pragma unroll
    for (int k = 0; k < W; k++) d[k] = (pred - 32 * k >= 0) ? ld_relaxed (desc + (pred - 32 * k)) : DESC_INCLUSIVE;
#pragma unroll
    for (int k = 0; k < W; k++) {
      if (done) break;
      while (true) {
        const uint32_t st = (uint32_t) (d[k] >> 62);
        const uint32_t m_wait = __ballot_sync (0xffffffffu, st == 0);
        const uint32_t m_incl = __ballot_sync (0xffffffffu, st == 2);
        const uint32_t m_stop = m_wait | m_incl;
        if (m_stop == 0) {                     // a full row of partial counts
          lane_sum += d[k] & DESC_VALUE_MASK;
          break;
        }
        const int first = __ffs (m_stop) - 1;
        if ((m_wait >> first) & 1u) {          // the nearest stopper has not posted yet: poll this row again
          if (polls) *polls += 1;
          d[k] = (pred - 32 * k >= 0) ? ld_relaxed (desc + (pred - 32 * k)) : DESC_INCLUSIVE;
          continue;
        }
        if (lane <= first) lane_sum += d[k] & DESC_VALUE_MASK;
        done = true;
        break;
      }
    }
    pred -= 32 * W;
    if (hops) *hops += 1;
  }
  const uint64_t exclusive = warp_sum_u64 (lane_sum);
  if (lane == 0) st_relaxed (desc + tile, DESC_INCLUSIVE | (exclusive + aggregate));
  return exclusive;
}

// ---- PTX wrappers shared by the pipelined kernels (mbarrier, 1-D TMA bulk copies, L2 prefetch) ----
__device__ __forceinline__ uint32_t smem_u32 (const void *p) { return (uint32_t) __cvta_generic_to_shared (p); }

__device__ __forceinline__ void mbar_init (uint64_t *bar, uint32_t count)
{
  asm volatile ("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32 (bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_arrive (uint64_t *bar)
{
  asm volatile ("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32 (bar)) : "memory");
}

// several arrivals at once (one thread standing in for `count` arrivers)
__device__ __forceinline__ void mbar_arrive_n (uint64_t *bar, uint32_t count)
{
  asm volatile ("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32 (bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx (uint64_t *bar, uint32_t bytes)
{
  asm volatile ("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32 (bar)), "r"(bytes) : "memory");
}

// try_wait suspends the thread in hardware until the phase completes or the time hint (ns) runs out, so a
// waiting warp does not burn issue slots polling
__device__ __forceinline__ bool mbar_try_wait (uint64_t *bar, uint32_t parity, uint32_t hint_ns)
{
  uint32_t ok;
  asm volatile ("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(ok) : "r"(smem_u32 (bar)), "r"(parity), "r"(hint_ns) : "memory");
  return ok != 0;
}

__device__ __forceinline__ void mbar_wait (uint64_t *bar, uint32_t parity)
{
  while (!mbar_try_wait (bar, parity, 2000u)) { }
}

// helper warps (producer, look-back, store) can afford to sleep longer
__device__ __forceinline__ void mbar_wait_relaxed (uint64_t *bar, uint32_t parity)
{
  while (!mbar_try_wait (bar, parity, 20000u)) { }
}

// same, giving the issue slots to the other warps between two looks (try_wait comes back long before its time hint)
__device__ __forceinline__ void mbar_wait_sleep (uint64_t *bar, uint32_t parity, unsigned ns)
{
  while (!mbar_try_wait (bar, parity, 20000u)) __nanosleep (ns);
}

// 1-D TMA bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s (void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
  asm volatile ("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                :: "r"(smem_u32 (dst)), "l"(src), "r"(bytes), "r"(smem_u32 (bar)) : "memory");
}

// L2 prefetch of a byte range (widened inwards to 16-byte boundaries; a few edge bytes do not matter)
__device__ __forceinline__ void prefetch_l2 (const void *p, uint64_t bytes)
{
  const uintptr_t lo = ((uintptr_t) p + 15) & ~(uintptr_t) 15;
  const uintptr_t hi = ((uintptr_t) p + bytes) & ~(uintptr_t) 15;
  if (hi > lo) asm volatile ("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(lo), "r"((uint32_t) (hi - lo)) : "memory");
}

__device__ __forceinline__ void fence_proxy_async () { asm volatile ("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init () { asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory"); }
template <int NC>
__device__ __forceinline__ void consumer_sync () { asm volatile ("bar.sync 1, %0;" :: "n"(NC) : "memory"); }

}  // namespace dev
}  // namespace gt4gpu
