// gt4gpu_device.cuh -- device helpers shared by the kernels: relaxed 64-bit descriptor accesses and the decoupled
// look-back over per-tile descriptors.  Not installed; not part of the ABI.
#pragma once

#include <stdint.h>

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

}  // namespace dev
}  // namespace gt4gpu
