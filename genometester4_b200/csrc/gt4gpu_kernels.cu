// gt4gpu_kernels.cu -- hand-written sm_100a kernels of the set-operation engine.
//
//   partition_kernel      merge-path co-rank of every tile boundary (one thread per diagonal)
//   setop2_tile_kernel    one CTA per tile of TILE = NT * VT merged slots:
//                           stage the tile's A and B slices (+1 halo / +1 peek) in shared memory,
//                           per-thread co-rank + serial merge of VT slots (gt4gpu_core.cuh),
//                           rule + cut-off per requested output stream,
//                           block scan of the survivors, decoupled look-back for the global
//                           offset, compaction through shared memory, coalesced SoA stores
//   deinterleave / interleave   12-byte AoS records <-> SoA, staged through shared memory
//   scatter_counts_kernel       count matrix of gt4_union / gt4_is_union
//
// Integer work only, HBM-bound: no tensor cores.  Algorithmic traffic is 12 B per input record
// read + 12 B per output record written.
#include <cuda_runtime.h>
#include <stdint.h>

#include "gt4gpu_device.cuh"
#include "gt4gpu_internal.h"

namespace gt4gpu {

using namespace dev;

// ------------------------------------------------------------------------------------------
// partition
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__ (256)
partition_kernel (const uint64_t *__restrict__ a, uint64_t na, const uint64_t *__restrict__ b, uint64_t nb,
                  uint32_t tile, uint64_t n_tiles, uint64_t *__restrict__ part)
{
  const uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t > n_tiles) return;
  const uint64_t total = na + nb;
  uint64_t diag = t * tile;
  if (diag > total) diag = total;
  part[t] = merge_path<uint64_t> (a, na, b, nb, diag);
}

// Two-level variant for long lists.  A full-range search touches ~log2(n) scattered cache lines
// (and TLB entries) per boundary; searching every PART_COARSE-th boundary first and the rest inside
// the window its two coarse neighbours span keeps all but ~1.5% of the searches inside a few MB.
static constexpr uint64_t PART_COARSE = 64;

__global__ void __launch_bounds__ (256)
partition_coarse_kernel (const uint64_t *__restrict__ a, uint64_t na, const uint64_t *__restrict__ b, uint64_t nb,
                         uint32_t tile, uint64_t n_tiles, uint64_t *__restrict__ part)
{
  uint64_t t = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) * PART_COARSE;
  if (t >= n_tiles + PART_COARSE) return;
  if (t > n_tiles) t = n_tiles;
  const uint64_t total = na + nb;
  uint64_t diag = t * tile;
  if (diag > total) diag = total;
  part[t] = merge_path<uint64_t> (a, na, b, nb, diag);
}

__global__ void __launch_bounds__ (256)
partition_fine_kernel (const uint64_t *__restrict__ a, uint64_t na, const uint64_t *__restrict__ b, uint64_t nb,
                       uint32_t tile, uint64_t n_tiles, uint64_t *__restrict__ part)
{
  const uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tiles || (t % PART_COARSE) == 0) return;     // coarse boundaries (and the last one) are done
  const uint64_t t0 = t - t % PART_COARSE;
  const uint64_t t1 = t0 + PART_COARSE < n_tiles ? t0 + PART_COARSE : n_tiles;
  part[t] = merge_path_window<uint64_t> (a, na, b, nb, t * tile, part[t0], part[t1]);
}

// ------------------------------------------------------------------------------------------
// the tile kernel
// ------------------------------------------------------------------------------------------
template <int NT, int VT>
struct TileSmem {
  static constexpr int TILE = NT * VT;
  static constexpr int SLOTS = TILE + VT + 4;   // halo + tile + peek + over-read slack of merge_slots
  static constexpr size_t BYTES = (size_t) SLOTS * (sizeof (uint64_t) + sizeof (uint32_t));
};

template <int NT, int VT, int NS, bool COUNT_ONLY>
__global__ void __launch_bounds__ (NT)
setop2_tile_kernel (const TileArgs args)
{
  constexpr int TILE = NT * VT;
  constexpr int SLOTS = TileSmem<NT, VT>::SLOTS;
  constexpr int NW = NT / 32;

  extern __shared__ __align__ (16) unsigned char smem_raw[];
  uint64_t *s_keys = reinterpret_cast<uint64_t *> (smem_raw);
  uint32_t *s_cnts = reinterpret_cast<uint32_t *> (s_keys + SLOTS);
  __shared__ uint64_t s_tile;
  __shared__ uint64_t s_base;
  __shared__ int s_warp_cnt[NW];
  __shared__ unsigned long long s_warp_sum[NW];

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;

  if (tid == 0) s_tile = atomicAdd (&args.hdr->ticket, 1u);
  __syncthreads ();
  const uint64_t tile = s_tile;

  const uint64_t total = args.na + args.nb;
  const uint64_t d_lo = tile * TILE;
  const uint64_t d_hi = (d_lo + TILE < total) ? d_lo + TILE : total;
  const uint64_t a_lo = args.part[tile];
  uint64_t a_hi = args.part[tile + 1];
  const bool sane = a_hi >= a_lo && a_hi - a_lo <= d_hi - d_lo;     // false only for inputs that are not strictly ascending
  if (!sane) {
    if (threadIdx.x == 0) args.hdr->overflow = 2u;
    a_hi = a_lo;
  }
  const uint64_t b_lo = d_lo - a_lo, b_hi = sane ? d_hi - a_hi : b_lo;
  const int na = (int) (a_hi - a_lo), nb = (int) (b_hi - b_lo);
  const bool has_halo = a_lo > 0;
  const bool has_peek = b_hi < args.nb;

  // ---- stage [halo | A slice | B slice | peek] : slot x <- A[a_lo - 1 + x] or B[b_lo + x - 1 - na]
  {
    const int n_stage = na + nb + 2;
    uint64_t k[VT];
    uint32_t c[VT];
#pragma unroll
    for (int r = 0; r < VT; r++) {
      const int x = tid + r * NT;
      k[r] = 0;
      c[r] = 0;
      if (x <= na) {
        if (x > 0 || has_halo) {
          k[r] = args.a_words[a_lo + x - 1];
          c[r] = args.a_counts[a_lo + x - 1];
        }
      } else if (x < n_stage) {
        const int j = x - 1 - na;
        if (j < nb || has_peek) {
          k[r] = args.b_words[b_lo + j];
          c[r] = args.b_counts[b_lo + j];
        }
      }
    }
#pragma unroll
    for (int r = 0; r < VT; r++) {
      const int x = tid + r * NT;
      s_keys[x] = k[r];
      s_cnts[x] = c[r];
    }
    // the last two staged slots (x = TILE, TILE + 1) when the tile is full
    if (tid < 2) {
      const int x = TILE + tid;
      uint64_t kk = 0;
      uint32_t cc = 0;
      if (x < n_stage) {
        if (x <= na) {
          if (x > 0 || has_halo) {
            kk = args.a_words[a_lo + x - 1];
            cc = args.a_counts[a_lo + x - 1];
          }
        } else {
          const int j = x - 1 - na;
          if (j < nb || has_peek) {
            kk = args.b_words[b_lo + j];
            cc = args.b_counts[b_lo + j];
          }
        }
      }
      s_keys[x] = kk;
      s_cnts[x] = cc;
    }
  }
  __syncthreads ();

  // ---- per-thread co-rank + serial merge of VT slots
  const uint64_t *ka = s_keys + 1;
  const uint32_t *ca = s_cnts + 1;
  const uint64_t *kb = ka + na;
  const uint32_t *cb = ca + na;
  const int n_tile = na + nb;
  const int d0 = (tid * VT < n_tile) ? tid * VT : n_tile;
  const int i0 = merge_path<int> (ka, na, kb, nb, d0);

  uint64_t o_key[VT];
  uint32_t o_freq[NS][VT];
  uint32_t o_mask[NS];
#pragma unroll
  for (int q = 0; q < NS; q++) o_mask[q] = 0;

  merge_slots<VT> (ka, ca, na, has_halo, kb, cb, nb, has_peek, i0, d0,
    [&] (int s, uint64_t key, uint32_t c1, uint32_t c2, bool in_a, bool in_b, bool live) {
      o_key[s] = key;
#pragma unroll
      for (int q = 0; q < NS; q++) {
        const int stream = (NS == 1) ? args.stream0 : q;
        const bool wanted = (NS == 1) ? true : ((args.p.ops >> q) & 1u);
        uint32_t f = 0;
        const bool keep = live && wanted && eval_stream (args.p, stream, c1, c2, in_a, in_b, f);
        o_freq[q][s] = f;
        o_mask[q] |= (keep ? 1u : 0u) << s;
      }
    });
  __syncthreads ();   // inputs consumed; shared memory becomes the compaction buffer

#pragma unroll
  for (int q = 0; q < NS; q++) {
    const int stream = (NS == 1) ? args.stream0 : q;
    if (NS > 1 && !((args.p.ops >> q) & 1u)) continue;   // uniform across the grid

    const int cnt = __popc (o_mask[q]);
    unsigned long long fsum = 0;
#pragma unroll
    for (int s = 0; s < VT; s++) fsum += ((o_mask[q] >> s) & 1u) ? o_freq[q][s] : 0u;

    // block scan of cnt, block sum of fsum
    int incl = cnt;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int t = __shfl_up_sync (0xffffffffu, incl, off);
      if (lane >= off) incl += t;
    }
    fsum = warp_sum_u64 (fsum);
    if (lane == 31) s_warp_cnt[warp] = incl;
    if (lane == 0) s_warp_sum[warp] = fsum;
    __syncthreads ();
    int warp_prefix = 0, tile_cnt = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) {
      const int v = s_warp_cnt[w];
      if (w < warp) warp_prefix += v;
      tile_cnt += v;
    }
    const int excl = warp_prefix + incl - cnt;

    if (tid == 0) {
      unsigned long long tile_sum = 0;
#pragma unroll
      for (int w = 0; w < NW; w++) tile_sum += s_warp_sum[w];
      unsigned long long *slot = args.hdr->totals[stream][tile & (TOTAL_SLOTS - 1)];
      atomicAdd (slot, (unsigned long long) tile_cnt);
      atomicAdd (slot + 1, tile_sum);
    }

    if (!COUNT_ONLY) {
      // warp 0 resolves the global offset while the other warps already compact
      if (warp == 0) {
        const uint64_t base = lookback_exclusive<1> (args.desc + (uint64_t) q * args.n_tiles, tile, (uint64_t) tile_cnt, lane);
        if (lane == 0) s_base = base;
      }
      int pos = excl;
#pragma unroll
      for (int s = 0; s < VT; s++) {
        if ((o_mask[q] >> s) & 1u) {
          s_keys[pos] = o_key[s];
          s_cnts[pos] = o_freq[q][s];
          pos += 1;
        }
      }
      __syncthreads ();
      const uint64_t base = s_base;
      if (base + (uint64_t) tile_cnt <= args.out_capacity[stream]) {
        uint64_t *ow = args.out_words[stream] + base;
        uint32_t *oc = args.out_counts[stream] + base;
        for (int x = tid; x < tile_cnt; x += NT) {
          ow[x] = s_keys[x];
          oc[x] = s_cnts[x];
        }
      } else if (tid == 0) {
        args.hdr->overflow = 1u;
      }
    }
    if (q + 1 < NS) __syncthreads ();   // compaction buffer and s_warp_* are reused by the next stream
  }
}

// ------------------------------------------------------------------------------------------
// AoS <-> SoA (12-byte packed records; /root/reference/src/word-map.h:89-99 is the AoS layout)
// ------------------------------------------------------------------------------------------
static constexpr int AOS_NT = 256;
static constexpr int AOS_PER_CTA = 1024;   // records per CTA

__global__ void __launch_bounds__ (AOS_NT)
deinterleave_kernel (const uint32_t *__restrict__ rec, uint64_t n, uint64_t *__restrict__ words, uint32_t *__restrict__ counts)
{
  __shared__ uint32_t s[AOS_PER_CTA * 3];
  const uint64_t first = (uint64_t) blockIdx.x * AOS_PER_CTA;
  const uint64_t left = n - first;
  const int m = left < AOS_PER_CTA ? (int) left : AOS_PER_CTA;
  const uint32_t *src = rec + first * 3;
  for (int x = threadIdx.x; x < m * 3; x += AOS_NT) s[x] = src[x];
  __syncthreads ();
  for (int r = threadIdx.x; r < m; r += AOS_NT) {
    words[first + r] = (uint64_t) s[3 * r] | ((uint64_t) s[3 * r + 1] << 32);
    counts[first + r] = s[3 * r + 2];
  }
}

__global__ void __launch_bounds__ (AOS_NT)
interleave_kernel (const uint64_t *__restrict__ words, const uint32_t *__restrict__ counts, uint64_t n, uint32_t *__restrict__ rec)
{
  __shared__ uint32_t s[AOS_PER_CTA * 3];
  const uint64_t first = (uint64_t) blockIdx.x * AOS_PER_CTA;
  const uint64_t left = n - first;
  const int m = left < AOS_PER_CTA ? (int) left : AOS_PER_CTA;
  for (int r = threadIdx.x; r < m; r += AOS_NT) {
    const uint64_t w = words[first + r];
    s[3 * r] = (uint32_t) w;
    s[3 * r + 1] = (uint32_t) (w >> 32);
    s[3 * r + 2] = counts[first + r];
  }
  __syncthreads ();
  uint32_t *dst = rec + first * 3;
  for (int x = threadIdx.x; x < m * 3; x += AOS_NT) dst[x] = s[x];
}

// GT4I index records (/root/reference/src/index-map.c:122-139): 16 bytes = u64 word + u64 offset of the word's first
// location; the count of a word is the distance to the next word's offset (num_locations closes the last one).
// rec holds n records, followed by one more record when has_next (its offset closes record n - 1), else end_loc does.
__global__ void __launch_bounds__ (256)
index16_kernel (const uint64_t *__restrict__ rec, uint64_t n, int has_next, uint64_t end_loc,
                uint64_t *__restrict__ words, uint32_t *__restrict__ counts)
{
  const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t loc = rec[2 * i + 1];
  const uint64_t next = (i + 1 < n || has_next) ? rec[2 * i + 3] : end_loc;
  words[i] = rec[2 * i];
  counts[i] = (uint32_t) (next - loc);
}

// counts of list j scattered into the row-major matrix; every word of the list is a row key
__global__ void __launch_bounds__ (256)
scatter_counts_kernel (const uint64_t *__restrict__ rows, uint64_t n_rows, const uint64_t *__restrict__ words,
                       const uint32_t *__restrict__ counts, uint64_t n, unsigned j, unsigned n_lists, uint32_t *__restrict__ matrix)
{
  const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t w = words[i];
  uint64_t lo = 0, hi = n_rows;
  while (lo < hi) {
    const uint64_t mid = lo + ((hi - lo) >> 1);
    if (rows[mid] < w) lo = mid + 1;
    else hi = mid;
  }
  if (lo < n_rows && rows[lo] == w) matrix[lo * n_lists + j] = counts[i];
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
#define GT4GPU_TILE_SHAPES(X) X (128, 15) X (128, 17) X (256, 7) X (256, 9) X (256, 11) X (256, 15) X (512, 7) X (512, 9)

bool tile_shape_supported (int threads, int items)
{
#define X(NT, VT) if (threads == NT && items == VT) return true;
  GT4GPU_TILE_SHAPES (X)
#undef X
  return false;
}

size_t tile_smem_bytes (int threads, int items)
{
  return (size_t) (threads * items + items + 4) * 12;
}

cudaError_t launch_partition (const uint64_t *a, uint64_t na, const uint64_t *b, uint64_t nb,
                              uint32_t tile, uint64_t n_tiles, uint64_t *part, cudaStream_t st)
{
  const uint64_t n = n_tiles + 1;
  const unsigned grid = (unsigned) ((n + 255) / 256);
  if (n_tiles < 16 * PART_COARSE) {
    partition_kernel<<<grid, 256, 0, st>>> (a, na, b, nb, tile, n_tiles, part);
    return cudaGetLastError ();
  }
  const uint64_t n_coarse = n_tiles / PART_COARSE + 2;
  partition_coarse_kernel<<<(unsigned) ((n_coarse + 255) / 256), 256, 0, st>>> (a, na, b, nb, tile, n_tiles, part);
  partition_fine_kernel<<<grid, 256, 0, st>>> (a, na, b, nb, tile, n_tiles, part);
  return cudaGetLastError ();
}

template <int NT, int VT, int NS, bool CO>
static cudaError_t launch_one (const TileArgs &args, cudaStream_t st)
{
  const size_t smem = TileSmem<NT, VT>::BYTES;
  static bool configured = false;   // benign race: the attribute is idempotent
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute (setop2_tile_kernel<NT, VT, NS, CO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  setop2_tile_kernel<NT, VT, NS, CO><<<(unsigned) args.n_tiles, NT, smem, st>>> (args);
  return cudaGetLastError ();
}

cudaError_t launch_setop2 (const TileArgs &args, TileShape shape, int n_streams, bool count_only, cudaStream_t st)
{
  if (args.n_tiles == 0) return cudaSuccess;
  if (args.n_tiles > 0x7fffffffull) return cudaErrorInvalidConfiguration;
#define X(NT, VT)                                                                          \
  if (shape.threads == NT && shape.items == VT) {                                          \
    if (n_streams == 1) return count_only ? launch_one<NT, VT, 1, true> (args, st)         \
                                          : launch_one<NT, VT, 1, false> (args, st);       \
    return count_only ? launch_one<NT, VT, 4, true> (args, st)                             \
                      : launch_one<NT, VT, 4, false> (args, st);                           \
  }
  GT4GPU_TILE_SHAPES (X)
#undef X
  return cudaErrorInvalidValue;
}

cudaError_t launch_deinterleave (const void *records, uint64_t n, uint64_t *words, uint32_t *counts, cudaStream_t st)
{
  if (n == 0) return cudaSuccess;
  const uint64_t grid = (n + AOS_PER_CTA - 1) / AOS_PER_CTA;
  deinterleave_kernel<<<(unsigned) grid, AOS_NT, 0, st>>> (static_cast<const uint32_t *> (records), n, words, counts);
  return cudaGetLastError ();
}

cudaError_t launch_interleave (const uint64_t *words, const uint32_t *counts, uint64_t n, void *records, cudaStream_t st)
{
  if (n == 0) return cudaSuccess;
  const uint64_t grid = (n + AOS_PER_CTA - 1) / AOS_PER_CTA;
  interleave_kernel<<<(unsigned) grid, AOS_NT, 0, st>>> (words, counts, n, static_cast<uint32_t *> (records));
  return cudaGetLastError ();
}

cudaError_t launch_index16 (const void *records, uint64_t n, int has_next, uint64_t end_loc, uint64_t *words, uint32_t *counts, cudaStream_t st)
{
  if (n == 0) return cudaSuccess;
  const uint64_t grid = (n + 255) / 256;
  index16_kernel<<<(unsigned) grid, 256, 0, st>>> (static_cast<const uint64_t *> (records), n, has_next, end_loc, words, counts);
  return cudaGetLastError ();
}

// ------------------------------------------------------------------------------------------
// batch lookups (word_map_lookup, src/word-map.c:134-163, for many words at once)
// ------------------------------------------------------------------------------------------
// 2-bit reverse complement of a k-mer word (get_reverse_complement, src/sequence.c:65-79), without the loop:
// complement, reverse the 32 two-bit groups of the 64-bit word, shift the k groups back down
__device__ __forceinline__ uint64_t reverse_complement (uint64_t w, unsigned k)
{
  w = ~w;
  w = ((w >> 2) & 0x3333333333333333ull) | ((w & 0x3333333333333333ull) << 2);
  w = ((w >> 4) & 0x0f0f0f0f0f0f0f0full) | ((w & 0x0f0f0f0f0f0f0f0full) << 4);
  w = __byte_perm ((uint32_t) (w >> 32), 0, 0x0123) | ((uint64_t) __byte_perm ((uint32_t) w, 0, 0x0123) << 32);
  return w >> (64 - 2 * k);
}

__global__ void __launch_bounds__ (256)
lookup_kernel (const uint64_t *__restrict__ words, const uint32_t *__restrict__ counts, uint64_t n, unsigned k, int canonize,
               const uint64_t *__restrict__ queries, uint64_t n_queries, uint64_t *__restrict__ canonical_out,
               uint32_t *__restrict__ counts_out)
{
  const uint64_t q = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_queries) return;
  uint64_t w = queries[q];
  if (canonize) {
    const uint64_t r = reverse_complement (w, k);
    if (r < w) w = r;
  }
  // lower bound; the upper levels of the search stay in L2 across the batch
  uint64_t lo = 0, hi = n;
  while (lo < hi) {
    const uint64_t mid = lo + ((hi - lo) >> 1);
    if (__ldg (words + mid) < w) lo = mid + 1;
    else hi = mid;
  }
  uint32_t c = 0;
  if (lo < n && __ldg (words + lo) == w) c = __ldg (counts + lo);
  if (canonical_out) canonical_out[q] = w;
  counts_out[q] = c;
}

__global__ void __launch_bounds__ (256)
canonize_kernel (const uint64_t *__restrict__ queries, uint64_t n_queries, unsigned k, int canonize, uint64_t *__restrict__ canonical)
{
  const uint64_t q = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_queries) return;
  uint64_t w = queries[q];
  if (canonize) {
    const uint64_t r = reverse_complement (w, k);
    if (r < w) w = r;
  }
  canonical[q] = w;
}

// queries in ascending order: neighbouring threads walk the same search path, so all but the last few probes hit L1 / L2
__global__ void __launch_bounds__ (256)
lookup_sorted_kernel (const uint64_t *__restrict__ words, const uint32_t *__restrict__ counts, uint64_t n,
                      const uint64_t *__restrict__ sorted_queries, const uint32_t *__restrict__ perm, uint64_t n_queries,
                      uint32_t *__restrict__ counts_out)
{
  const uint64_t q = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_queries) return;
  const uint64_t w = sorted_queries[q];
  uint64_t lo = 0, hi = n;
  while (lo < hi) {
    const uint64_t mid = lo + ((hi - lo) >> 1);
    if (__ldg (words + mid) < w) lo = mid + 1;
    else hi = mid;
  }
  uint32_t c = 0;
  if (lo < n && __ldg (words + lo) == w) c = __ldg (counts + lo);
  counts_out[perm[q]] = c;
}

// *unsorted |= 1 when some element is smaller than its predecessor
__global__ void __launch_bounds__ (256)
check_sorted_kernel (const uint64_t *__restrict__ keys, uint64_t n, uint32_t *unsorted)
{
  const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  const bool bad = i + 1 < n && keys[i] > keys[i + 1];
  if (__syncthreads_or (bad) && threadIdx.x == 0) atomicOr (unsorted, 1u);
}

cudaError_t launch_check_sorted (const uint64_t *keys, uint64_t n, uint32_t *unsorted, cudaStream_t st)
{
  cudaError_t e = cudaMemsetAsync (unsorted, 0, sizeof (uint32_t), st);
  if (e != cudaSuccess || n < 2) return e;
  check_sorted_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, st>>> (keys, n, unsorted);
  return cudaGetLastError ();
}

cudaError_t launch_canonize (const uint64_t *queries, uint64_t n_queries, unsigned k, int canonize, uint64_t *canonical, cudaStream_t st)
{
  if (n_queries == 0) return cudaSuccess;
  canonize_kernel<<<(unsigned) ((n_queries + 255) / 256), 256, 0, st>>> (queries, n_queries, k, canonize, canonical);
  return cudaGetLastError ();
}

cudaError_t launch_lookup_sorted (const uint64_t *words, const uint32_t *counts, uint64_t n, const uint64_t *sorted_queries,
                                  const uint32_t *perm, uint64_t n_queries, uint32_t *counts_out, cudaStream_t st)
{
  if (n_queries == 0) return cudaSuccess;
  lookup_sorted_kernel<<<(unsigned) ((n_queries + 255) / 256), 256, 0, st>>> (words, counts, n, sorted_queries, perm, n_queries, counts_out);
  return cudaGetLastError ();
}

cudaError_t launch_lookup (const uint64_t *words, const uint32_t *counts, uint64_t n, unsigned k, int canonize,
                           const uint64_t *queries, uint64_t n_queries, uint64_t *canonical_out, uint32_t *counts_out, cudaStream_t st)
{
  if (n_queries == 0) return cudaSuccess;
  const uint64_t grid = (n_queries + 255) / 256;
  if (grid > 0x7fffffffull) return cudaErrorInvalidConfiguration;
  lookup_kernel<<<(unsigned) grid, 256, 0, st>>> (words, counts, n, k, canonize, queries, n_queries, canonical_out, counts_out);
  return cudaGetLastError ();
}

cudaError_t launch_scatter_counts (const uint64_t *rows, uint64_t n_rows, const uint64_t *words, const uint32_t *counts,
                                   uint64_t n, unsigned j, unsigned n_lists, uint32_t *matrix, cudaStream_t st)
{
  if (n == 0) return cudaSuccess;
  const uint64_t grid = (n + 255) / 256;
  scatter_counts_kernel<<<(unsigned) grid, 256, 0, st>>> (rows, n_rows, words, counts, n, j, n_lists, matrix);
  return cudaGetLastError ();
}

}  // namespace gt4gpu
