// gt4gpu_internal.h -- launch interface between the C-ABI layer (gt4gpu_api.cu)
// and the sm_100a kernels (gt4gpu_kernels.cu).  Not installed; not part of the ABI.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "gt4gpu_core.cuh"

namespace gt4gpu {

// Experiment switches (environment GT4GPU_DEBUG, a bit mask).  Measurement-only bits are always honoured: 2 = look-back
// statistics, 4 = no L2 prefetch, 32 = per-phase cycle counters of the merge consumers.  Bits that change what a kernel
// computes (1 / 8 / 16: skip a look-back or the stores -- WRONG output, timing experiments only) are ignored unless the
// library is built with -DGT4GPU_UNSAFE_EXPERIMENTS.
int debug_flags ();

// Per-call scratch living in HBM, zeroed before every launch.
//   [0]      u32 ticket       -- dynamic tile counter (tiles are claimed in launch order so the
//                                decoupled look-back never waits on a tile that has not started)
//   [1]      u32 overflow     -- set when a caller-provided output buffer is too small
//   [2..]    u64 totals[4][TOTAL_SLOTS][2] -- per stream: records emitted, sum of emitted counts,
//                                spread over TOTAL_SLOTS addresses (tile % TOTAL_SLOTS) so the
//                                per-tile reductions do not serialise on one L2 atomic unit;
//                                the host adds the slots up
static constexpr int TOTAL_SLOTS = 64;
struct CallHeader {
  uint32_t ticket;
  uint32_t overflow;
  unsigned long long totals[4][TOTAL_SLOTS][2];
  unsigned long long dbg[8];      // GT4GPU_DEBUG statistics (look-back latency, consumer phase cycles); never read by the merge logic
};

struct TileArgs {
  const uint64_t *a_words;
  const uint32_t *a_counts;
  uint64_t na;
  const uint64_t *b_words;
  const uint32_t *b_counts;
  uint64_t nb;
  const uint64_t *part;        // n_tiles + 1 co-ranks: A index at diagonal t * TILE
  uint64_t n_tiles;
  uint64_t *out_words[4];
  uint32_t *out_counts[4];
  uint64_t out_capacity[4];
  CallHeader *hdr;
  uint64_t *desc;              // look-back descriptors, [stream slot][n_tiles]
  uint32_t tile_stride;        // stream kernel, count-only: > 1 visits every tile_stride-th tile only (density sample)
  int side_hint;               // stream kernel: the output is sparse, use the side-buffer variant where it exists
  int stream0;                 // the single requested stream when the 1-stream kernel is used
  int debug;                   // experiments only (GT4GPU_DEBUG): bit 0 = skip the look-back (WRONG output offsets)
  SetOpParams p;
};

struct TileShape { int threads, items; };

bool tile_shape_supported (int threads, int items);
size_t tile_smem_bytes (int threads, int items);

// co-rank of every tile boundary
cudaError_t launch_partition (const uint64_t *a, uint64_t na, const uint64_t *b, uint64_t nb,
                              uint32_t tile, uint64_t n_tiles, uint64_t *part, cudaStream_t st);
// the merge itself; n_streams is 1 or 4 (4 = fused multi-output pass)
cudaError_t launch_setop2 (const TileArgs &args, TileShape shape, int n_streams, bool count_only, cudaStream_t st);

// second-generation single-output kernel (gt4gpu_stream_kernel.cu): persistent, warp-specialised, TMA-staged;
// tiles are consumers * items merged slots
bool stream_shape_supported (int consumers, int items);
extern int g_stream_side;
// does a side-buffer variant exist for this output (and is it enabled)?
bool stream_side_capable (const SetOpParams &p, int stream, int consumers, int items);
cudaError_t launch_setop2_stream (const TileArgs &args, int consumers, int items, bool count_only, int sm_count, cudaStream_t st);

// multi-output kernel on the same pipeline (gt4gpu_fused_kernel.cu): one read of the lists, up to four outputs.
// Applies to two or more outputs without -du; its tiles (fused_tile_slots) differ from the single-output kernel's, and its
// look-back descriptors are 32 bytes per tile (fused_desc_bytes).
bool fused_applicable (const SetOpParams &p, uint32_t ops);
uint32_t fused_tile_slots (uint32_t ops);
size_t fused_desc_bytes (uint64_t n_tiles);
cudaError_t launch_setop2_fused (const TileArgs &args, int sm_count, cudaStream_t st);

// ---- single-pass N-list union / intersection (gt4gpu_kway_kernel.cu)
static constexpr int KWAY_MAX_LISTS = 8;        // lists per pass (their heads live in registers); more lists go through several passes
#ifndef GT4_KWAY_SAMPLE
#define GT4_KWAY_SAMPLE 128
#endif
static constexpr int KWAY_SAMPLE = GT4_KWAY_SAMPLE;         // every KWAY_SAMPLE-th word of every list is a boundary candidate
#ifndef GT4_KWAY_CAP
#define GT4_KWAY_CAP 4096
#endif
#ifndef GT4_KWAY_STAGES
#define GT4_KWAY_STAGES 3
#endif
static constexpr int KWAY_TILE_CAP = GT4_KWAY_CAP;      // records a tile can hold
static constexpr int KWAY_STAGES = GT4_KWAY_STAGES;     // tiles in flight per CTA (shared memory: stages + one scratch buffer)
static constexpr int KWAY_CONSUMERS = 256;
enum KwayMode : int { KWAY_MODE_GENERIC = 0, KWAY_MODE_U_ADD = 1, KWAY_MODE_I_MIN = 2 };

struct KwayArgs {
  const uint64_t *words[KWAY_MAX_LISTS];        // 16-byte aligned arrays; unused entries: n = 0
  const uint32_t *counts[KWAY_MAX_LISTS];
  uint64_t n[KWAY_MAX_LISTS];
  uint64_t sample_off[KWAY_MAX_LISTS];          // position of list j's samples in the sample array
  int n_lists;                 // lists taking part
  int n_real;                  // an intersection keeps a word found in n_real lists
  const uint64_t *cuts;        // [n_tiles + 1][NL]: first record of every tile in every list
  const uint64_t *bounds;      // [n_tiles + 1]: key range of tile t = [bounds[t], bounds[t + 1]]
  uint64_t n_tiles;
  uint64_t *out_words;
  uint32_t *out_counts;
  uint64_t out_capacity;
  CallHeader *hdr;
  uint64_t *desc;
  int op;                      // 0 union, 1 intersection
  int rule;                    // RULE_ADD / MAX / MIN / NUMBER
  uint32_t cutoff;
  uint32_t count_override;
  int final_pass;              // apply the cut-off (inner passes of a many-list call keep everything)
  int debug;
};

int kway_select_mode (int op, int rule);
cudaError_t launch_kway_samples (const KwayArgs &args, uint64_t n_samples, uint64_t *samples, cudaStream_t st);
// every-th sorted sample is a boundary; nl = 4 or 8 (the kernel variant: lists are padded to it)
cudaError_t launch_kway_cuts (const KwayArgs &args, const uint64_t *sorted_samples, uint64_t every, int nl, uint64_t *cuts, uint64_t *bounds, cudaStream_t st);
cudaError_t launch_kway_tiles (const KwayArgs &args, int nl, bool count_only, int sm_count, cudaStream_t st);

cudaError_t launch_deinterleave (const void *records, uint64_t n, uint64_t *words, uint32_t *counts, cudaStream_t st);
cudaError_t launch_interleave (const uint64_t *words, const uint32_t *counts, uint64_t n, void *records, cudaStream_t st);

// GT4I index records (16 bytes: word, first-location offset) -> SoA words / counts (offset differences)
cudaError_t launch_index16 (const void *records, uint64_t n, int has_next, uint64_t end_loc, uint64_t *words, uint32_t *counts, cudaStream_t st);

// counts[row(words_j[i]) * n_lists + j] = counts_j[i], rows found by binary search in `rows`
cudaError_t launch_scatter_counts (const uint64_t *rows, uint64_t n_rows, const uint64_t *words, const uint32_t *counts,
                                   uint64_t n, unsigned j, unsigned n_lists, uint32_t *matrix, cudaStream_t st);

// counts_out[i] = count of queries[i] (canonical form when canonize) in the list, 0 when absent
cudaError_t launch_lookup (const uint64_t *words, const uint32_t *counts, uint64_t n, unsigned k, int canonize,
                           const uint64_t *queries, uint64_t n_queries, uint64_t *canonical_out, uint32_t *counts_out, cudaStream_t st);
// the sorted-batch route: canonical form of every query; then, on the sorted queries, counts_out[perm[i]] = count of sorted[i]
cudaError_t launch_canonize (const uint64_t *queries, uint64_t n_queries, unsigned k, int canonize, uint64_t *canonical, cudaStream_t st);
cudaError_t launch_check_sorted (const uint64_t *keys, uint64_t n, uint32_t *unsorted, cudaStream_t st);
cudaError_t launch_lookup_sorted (const uint64_t *words, const uint32_t *counts, uint64_t n, const uint64_t *sorted_queries,
                                  const uint32_t *perm, uint64_t n_queries, uint32_t *counts_out, cudaStream_t st);

// ---- list building (gt4gpu_sort_kernel.cu): least-significant-digit radix sort of raw words + run-length counts
static constexpr int SORT_MAX_PASSES = 8;                                        // 8-bit digits of a 64-bit word
static constexpr size_t SORT_SCRATCH_HEAD = 2 * SORT_MAX_PASSES * 256 * 8 + 256;  // histograms, bin starts, tickets
static constexpr int SORT_OR_SLOT = 8;           // u64 slot after the tickets: OR of all keys (written by the histogram pass)
static constexpr size_t SORT_OR_OFFSET = 2 * SORT_MAX_PASSES * 256 * 8 + SORT_OR_SLOT * 8;   // its byte offset in the scratch
size_t sort_scratch_bytes (uint64_t n);
cudaError_t launch_radix_sort (const uint64_t *input, uint64_t *keys, uint64_t *alt, uint64_t n, int n_pass, unsigned char *scratch,
                               int sm_count, uint64_t **sorted, cudaStream_t st);
// same with a 32-bit payload per key; the first pass fills it with the key's index, so *sorted_vals is the sorting permutation
cudaError_t launch_radix_sort_pairs (uint64_t *keys, uint64_t *alt, uint32_t *vals, uint32_t *valt, uint64_t n, int n_pass,
                                     unsigned char *scratch, int sm_count, uint64_t **sorted, uint32_t **sorted_vals, cudaStream_t st);
size_t rle_scratch_bytes (uint64_t n);
cudaError_t launch_rle_heads (const uint64_t *sorted, uint64_t n, uint64_t *words_tmp, uint64_t *first, unsigned char *scratch,
                              int sm_count, unsigned long long **d_n_unique, cudaStream_t st);
cudaError_t launch_rle_counts (const uint64_t *words_tmp, const uint64_t *first, uint64_t n_unique, uint64_t n,
                               uint64_t *words, uint32_t *counts, cudaStream_t st);

// ---- FastA text -> canonical words on the device (gt4gpu_fasta_kernel.cu)
uint64_t fasta_chunks (uint64_t n);
size_t fasta_scratch_bytes (uint64_t n_chunks);
cudaError_t launch_fasta_codes (const uint8_t *text, uint64_t n, unsigned char *scratch, uint8_t *codes, const uint64_t **d_n_codes,
                                cudaStream_t st);
cudaError_t launch_fastq_codes (const uint8_t *text, uint64_t n, unsigned char *scratch, uint8_t *codes, const uint64_t **d_n_codes,
                                const uint32_t **d_malformed, const uint64_t **d_n_lines, cudaStream_t st);
cudaError_t launch_fasta_word_counts (const uint8_t *codes, uint64_t n_codes, unsigned k, unsigned char *scratch,
                                      const uint64_t **d_n_words, cudaStream_t st);
cudaError_t launch_fasta_words (const uint8_t *codes, uint64_t n_codes, unsigned k, const unsigned char *scratch, uint64_t *words,
                                cudaStream_t st);

}  // namespace gt4gpu
