// gt4gpu_fasta_kernel.cu -- FastA text -> canonical k-mer words on the device.
//
// Device form of fasta_reader_read_nwords (src/fasta.c:88-290) for FastA images, as glistmaker's read_table drives it
// (canonize = 1).  The reader is a byte-serial state machine; its state at any byte follows from two facts that can be
// computed in parallel:
//
//   * a byte belongs to a NAME iff a '>' occurs between the start of its line and itself (names start at any '>' met
//     in sequence state and run to the end of the line, :147-190; every line starts in sequence state except the
//     first, which the caller checks to start with '>')
//   * in sequence state nucleotides (ACGTU, either case) extend the current word, control characters (< ' ', i.e. line
//     ends) are transparent, every other character restarts the word (:221-269); the end of a name restarts it too
//
//   fasta_text_kernel<LINES>   per 4 KiB chunk of text: its line state (has a line end, '>' after the last line end)
//   line_carry_kernel          exclusive scan of the line states over the chunks (one CTA)
//   fasta_text_kernel<COUNT>   with the carried-in state: number of codes the chunk keeps
//   offsets_kernel             exclusive sum over the chunks (one CTA)
//   fasta_text_kernel<EMIT>    writes the compacted code stream (0..3 nucleotide, 4 = restart): names and transparent
//                              bytes are gone
//   fastq_text_kernel<...>     the same three passes for four-line FastQ records: the state is the line number mod 4
//   fasta_words_kernel<COUNT/EMIT>   a k-mer ends at code j iff codes j-k+1..j are nucleotides: count per chunk, then
//                              (after offsets_kernel) write the canonical words in file order
//
// HBM traffic: the text is read three times (1 B per byte each), the codes written once and read twice, 8 B written per
// word: with ~1 word per byte the word stream dominates.  Integer / byte work, no tensor cores.
#include <cuda_runtime.h>
#include <stdint.h>

#include "gt4gpu_fasta_core.cuh"
#include "gt4gpu_internal.h"

namespace gt4gpu {

namespace {

using namespace reader;

constexpr int FA_NT = 256;
constexpr int FA_BYTES = BYTES_PER_THREAD;         // bytes (or codes) per thread
constexpr int FA_CHUNK = FA_NT * FA_BYTES;         // 4 KiB per CTA
enum { MODE_LINES = 0, MODE_COUNT = 1, MODE_EMIT = 2 };

__device__ __forceinline__ uint32_t block_exclusive_sum (uint32_t v, uint32_t *s_warp, uint32_t *total)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t t = __shfl_up_sync (0xffffffffu, incl, off);
    if (lane >= off) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads ();
  uint32_t before = 0, sum = 0;
  for (int w = 0; w < FA_NT / 32; w++) {
    if (w < warp) before += s_warp[w];
    sum += s_warp[w];
  }
  *total = sum;
  return before + incl - v;
}

template <int MODE>
__global__ void __launch_bounds__ (FA_NT)
fasta_text_kernel (const uint8_t *__restrict__ text, uint64_t n, LineState *__restrict__ lines, const LineState *__restrict__ carry,
                   uint32_t *__restrict__ counts, const uint64_t *__restrict__ offsets, uint8_t *__restrict__ codes)
{
  __shared__ LineState s_ls[FA_NT / 32];
  __shared__ uint32_t s_sum[FA_NT / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint64_t base = (uint64_t) blockIdx.x * FA_CHUNK + (uint64_t) tid * FA_BYTES;

  uint8_t cls[FA_BYTES];
  if (base + FA_BYTES <= n && (reinterpret_cast<uintptr_t> (text + base) & 15) == 0) {
    const uint4 v = *reinterpret_cast<const uint4 *> (text + base);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < FA_BYTES; i++) cls[i] = classify ((uint8_t) (w[i >> 2] >> (8 * (i & 3))));
  } else {
#pragma unroll
    for (int i = 0; i < FA_BYTES; i++) cls[i] = (base + i < n) ? classify (text[base + i]) : CODE_SKIP;
  }
  // line state of this thread's 16 bytes, then of everything in the chunk before them
  const LineState mine = span_state (cls, FA_BYTES);
  LineState incl = mine;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    LineState o;
    o.has_nl = __shfl_up_sync (0xffffffffu, incl.has_nl, off);
    o.gt = __shfl_up_sync (0xffffffffu, incl.gt, off);
    if (lane >= off) incl = combine (o, incl);
  }
  if (lane == 31) s_ls[warp] = incl;
  __syncthreads ();
  LineState before = {0u, 0u};
  for (int w = 0; w < warp; w++) before = combine (before, s_ls[w]);
  if (MODE == MODE_LINES) {
    if (tid == FA_NT - 1) lines[blockIdx.x] = combine (before, incl);
    return;
  }
  LineState prev;
  prev.has_nl = __shfl_up_sync (0xffffffffu, incl.has_nl, 1);
  prev.gt = __shfl_up_sync (0xffffffffu, incl.gt, 1);
  if (lane == 0) { prev.has_nl = 0u; prev.gt = 0u; }
  const LineState cur0 = combine (carry[blockIdx.x], combine (before, prev));

  uint8_t out[FA_BYTES];
  const uint32_t n_out = (uint32_t) walk_fasta (cls, FA_BYTES, cur0, out);
  uint32_t total;
  const uint32_t at = block_exclusive_sum (n_out, s_sum, &total);
  if (MODE == MODE_COUNT) {
    if (tid == 0) counts[blockIdx.x] = total;
    return;
  }
  uint8_t *dst = codes + offsets[blockIdx.x] + at;
#pragma unroll
  for (int i = 0; i < FA_BYTES; i++) if (i < (int) n_out) dst[i] = out[i];
}

// FastQ (src/fasta.c:191-217, :272-295): records are four lines -- "@name", the sequence, "+...", the quality -- so the
// reader's state at a byte is its line number modulo 4, i.e. a prefix count of line ends.  Only sequence lines produce
// codes (plus one restart per record); a line 0 that does not start with '@' or a line 2 that does not start with '+'
// is where the reference's reader gives up: the kernel only raises *malformed and the caller hands such images to the
// serial host reader, which reproduces the reference's partial result.
//   MODE_LINES: newline count of the chunk.  MODE_COUNT / MODE_EMIT: carry = line number at the start of the chunk.
template <int MODE>
__global__ void __launch_bounds__ (FA_NT)
fastq_text_kernel (const uint8_t *__restrict__ text, uint64_t n, uint32_t *__restrict__ newlines, const uint64_t *__restrict__ line0,
                   uint32_t *__restrict__ counts, const uint64_t *__restrict__ offsets, uint8_t *__restrict__ codes,
                   uint32_t *__restrict__ malformed)
{
  __shared__ uint32_t s_sum[FA_NT / 32];
  const int tid = threadIdx.x;
  const uint64_t base = (uint64_t) blockIdx.x * FA_CHUNK + (uint64_t) tid * FA_BYTES;
  uint8_t ch[FA_BYTES];
  if (base + FA_BYTES <= n && (reinterpret_cast<uintptr_t> (text + base) & 15) == 0) {
    const uint4 v = *reinterpret_cast<const uint4 *> (text + base);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < FA_BYTES; i++) ch[i] = (uint8_t) (w[i >> 2] >> (8 * (i & 3)));
  } else {
#pragma unroll
    for (int i = 0; i < FA_BYTES; i++) ch[i] = (base + i < n) ? text[base + i] : (uint8_t) 1;    // padding: a transparent control character
  }
  uint32_t n_nl = 0;
#pragma unroll
  for (int i = 0; i < FA_BYTES; i++) n_nl += (ch[i] == '\n' && base + i < n) ? 1u : 0u;
  uint32_t total;
  const uint32_t nl_before = block_exclusive_sum (n_nl, s_sum, &total);
  if (MODE == MODE_LINES) {
    if (tid == 0) newlines[blockIdx.x] = total;
    return;
  }
  __syncthreads ();                   // s_sum is reused below
  const uint64_t line = line0[blockIdx.x] + nl_before;
  const bool at_line_start = base == 0 || (base <= n && text[base - 1] == '\n');
  uint8_t out[FA_BYTES];
  bool bad = false;
  const int n_mine = base >= n ? 0 : (n - base < (uint64_t) FA_BYTES ? (int) (n - base) : FA_BYTES);
  const uint32_t n_out = (uint32_t) walk_fastq (ch, n_mine, line, at_line_start, out, &bad);
  if (bad) atomicOr (malformed, 1u);
  const uint32_t at = block_exclusive_sum (n_out, s_sum, &total);
  if (MODE == MODE_COUNT) {
    if (tid == 0) counts[blockIdx.x] = total;
    return;
  }
  uint8_t *dst = codes + offsets[blockIdx.x] + at;
#pragma unroll
  for (int i = 0; i < FA_BYTES; i++) if (i < (int) n_out) dst[i] = out[i];
}

// exclusive scan of the chunks' line states; one CTA of 1024 threads, each owning a contiguous run of chunks
__global__ void __launch_bounds__ (1024)
line_carry_kernel (const LineState *__restrict__ lines, uint64_t n_chunks, LineState *__restrict__ carry)
{
  __shared__ LineState s_part[1024];
  const uint64_t per = (n_chunks + 1023) / 1024;
  const uint64_t lo = per * threadIdx.x, hi = lo + per < n_chunks ? lo + per : n_chunks;
  LineState acc = {0u, 0u};
  for (uint64_t i = lo; i < hi; i++) acc = combine (acc, lines[i]);
  s_part[threadIdx.x] = acc;
  __syncthreads ();
  if (threadIdx.x == 0) {
    LineState run = {0u, 0u};
    for (int t = 0; t < 1024; t++) {
      const LineState own = s_part[t];
      s_part[t] = run;
      run = combine (run, own);
    }
  }
  __syncthreads ();
  acc = s_part[threadIdx.x];
  for (uint64_t i = lo; i < hi; i++) {
    carry[i] = acc;
    acc = combine (acc, lines[i]);
  }
}

// exclusive sum of per-chunk counts -> 64-bit offsets; offsets[n_chunks] = total
__global__ void __launch_bounds__ (1024)
offsets_kernel (const uint32_t *__restrict__ counts, uint64_t n_chunks, uint64_t *__restrict__ offsets)
{
  __shared__ unsigned long long s_part[1024];
  const uint64_t per = (n_chunks + 1023) / 1024;
  const uint64_t lo = per * threadIdx.x, hi = lo + per < n_chunks ? lo + per : n_chunks;
  unsigned long long acc = 0;
  for (uint64_t i = lo; i < hi; i++) acc += counts[i];
  s_part[threadIdx.x] = acc;
  __syncthreads ();
  if (threadIdx.x == 0) {
    unsigned long long run = 0;
    for (int t = 0; t < 1024; t++) {
      const unsigned long long own = s_part[t];
      s_part[t] = run;
      run += own;
    }
    offsets[n_chunks] = run;
  }
  __syncthreads ();
  acc = s_part[threadIdx.x];
  for (uint64_t i = lo; i < hi; i++) {
    offsets[i] = acc;
    acc += counts[i];
  }
}

// codes -> canonical words.  Thread t of a chunk owns code positions [j0, j0 + 16) and replays the k - 1 codes before
// them to prime its window.
template <int MODE>
__global__ void __launch_bounds__ (FA_NT)
fasta_words_kernel (const uint8_t *__restrict__ codes, uint64_t n_codes, unsigned k, uint32_t *__restrict__ counts,
                    const uint64_t *__restrict__ offsets, uint64_t *__restrict__ words)
{
  __shared__ uint32_t s_sum[FA_NT / 32];
  const uint64_t j0 = (uint64_t) blockIdx.x * FA_CHUNK + (uint64_t) threadIdx.x * FA_BYTES;
  uint64_t out[FA_BYTES];
  uint32_t n_out = 0;
  if (j0 < n_codes) n_out = (uint32_t) window_words (codes, j0, j0 + FA_BYTES < n_codes ? j0 + FA_BYTES : n_codes, k, out);
  uint32_t total;
  const uint32_t at = block_exclusive_sum (n_out, s_sum, &total);
  if (MODE == MODE_COUNT) {
    if (threadIdx.x == 0) counts[blockIdx.x] = total;
    return;
  }
  uint64_t *dst = words + offsets[blockIdx.x] + at;
  for (uint32_t i = 0; i < n_out; i++) dst[i] = out[i];
}

}  // namespace

// ------------------------------------------------------------------------------------------
// launchers (the caller allocates; see gt4gpu_fasta_words_device in gt4gpu_api.cu)
// ------------------------------------------------------------------------------------------
uint64_t fasta_chunks (uint64_t n) { return (n + FA_CHUNK - 1) / FA_CHUNK; }

// scratch per chunk: line state (8) + carry (8) + count (4) + offset (8, one extra); FastQ: line number (8) + offset (8) + 2 counts (8) + flag
size_t fasta_scratch_bytes (uint64_t n_chunks) { return (size_t) (n_chunks + 2) * 32; }

// text (device, n bytes) -> codes (device, capacity n + n_chunks... at most one code per byte); *d_n_codes (device u64) is
// offsets[n_chunks] inside the scratch
cudaError_t launch_fasta_codes (const uint8_t *text, uint64_t n, unsigned char *scratch, uint8_t *codes, const uint64_t **d_n_codes,
                                cudaStream_t st)
{
  const uint64_t nc = fasta_chunks (n);
  if (nc == 0 || nc > 0x7fffffffull) return nc ? cudaErrorInvalidConfiguration : cudaSuccess;
  LineState *lines = reinterpret_cast<LineState *> (scratch);
  LineState *carry = lines + (nc + 1);
  uint64_t *offsets = reinterpret_cast<uint64_t *> (carry + (nc + 1));
  uint32_t *counts = reinterpret_cast<uint32_t *> (offsets + (nc + 1));
  fasta_text_kernel<MODE_LINES><<<(unsigned) nc, FA_NT, 0, st>>> (text, n, lines, nullptr, nullptr, nullptr, nullptr);
  line_carry_kernel<<<1, 1024, 0, st>>> (lines, nc, carry);
  fasta_text_kernel<MODE_COUNT><<<(unsigned) nc, FA_NT, 0, st>>> (text, n, nullptr, carry, counts, nullptr, nullptr);
  offsets_kernel<<<1, 1024, 0, st>>> (counts, nc, offsets);
  fasta_text_kernel<MODE_EMIT><<<(unsigned) nc, FA_NT, 0, st>>> (text, n, nullptr, carry, nullptr, offsets, codes);
  *d_n_codes = offsets + nc;
  return cudaGetLastError ();
}

// FastQ text -> codes; *d_malformed (device u32) is non-zero when a record line starts with the wrong tag; *d_n_lines
// (device u64) is the number of line ends, from which the caller sees an image that stops inside a '+' line
cudaError_t launch_fastq_codes (const uint8_t *text, uint64_t n, unsigned char *scratch, uint8_t *codes, const uint64_t **d_n_codes,
                                const uint32_t **d_malformed, const uint64_t **d_n_lines, cudaStream_t st)
{
  const uint64_t nc = fasta_chunks (n);
  if (nc == 0 || nc > 0x7fffffffull) return nc ? cudaErrorInvalidConfiguration : cudaSuccess;
  // scratch: line0 u64 [nc + 1] | offsets u64 [nc + 1] | newlines u32 [nc] | counts u32 [nc] | malformed u32
  uint64_t *line0 = reinterpret_cast<uint64_t *> (scratch);
  uint64_t *offsets = line0 + (nc + 1);
  uint32_t *newlines = reinterpret_cast<uint32_t *> (offsets + (nc + 1));
  uint32_t *counts = newlines + nc;
  uint32_t *malformed = counts + nc;
  cudaError_t e = cudaMemsetAsync (malformed, 0, sizeof (uint32_t), st);
  if (e != cudaSuccess) return e;
  fastq_text_kernel<MODE_LINES><<<(unsigned) nc, FA_NT, 0, st>>> (text, n, newlines, nullptr, nullptr, nullptr, nullptr, nullptr);
  offsets_kernel<<<1, 1024, 0, st>>> (newlines, nc, line0);
  fastq_text_kernel<MODE_COUNT><<<(unsigned) nc, FA_NT, 0, st>>> (text, n, nullptr, line0, counts, nullptr, nullptr, malformed);
  offsets_kernel<<<1, 1024, 0, st>>> (counts, nc, offsets);
  fastq_text_kernel<MODE_EMIT><<<(unsigned) nc, FA_NT, 0, st>>> (text, n, nullptr, line0, nullptr, offsets, codes, malformed);
  *d_n_codes = offsets + nc;
  *d_malformed = malformed;
  *d_n_lines = line0 + nc;          // line ends in the whole image
  return cudaGetLastError ();
}

// codes -> number of words (pass 1: fills the scratch offsets; *d_n_words = device u64)
cudaError_t launch_fasta_word_counts (const uint8_t *codes, uint64_t n_codes, unsigned k, unsigned char *scratch,
                                      const uint64_t **d_n_words, cudaStream_t st)
{
  const uint64_t nc = fasta_chunks (n_codes);
  uint64_t *offsets = reinterpret_cast<uint64_t *> (scratch);
  uint32_t *counts = reinterpret_cast<uint32_t *> (offsets + (nc + 1));
  *d_n_words = offsets + nc;
  if (nc == 0) return cudaMemsetAsync (offsets, 0, sizeof (uint64_t), st);
  if (nc > 0x7fffffffull) return cudaErrorInvalidConfiguration;
  fasta_words_kernel<MODE_COUNT><<<(unsigned) nc, FA_NT, 0, st>>> (codes, n_codes, k, counts, nullptr, nullptr);
  offsets_kernel<<<1, 1024, 0, st>>> (counts, nc, offsets);
  return cudaGetLastError ();
}

cudaError_t launch_fasta_words (const uint8_t *codes, uint64_t n_codes, unsigned k, const unsigned char *scratch, uint64_t *words,
                                cudaStream_t st)
{
  const uint64_t nc = fasta_chunks (n_codes);
  if (nc == 0) return cudaSuccess;
  const uint64_t *offsets = reinterpret_cast<const uint64_t *> (scratch);
  fasta_words_kernel<MODE_EMIT><<<(unsigned) nc, FA_NT, 0, st>>> (codes, n_codes, k, nullptr, offsets, words);
  return cudaGetLastError ();
}

}  // namespace gt4gpu
