// emulate_reader.cpp -- TEST HELPER.  Replays the chunk / thread decomposition of the device sequence readers
// (fasta_text_kernel, fastq_text_kernel, fasta_words_kernel in genometester4_b200/csrc/gt4gpu_fasta_kernel.cu) on the
// CPU with the very same __host__ __device__ functions the kernels call (gt4gpu_fasta_core.cuh): classify, span_state,
// combine, walk_fasta, walk_fastq, window_words.  The chunk size is a parameter, so that names, lines and N runs
// straddle thread and chunk boundaries in every possible way.  What it does NOT cover is the CUDA glue (vector loads,
// shuffle scans, single-CTA carry scans, stores) -- that is what the -m gpu tests are for.
// Built by tests/test_reader_emulation.py with g++.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "gt4gpu_fasta_core.cuh"

using namespace gt4gpu::reader;

// returns 0, or 3 (format) for a FastQ image the reference's reader gives up on; words: capacity n_bytes
extern "C" int emu_sequence_words (const uint8_t *text, uint64_t n, unsigned k, int threads_per_chunk, int bytes_per_thread,
                                   uint64_t *words, uint64_t *n_words)
{
  *n_words = 0;
  if (n == 0 || text[0] == 0) return 0;
  const bool fastq = text[0] == '@';
  if (!fastq && text[0] != '>') return 3;
  if (const void *z = memchr (text, 0, n)) n = (uint64_t) ((const uint8_t *) z - text);
  const int bpt = bytes_per_thread;
  if (bpt > BYTES_PER_THREAD) return 1;
  const uint64_t chunk = (uint64_t) threads_per_chunk * bpt;
  const uint64_t n_chunks = (n + chunk - 1) / chunk;
  std::vector<uint8_t> codes;
  bool malformed = false;

  if (!fastq) {
    // pass 1 (MODE_LINES): line state of every chunk, from the per-thread span states scanned in thread order
    std::vector<LineState> lines (n_chunks), carry (n_chunks);
    for (uint64_t c = 0; c < n_chunks; c++) {
      LineState acc = {0u, 0u};
      for (int t = 0; t < threads_per_chunk; t++) {
        uint8_t cls[BYTES_PER_THREAD];
        const uint64_t base = c * chunk + (uint64_t) t * bpt;
        for (int i = 0; i < bpt; i++) cls[i] = (base + i < n) ? classify (text[base + i]) : CODE_SKIP;
        acc = combine (acc, span_state (cls, bpt));
      }
      lines[c] = acc;
    }
    // line_carry_kernel: exclusive scan
    LineState run = {0u, 0u};
    for (uint64_t c = 0; c < n_chunks; c++) { carry[c] = run; run = combine (run, lines[c]); }
    // passes 2 + 3 (MODE_COUNT / MODE_EMIT): every thread walks its bytes from the carried-in state
    for (uint64_t c = 0; c < n_chunks; c++) {
      LineState before = {0u, 0u};
      for (int t = 0; t < threads_per_chunk; t++) {
        uint8_t cls[BYTES_PER_THREAD], out[BYTES_PER_THREAD];
        const uint64_t base = c * chunk + (uint64_t) t * bpt;
        for (int i = 0; i < bpt; i++) cls[i] = (base + i < n) ? classify (text[base + i]) : CODE_SKIP;
        const int m = walk_fasta (cls, bpt, combine (carry[c], before), out);
        codes.insert (codes.end (), out, out + m);
        before = combine (before, span_state (cls, bpt));
      }
    }
  } else {
    // pass 1: line ends per chunk; exclusive sum = line number at the start of every chunk
    std::vector<uint64_t> line0 (n_chunks + 1, 0);
    for (uint64_t c = 0; c < n_chunks; c++) {
      uint64_t nl = 0;
      for (uint64_t i = c * chunk; i < (c + 1) * chunk && i < n; i++) nl += text[i] == '\n';
      line0[c + 1] = line0[c] + nl;
    }
    for (uint64_t c = 0; c < n_chunks; c++) {
      uint64_t line = line0[c];
      for (int t = 0; t < threads_per_chunk; t++) {
        const uint64_t base = c * chunk + (uint64_t) t * bpt;
        if (base >= n) break;
        const int m_in = n - base < (uint64_t) bpt ? (int) (n - base) : bpt;
        uint8_t out[BYTES_PER_THREAD];
        const bool at_line_start = base == 0 || text[base - 1] == '\n';
        const int m = walk_fastq (text + base, m_in, line, at_line_start, out, &malformed);
        codes.insert (codes.end (), out, out + m);
        for (int i = 0; i < m_in; i++) line += text[base + i] == '\n';
      }
    }
    if ((line0[n_chunks] & 3) == 2) malformed = true;      // the image stops on or inside a '+' line
    if (malformed) return 3;
  }

  // fasta_words_kernel: every thread owns bpt consecutive code positions
  const uint64_t n_codes = codes.size ();
  uint64_t n_w = 0;
  for (uint64_t j0 = 0; j0 < n_codes; j0 += bpt) {
    uint64_t out[BYTES_PER_THREAD];
    const uint64_t j1 = j0 + bpt < n_codes ? j0 + bpt : n_codes;
    const int m = window_words (codes.data (), j0, j1, k, out);
    for (int i = 0; i < m; i++) words[n_w++] = out[i];
  }
  *n_words = n_w;
  return 0;
}
