// gt4gpu_fasta_core.cuh -- the per-thread logic of the device sequence readers, as __host__ __device__ functions so
// that tests/emulate_reader.cpp can replay the chunk / thread decomposition on the CPU with the very same code
// (gt4gpu_fasta_kernel.cu holds the CUDA glue: loads, block scans, carries, stores).
//
// Reference: fasta_reader_read_nwords, src/fasta.c:88-290.
#pragma once

#include <stdint.h>

#include "gt4gpu_core.cuh"      // GT4_HD

namespace gt4gpu {
namespace reader {

constexpr int BYTES_PER_THREAD = 16;
constexpr uint8_t CODE_BREAK = 4, CODE_SKIP = 5, CODE_NL = 6, CODE_GT = 7;

// c2n of src/fasta.c:62-69 plus the two characters that steer the FastA state machine
GT4_HD uint8_t classify (uint8_t c)
{
  switch (c) {
  case 'A': case 'a': return 0;
  case 'C': case 'c': return 1;
  case 'G': case 'g': return 2;
  case 'T': case 't': case 'U': case 'u': return 3;
  case '\n': return CODE_NL;
  case '>': return CODE_GT;
  default: return c < ' ' ? CODE_SKIP : CODE_BREAK;
  }
}

// Line state of a span of text: does it contain a line end, and is there a '>' after its last line end (anywhere, if it
// has none).  combine (a, b) is the state of the concatenation "a b"; the operator is associative.
struct LineState { uint32_t has_nl, gt; };

GT4_HD LineState combine (LineState a, LineState b)
{
  LineState r;
  r.has_nl = a.has_nl | b.has_nl;
  r.gt = b.has_nl ? b.gt : (a.gt | b.gt);
  return r;
}

// line state of one thread's classified bytes
GT4_HD LineState span_state (const uint8_t *cls, int n)
{
  LineState s = {0u, 0u};
  for (int i = 0; i < n; i++) {
    if (cls[i] == CODE_NL) { s.has_nl = 1u; s.gt = 0u; }
    else if (cls[i] == CODE_GT) s.gt = 1u;
  }
  return s;
}

// FastA: codes one thread keeps of its n classified bytes, given the line state just before them.
// A byte is inside a name iff a '>' lies between the start of its line and itself (:147-190); the end of a name
// restarts the word (:152-156); in sequence state nucleotides and restarts are kept, control characters vanish (:221-269).
GT4_HD int walk_fasta (const uint8_t *cls, int n, LineState before, uint8_t *out)
{
  bool in_name = before.gt != 0u;
  int n_out = 0;
  for (int i = 0; i < n; i++) {
    const uint8_t k = cls[i];
    if (k == CODE_NL) {
      if (in_name) out[n_out++] = CODE_BREAK;
      in_name = false;
    } else if (k == CODE_GT) {
      in_name = true;
    } else if (!in_name && k <= CODE_BREAK) {
      out[n_out++] = k;
    }
  }
  return n_out;
}

// FastQ (:191-217, :272-295): four-line records, so the state is the line number modulo 4.  ch: the thread's n raw bytes,
// line: line number at its first byte, at_line_start: whether that byte opens a line.  Returns the codes kept; *bad is set
// where the reference's reader gives up (line 0 not opened by '@', line 2 not opened by '+').
GT4_HD int walk_fastq (const uint8_t *ch, int n, uint64_t line, bool at_line_start, uint8_t *out, bool *bad)
{
  int n_out = 0;
  for (int i = 0; i < n; i++) {
    const uint8_t c = ch[i];
    const unsigned phase = (unsigned) (line & 3u);
    if (at_line_start) {
      if (phase == 0 && c != '@') *bad = true;       // :284-287
      if (phase == 2 && c != '+') *bad = true;       // :203-206
    }
    at_line_start = false;
    if (c == '\n') {
      if (phase == 1) out[n_out++] = CODE_BREAK;      // the next record starts a new word (its name end resets the reader)
      line++;
      at_line_start = true;
    } else if (phase == 1) {
      const uint8_t k = classify (c);
      if (k <= 3) out[n_out++] = k;
      else if (k == CODE_BREAK || k == CODE_GT) out[n_out++] = CODE_BREAK;   // '>' is an ordinary character in FastQ
    }
  }
  return n_out;
}

// Canonical words that end at code positions [j0, j1) of the compacted code stream; the k - 1 codes before j0 prime the
// window.  A k-mer ends at code j iff codes j-k+1..j are nucleotides; canonical = min (word, reverse complement) (:243).
GT4_HD int window_words (const uint8_t *codes, uint64_t j0, uint64_t j1, unsigned k, uint64_t *out)
{
  const uint64_t mask = k >= 32 ? ~0ull : (1ull << (2 * k)) - 1;
  const unsigned top = 2 * (k - 1);
  uint64_t fw = 0, rc = 0;
  unsigned have = 0;
  int n_out = 0;
  for (uint64_t j = j0 >= k - 1 ? j0 - (k - 1) : 0; j < j1; j++) {
    const uint8_t c = codes[j];
    if (c > 3) { have = 0; fw = rc = 0; continue; }
    fw = ((fw << 2) | c) & mask;
    rc = (rc >> 2) | ((uint64_t) (3u - c) << top);
    if (have < k) have++;
    if (have == k && j >= j0) out[n_out++] = fw < rc ? fw : rc;
  }
  return n_out;
}

}  // namespace reader
}  // namespace gt4gpu
