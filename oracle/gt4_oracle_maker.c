/*
 * gt4_oracle_maker.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Plain-C restatement of the list-building path of glistmaker 4.2.16:
 *
 *   FastA/FastQ text -> canonical words     fasta_reader_read_nwords, src/fasta.c:88-290
 *   table of words   -> sorted (word,count)  wordtable_sort (src/word-table.c; radix sort of
 *                                            src/utils.c:127-198) + merge_tables_to_file,
 *                                            src/glistmaker.c:1080-1144
 *
 * Parity status: PINNED -- tests/test_oracle_vs_reference.py and tests/test_listmaker_host.py run the unmodified
 * glistmaker binary (oracle/_ref) on synthetic FastA/FastQ files and compare the list files byte for byte with
 * what these functions produce; the inputs and expected lists are committed under tests/golden/maker/.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* c2n of src/fasta.c:62-69: A/a 0, C/c 1, G/g 2, T/t/U/u 3, everything else "not a nucleotide" */
static unsigned nucl_value (int c)
{
  switch (c) {
  case 'A': case 'a': return 0;
  case 'C': case 'c': return 1;
  case 'G': case 'g': return 2;
  case 'T': case 't': case 'U': case 'u': return 3;
  default: return 0xffffffffu;
  }
}

enum { ST_NONE, ST_NAME, ST_SEQUENCE, ST_QUALITY };

/*
 * Words of one sequence file image, in file order (duplicates included), canonical (the smaller of the word and
 * its reverse complement, src/fasta.c:243).  Returns 0, or -1 where the reader reports a format error
 * (invalid start tag :136-139, missing '+' :203-206 or '@' :284-287).  A zero byte ends the input like EOF
 * (:107-118: the source returns 0 at end of data).  out may be NULL to only count.
 */
int gt4o_sequence_words (const unsigned char *text, uint64_t n_bytes, unsigned word_length, uint64_t *out, uint64_t capacity,
                         uint64_t *n_words)
{
  const uint64_t mask = (word_length >= 32) ? ~0ull : ((1ull << (2 * word_length)) - 1);   /* create_mask */
  int state = ST_NONE, fastq = 0;
  uint64_t fw = 0, rv = 0, n = 0, i = 0;
  unsigned cur = 0;
  *n_words = 0;
  while (i < n_bytes) {
    int c = text[i++];
    if (c == 0) break;
    switch (state) {
    case ST_NONE:                                              /* :131-146 */
      if (c == '>') fastq = 0;
      else if (c == '@') fastq = 1;
      else return -1;
      state = ST_NAME;
      break;
    case ST_NAME:                                              /* :147-175 */
      if (c == '\n') {
        state = ST_SEQUENCE;
        fw = rv = 0;
        cur = 0;
      }
      break;
    case ST_SEQUENCE:
      if (!fastq && c == '>') {                                /* :177-190 */
        state = ST_NAME;
      } else if (fastq && c == '\n') {                         /* :191-217: "+...\n" then quality */
        if (i >= n_bytes || text[i] != '+') return -1;
        i++;
        for (;;) {
          if (i >= n_bytes || text[i] == 0) { *n_words = n; return -1; }   /* cval <= 0 inside the '+' line, :210-214 */
          if (text[i++] == '\n') break;
        }
        state = ST_QUALITY;
      } else {
        const unsigned v = nucl_value (c);
        if (v <= 3) {                                          /* :221-262 */
          fw = (fw << 2) | v;
          rv = (rv >> 2) | ((uint64_t) (~v & 3u) << ((word_length - 1) * 2));
          cur += 1;
          if (cur > word_length) {
            fw &= mask;
            cur = word_length;
          }
          if (cur == word_length) {
            const uint64_t w = fw < rv ? fw : rv;
            if (out) {
              if (n >= capacity) return -2;
              out[n] = w;
            }
            n += 1;
          }
        } else if (c >= ' ') {                                 /* :263-269: any other printable character restarts the word */
          fw = rv = 0;
          cur = 0;
        }                                                      /* control characters (line ends) are skipped */
      }
      break;
    case ST_QUALITY:                                           /* :272-295 */
      if (c == '\n') {
        if (i >= n_bytes || text[i] == 0) { *n_words = n; return 0; }
        if (text[i] != '@') { *n_words = n; return -1; }
        i++;
        state = ST_NAME;
      }
      break;
    }
  }
  *n_words = n;
  return 0;
}

static int cmp_u64 (const void *a, const void *b)
{
  const uint64_t x = *(const uint64_t *) a, y = *(const uint64_t *) b;
  return (x > y) - (x < y);
}

/*
 * One table: sort the words ascending (wordtable_sort) and emit every distinct word with the number of times it
 * occurs (merge_tables_to_file over a single table, src/glistmaker.c:1109-1141: freq += 1 per equal word; freq is an
 * unsigned int).  words is sorted in place.  Returns the number of distinct words; out arrays hold up to n entries.
 */
uint64_t gt4o_count_words (uint64_t *words, uint64_t n, uint64_t *out_words, uint32_t *out_counts)
{
  uint64_t u = 0, i = 0;
  if (n == 0) return 0;
  qsort (words, n, sizeof (uint64_t), cmp_u64);
  while (i < n) {
    uint64_t j = i;
    uint32_t freq = 0;
    while (j < n && words[j] == words[i]) { freq += 1; j += 1; }
    out_words[u] = words[i];
    out_counts[u] = freq;
    u += 1;
    i = j;
  }
  return u;
}
