/*
 * gt4_oracle_query.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Plain-C restatement of the exact-match lookups of glistquery 4.2.16:
 *   get_reverse_complement / canonical word   src/sequence.c:65-86
 *   word_map_lookup (binary search)           src/word-map.c:134-163
 *   search_one_word with n_mm = 0             src/glistquery.c:544-568
 *
 * Parity status: PINNED -- tests/test_query_host.py compares these functions with the text the unmodified
 * glistquery binary (oracle/_ref) prints for "-f queries.txt"; inputs and outputs are committed under
 * tests/golden/query/.
 */
#include <stdint.h>

/* src/sequence.c:65-79 */
uint64_t gt4o_reverse_complement (uint64_t word, unsigned word_length)
{
  uint64_t rc = 0;
  unsigned i;
  word = ~word;
  for (i = 0; i < word_length; i++) {
    rc = (rc << 2) | (word & 3);
    word >>= 2;
  }
  return rc;
}

/* src/word-map.c:145-161: returns 1 and *value when the word is in the list */
static int lookup_one (const uint64_t *words, const uint32_t *counts, uint64_t n, uint64_t word, uint32_t *value)
{
  uint64_t low = 0, high, mid;
  if (n == 0) return 0;
  high = n - 1;
  mid = (low + high) / 2;
  while (low <= high) {
    const uint64_t cur = words[mid];
    if (cur < word) {
      low = mid + 1;
    } else if (cur > word) {
      if (mid == 0) break;
      high = mid - 1;
    } else {
      *value = counts[mid];
      return 1;
    }
    mid = (low + high) / 2;
  }
  return 0;
}

/*
 * Batch form of search_one_word (src/glistquery.c:544-568, n_mm = 0): each query is replaced by its reverse
 * complement when that is smaller (:546-551), looked up, and reported as (canonical word, count) with count 0 for
 * "not in the list".
 */
void gt4o_lookup (const uint64_t *words, const uint32_t *counts, uint64_t n, unsigned word_length, const uint64_t *queries,
                  uint64_t n_queries, int canonize, uint64_t *canonical_out, uint32_t *counts_out)
{
  uint64_t i;
  for (i = 0; i < n_queries; i++) {
    uint64_t w = queries[i];
    uint32_t v = 0;
    if (canonize) {
      const uint64_t r = gt4o_reverse_complement (w, word_length);
      if (r < w) w = r;
    }
    if (!lookup_one (words, counts, n, w, &v)) v = 0;
    if (canonical_out) canonical_out[i] = w;
    counts_out[i] = v;
  }
}
