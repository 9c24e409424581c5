/*
 * gt4_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement, over SoA arrays, of the sorted-merge set operations of
 * GenomeTester4 4.2.16.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (libgt4gpu.so) never links, imports or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py runs the unmodified
 * reference binaries (oracle/_ref, built by oracle/Makefile from
 * /root/reference/src) over the flag matrix and compares byte-for-byte; the
 * outputs of those runs are also committed under tests/golden/ so the check
 * still runs where /root/reference is absent.
 *
 * Every function cites the reference lines it follows (paths relative to
 * /root/reference/).
 */
#ifndef GT4_ORACLE_H
#define GT4_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* enum Rules, src/glistcompare.c:45-54 */
enum {
  GT4O_RULE_DEFAULT = 0,
  GT4O_RULE_ADD = 1,
  GT4O_RULE_SUBTRACT = 2,
  GT4O_RULE_MIN = 3,
  GT4O_RULE_MAX = 4,
  GT4O_RULE_FIRST = 5,
  GT4O_RULE_SECOND = 6,
  GT4O_RULE_NUMBER = 7
};

/* One sorted list: strictly ascending u64 words + u32 counts (SoA view of the
 * 12-byte records of a .list file, src/word-map.h:89-99). */
typedef struct {
  const uint64_t *words;
  const uint32_t *counts;
  uint64_t n_words;
  uint32_t word_length;
} gt4o_list;

/* One output stream.  words/counts may be NULL (count-only); otherwise they
 * must hold `capacity` entries.  n_words/total_count are always filled. */
typedef struct {
  uint64_t *words;
  uint32_t *counts;
  uint64_t capacity;
  uint64_t n_words;
  uint64_t total_count;
} gt4o_out;

/* 48-byte header image, src/word-list.h:61-72 */
typedef struct {
  uint32_t code;
  uint32_t version_major;
  uint32_t version_minor;
  uint32_t word_length;
  uint64_t n_words;
  uint64_t total_count;
  uint64_t list_start;
  uint32_t word_bytes;
  uint32_t count_bytes;
} gt4o_header;

/* src/word-list.c:31-44 */
void gt4o_header_init (gt4o_header *h, uint32_t word_length);

/* Header acceptance rules.  mode 0 = mmap container (src/word-map.c:179-215),
 * mode 1 = stream container (src/word-list-stream.c:150-168).
 * Returns 0 on success, else: 1 bad tag, 2 bad major version, 3 file too small
 * (map only), 4 short header. */
int gt4o_header_parse (const unsigned char *file, uint64_t file_size, int mode, gt4o_header *out);

/* src/glistcompare.c:433-455 */
uint32_t gt4o_calculate_freq (uint32_t f1, uint32_t f2, int rule, uint32_t count_override);

/* src/glistcompare.c:789-955 (merge loop :843-905).  out[0..3] = union,
 * intrsec, diff1, diff2; streams not requested are left untouched.
 * `find_diff` must already include the "-dd implies -d" rule of main (:334). */
int gt4o_compare2 (const gt4o_list *a, const gt4o_list *b,
                   int find_union, int find_intrsec, int find_diff, int find_ddiff,
                   int subtract, uint32_t cutoff, int rule, uint32_t count_override,
                   gt4o_out out[4]);

/* src/glistcompare.c:500-603.  Returns 1 for a rule outside {default,add,max,number}. */
int gt4o_union_multi (const gt4o_list *lists, unsigned n_lists, uint32_t cutoff, int rule,
                      uint32_t count_override, gt4o_out *out, uint32_t *word_length_out);

/* src/glistcompare.c:605-717.  Returns 1 for a rule outside {default,add,min,max,number}. */
int gt4o_intersect_multi (const gt4o_list *lists, unsigned n_lists, uint32_t cutoff, int rule,
                          uint32_t count_override, gt4o_out *out, uint32_t *word_length_out);

/* src/set-operations.c:41-129 (N-way union, rule fixed to add). */
int gt4o_write_union (const gt4o_list *lists, unsigned n_lists, uint32_t cutoff,
                      gt4o_out *out, uint32_t *word_length_out);

/* src/set-operations.c:132-183 / :186-228.  The callback variants, restated as
 * "fill a row-major matrix": row r = word[r], counts[r*n_lists + j].
 * max_rows bounds the buffers; returns the number of rows produced through
 * *n_rows (rows beyond max_rows are counted but not stored).
 * Defined only for all-non-empty inputs (the reference reads uninitialised
 * pointers otherwise); returns 2 if a list is empty. */
int gt4o_union_matrix (const gt4o_list *lists, unsigned n_lists,
                       uint64_t *words, uint32_t *counts, uint64_t max_rows, uint64_t *n_rows);
int gt4o_is_union_matrix (const gt4o_list *lists, unsigned n_lists,
                          uint64_t *words, uint32_t *counts, uint64_t max_rows, uint64_t *n_rows);

#ifdef __cplusplus
}
#endif

#endif
