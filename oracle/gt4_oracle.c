/*
 * gt4_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Plain-C restatement of the reference's sorted-merge set operations over SoA
 * arrays.  See gt4_oracle.h for the usage contract and parity status (pinned
 * against the unmodified reference binaries by tests/test_oracle_vs_reference.py
 * and the committed tests/golden/ fixtures).
 *
 * The reference walks its inputs through an iterator interface whose
 * end-of-list behaviour matters (the last word/count stay in the instance when
 * the list is exhausted, src/word-list-sorted.c:69-78).  The `cursor` type below
 * reproduces exactly that contract so the loops can be restated one-to-one.
 */
#include <string.h>

#include "gt4_oracle.h"

/* ---- iterator contract: src/word-list-sorted.h:50-57, src/word-list-sorted.c:59-78 ---- */

typedef struct {
  const gt4o_list *list;
  uint64_t idx;
  uint64_t word;   /* stale after the end, like GT4WordSListInstance.word */
  uint32_t count;
} cursor;

static void
cursor_bind (cursor *c, const gt4o_list *l)
{
  /* Containers preload element 0 at construction (src/word-map.c:222-225,
   * src/word-list-stream.c:174-183); an empty list leaves word/count zeroed. */
  c->list = l;
  c->idx = 0;
  c->word = 0;
  c->count = 0;
  if (l->n_words) {
    c->word = l->words[0];
    c->count = l->counts[0];
  }
}

/* gt4_word_slist_get_first_word, src/word-list-sorted.c:59-67 */
static int
cursor_first (cursor *c)
{
  c->idx = 0;
  if (!c->list->n_words) return 0;
  c->word = c->list->words[0];
  c->count = c->list->counts[0];
  return 1;
}

/* gt4_word_slist_get_next_word, src/word-list-sorted.c:69-78 */
static int
cursor_next (cursor *c)
{
  if (c->idx >= c->list->n_words) return 0;
  c->idx += 1;
  if (c->idx >= c->list->n_words) return 0;
  c->word = c->list->words[c->idx];
  c->count = c->list->counts[c->idx];
  return 1;
}

static int
cursor_live (const cursor *c)
{
  return c->idx < c->list->n_words;
}

static void
emit (gt4o_out *o, uint64_t word, uint32_t freq)
{
  if (o->words && o->n_words < o->capacity) {
    o->words[o->n_words] = word;
    o->counts[o->n_words] = freq;
  }
  o->n_words += 1;
  o->total_count += freq;
}

/* ---- header: src/word-list.c:31-44, src/word-list.h:40-72 ---- */

#define GT4O_LIST_CODE ((uint32_t) ('G' << 24 | 'T' << 16 | '4' << 8 | 'C'))

void
gt4o_header_init (gt4o_header *h, uint32_t word_length)
{
  memset (h, 0, sizeof (*h));
  h->code = GT4O_LIST_CODE;
  h->version_major = 4;   /* src/version.h:27-30 */
  h->version_minor = 2;
  h->word_length = word_length;
  h->list_start = sizeof (gt4o_header);
  h->word_bytes = 8;
  h->count_bytes = 4;
}

int
gt4o_header_parse (const unsigned char *file, uint64_t file_size, int mode, gt4o_header *out)
{
  gt4o_header h;
  memset (&h, 0, sizeof (h));
  if (mode == 0) {
    /* mmap container, src/word-map.c:179-215 */
    uint32_t code, major, minor;
    if (file_size < 12) return 4;
    memcpy (&code, file, 4);
    memcpy (&major, file + 4, 4);
    memcpy (&minor, file + 8, 4);
    if (code != GT4O_LIST_CODE) return 1;
    if (major != 4) return 2;
    if (minor == 0) {
      if (file_size < 40) return 4;
      memcpy (&h, file, 40);
      h.list_start = 40;
      h.word_bytes = 8;
      h.count_bytes = 4;
    } else if (minor <= 2) {
      if (file_size < 40) return 4;
      memcpy (&h, file, 40);
      h.word_bytes = 8;
      h.count_bytes = 4;
    } else {
      if (file_size < 48) return 4;
      memcpy (&h, file, 48);
    }
    if (file_size < h.list_start + h.n_words * (uint64_t) (h.word_bytes + h.count_bytes)) return 3;
  } else {
    /* stream container, src/word-list-stream.c:150-168 */
    if (file_size < 48) return 4;
    memcpy (&h, file, 48);
    if (h.code != GT4O_LIST_CODE) return 1;
    if (h.version_major > 4) return 2;
    if (h.version_major == 4 && h.version_minor == 0) h.list_start = 48;
  }
  *out = h;
  return 0;
}

/* ---- count rules and the three predicates: src/glistcompare.c:433-489 ---- */

uint32_t
gt4o_calculate_freq (uint32_t f1, uint32_t f2, int rule, uint32_t count_override)
{
  switch (rule) {
  case GT4O_RULE_ADD:      return f1 + f2;                 /* u32 wrap-around */
  case GT4O_RULE_SUBTRACT: return (f1 > f2) ? f1 - f2 : 0;
  case GT4O_RULE_MIN:      return (f1 < f2) ? f1 : f2;
  case GT4O_RULE_MAX:      return (f1 > f2) ? f1 : f2;
  case GT4O_RULE_FIRST:    return f1;
  case GT4O_RULE_SECOND:   return f2;
  case GT4O_RULE_NUMBER:   return count_override;
  default:                 return 0;
  }
}

/* include_in_union, :459-466 */
static int
keep_union (uint32_t f1, uint32_t f2, uint32_t *f, int rule, uint32_t cutoff, uint32_t ov)
{
  if (f1 < cutoff && f2 < cutoff) return 0;
  if (rule == GT4O_RULE_DEFAULT) rule = GT4O_RULE_ADD;
  *f = gt4o_calculate_freq (f1, f2, rule, ov);
  return *f != 0;
}

/* include_in_intersection, :468-475 */
static int
keep_intersection (uint32_t f1, uint32_t f2, uint32_t *f, int rule, uint32_t cutoff, uint32_t ov)
{
  if (f1 < cutoff || f2 < cutoff) return 0;
  if (rule == GT4O_RULE_DEFAULT) rule = GT4O_RULE_MIN;
  *f = gt4o_calculate_freq (f1, f2, rule, ov);
  return *f != 0;
}

/* include_in_complement, :477-489 */
static int
keep_complement (uint32_t f1, uint32_t f2, uint32_t *f, int rule, uint32_t cutoff, int subtract, uint32_t ov)
{
  if (subtract) {
    if (f1 != f2 || f1 < cutoff) return 0;
    *f = f1;
    return 1;
  }
  if (f1 < cutoff || f2 >= cutoff) return 0;
  if (rule == GT4O_RULE_DEFAULT) rule = GT4O_RULE_SUBTRACT;
  *f = gt4o_calculate_freq (f1, f2, rule, ov);
  return *f != 0;
}

/* ---- two-list merge: src/glistcompare.c:789-955, loop :843-905 ---- */

int
gt4o_compare2 (const gt4o_list *a, const gt4o_list *b,
               int find_union, int find_intrsec, int find_diff, int find_ddiff,
               int subtract, uint32_t cutoff, int rule, uint32_t ov,
               gt4o_out out[4])
{
  cursor c1, c2;
  unsigned k;
  for (k = 0; k < 4; k++) {
    int wanted = (k == 0) ? find_union : (k == 1) ? find_intrsec : (k == 2) ? find_diff : find_ddiff;
    if (wanted) {
      out[k].n_words = 0;
      out[k].total_count = 0;
    }
  }
  cursor_bind (&c1, a);
  cursor_bind (&c2, b);
  cursor_first (&c1);
  cursor_first (&c2);

  while (cursor_live (&c1) || cursor_live (&c2)) {
    uint32_t f = 0;
    if (cursor_live (&c1) && cursor_live (&c2) && c1.word == c2.word) {
      /* key in both lists (:844-872) */
      if (find_union && keep_union (c1.count, c2.count, &f, rule, cutoff, ov)) emit (&out[0], c1.word, f);
      if (find_intrsec && keep_intersection (c1.count, c2.count, &f, rule, cutoff, ov)) emit (&out[1], c1.word, f);
      if (find_diff && keep_complement (c1.count, c2.count, &f, rule, cutoff, subtract, ov)) emit (&out[2], c1.word, f);
      if (find_ddiff && keep_complement (c2.count, c1.count, &f, rule, cutoff, 0, ov)) emit (&out[3], c2.word, f);
      cursor_next (&c1);
      cursor_next (&c2);
    } else if (cursor_live (&c1) && (!cursor_live (&c2) || c1.word < c2.word)) {
      /* key only in list 1 (:873-888): the other side contributes count 0 */
      if (find_union && keep_union (c1.count, 0, &f, rule, cutoff, ov)) emit (&out[0], c1.word, f);
      if (find_diff && keep_complement (c1.count, 0, &f, rule, cutoff, subtract, ov)) emit (&out[2], c1.word, f);
      cursor_next (&c1);
    } else if (cursor_live (&c2) && (!cursor_live (&c1) || c2.word < c1.word)) {
      /* key only in list 2 (:889-904) */
      if (find_union && keep_union (0, c2.count, &f, rule, cutoff, ov)) emit (&out[0], c2.word, f);
      if (find_ddiff && keep_complement (c2.count, 0, &f, rule, cutoff, 0, ov)) emit (&out[3], c2.word, f);
      cursor_next (&c2);
    }
  }
  return 0;
}

/* ---- N-list union: src/glistcompare.c:500-603 and src/set-operations.c:41-129 ---- */

#define GT4O_MAX_SETS 4096   /* src/set-operations.h:29; glistcompare's MAX_FILES is 1024 (:76) */

static int
union_n (const gt4o_list *lists, unsigned n_lists, uint32_t cutoff, int rule, uint32_t ov,
         gt4o_out *out, uint32_t *word_length_out)
{
  static cursor store[GT4O_MAX_SETS];   /* not re-entrant: test infrastructure only */
  cursor *src[GT4O_MAX_SETS];
  unsigned n_src = 0, j;
  uint64_t word;

  out->n_words = 0;
  out->total_count = 0;
  if (n_lists == 0 || n_lists > GT4O_MAX_SETS) return 1;

  /* keep only non-empty lists, preserving order (:526-533) */
  for (j = 0; j < n_lists; j++) {
    cursor_bind (&store[n_src], &lists[j]);
    src[n_src] = &store[n_src];
    if (lists[j].n_words) {
      cursor_first (src[n_src]);
      n_src += 1;
    }
  }
  /* header word length comes from slot 0 of the compacted array (:535): the
   * first non-empty list, or the last list when all are empty */
  if (word_length_out) *word_length_out = store[0].list->word_length;

  word = 0xffffffffffffffffULL;
  for (j = 0; j < n_src; j++) if (src[j]->word < word) word = src[j]->word;

  while (n_src) {
    uint64_t next = 0xffffffffffffffffULL;
    uint32_t freq = 0;
    j = 0;
    while (j < n_src) {
      if (src[j]->word == word) {
        if (rule == GT4O_RULE_ADD) freq += src[j]->count;
        else if (rule == GT4O_RULE_MAX) { if (src[j]->count > freq) freq = src[j]->count; }
        else freq = ov;
        if (!cursor_next (src[j])) {
          /* exhausted: swap-remove and re-examine the slot (:558-567) */
          n_src -= 1;
          if (n_src > 0) {
            src[j] = src[n_src];
            continue;
          }
          break;
        }
      }
      if (src[j]->word < next) next = src[j]->word;
      j += 1;
    }
    /* zero is NOT filtered here, only the combined count against the cutoff (:574) */
    if (freq >= cutoff) emit (out, word, freq);
    word = next;
  }
  return 0;
}

int
gt4o_union_multi (const gt4o_list *lists, unsigned n_lists, uint32_t cutoff, int rule,
                  uint32_t ov, gt4o_out *out, uint32_t *word_length_out)
{
  /* allowed rules, :518-523 */
  if (rule == GT4O_RULE_DEFAULT) rule = GT4O_RULE_ADD;
  else if (rule != GT4O_RULE_ADD && rule != GT4O_RULE_MAX && rule != GT4O_RULE_NUMBER) return 1;
  return union_n (lists, n_lists, cutoff, rule, ov, out, word_length_out);
}

int
gt4o_write_union (const gt4o_list *lists, unsigned n_lists, uint32_t cutoff,
                  gt4o_out *out, uint32_t *word_length_out)
{
  /* same loop with freq += count only (src/set-operations.c:77-116) */
  return union_n (lists, n_lists, cutoff, GT4O_RULE_ADD, 0, out, word_length_out);
}

/* ---- N-list intersection: src/glistcompare.c:605-717 ---- */

int
gt4o_intersect_multi (const gt4o_list *lists, unsigned n_lists, uint32_t cutoff, int rule,
                      uint32_t ov, gt4o_out *out, uint32_t *word_length_out)
{
  static cursor cur[GT4O_MAX_SETS];
  unsigned j;
  int finished = 0;
  uint64_t word = 0;

  out->n_words = 0;
  out->total_count = 0;
  if (n_lists == 0 || n_lists > GT4O_MAX_SETS) return 1;
  /* allowed rules, :622-627 */
  if (rule == GT4O_RULE_DEFAULT) rule = GT4O_RULE_MIN;
  else if (rule != GT4O_RULE_ADD && rule != GT4O_RULE_MIN && rule != GT4O_RULE_MAX && rule != GT4O_RULE_NUMBER) return 1;

  for (j = 0; j < n_lists; j++) {
    cursor_bind (&cur[j], &lists[j]);
    if (lists[j].n_words) cursor_first (&cur[j]);
    else { finished = 1; break; }   /* any empty list => empty result (:631-636) */
  }
  if (word_length_out) *word_length_out = lists[0].word_length;   /* :639 */

  while (!finished) {
    uint32_t freq = 0;
    unsigned n_equal = 0;
    /* candidate = largest current word (:651-653) */
    for (j = 0; j < n_lists; j++) if (cur[j].word > word) word = cur[j].word;
    /* advance every list to >= candidate, restarting on overshoot (:655-680) */
    for (j = 0; j < n_lists; j++) {
      while (cur[j].word < word) {
        if (!cursor_next (&cur[j])) { finished = 1; break; }
      }
      if (finished) break;
      if (cur[j].word > word) {
        word = cur[j].word;
        break;
      }
      n_equal += 1;
      if (rule == GT4O_RULE_MIN) { if (!freq || cur[j].count < freq) freq = cur[j].count; }
      else if (rule == GT4O_RULE_MAX) { if (cur[j].count > freq) freq = cur[j].count; }
      else if (rule == GT4O_RULE_ADD) freq += cur[j].count;
      else freq = ov;
    }
    if (n_equal == n_lists) {
      if (freq >= cutoff) emit (out, word, freq);   /* :682 */
      /* step all lists; the first one to run out ends the whole loop (:697-704) */
      for (j = 0; j < n_lists; j++) {
        if (!cursor_next (&cur[j])) { finished = 1; break; }
        if (cur[j].word > word) word = cur[j].word;
      }
    }
  }
  return 0;
}

/* ---- callback variants as matrices: src/set-operations.c:132-228 ---- */

int
gt4o_union_matrix (const gt4o_list *lists, unsigned n_lists,
                   uint64_t *words, uint32_t *counts, uint64_t max_rows, uint64_t *n_rows)
{
  static cursor cur[GT4O_MAX_SETS];
  unsigned j, n_src = 0;
  uint64_t word, rows = 0;
  *n_rows = 0;
  if (n_lists == 0 || n_lists > GT4O_MAX_SETS) return 1;
  for (j = 0; j < n_lists; j++) {
    if (!lists[j].n_words) return 2;   /* undefined in the reference (reads past compacted insts[]) */
    cursor_bind (&cur[j], &lists[j]);
    cursor_first (&cur[j]);
    n_src += 1;
  }
  word = 0xffffffffffffffffULL;
  for (j = 0; j < n_lists; j++) if (cursor_live (&cur[j]) && cur[j].word < word) word = cur[j].word;
  while (n_src) {
    uint64_t next = 0xffffffffffffffffULL;
    for (j = 0; j < n_lists; j++) {
      uint32_t c = 0;
      if (cursor_live (&cur[j])) {
        if (cur[j].word == word) {
          c = cur[j].count;
          if (!cursor_next (&cur[j])) n_src -= 1;
        }
        /* the stale last word of a just-exhausted list still feeds `next`
         * (:168-172), which is what produces the reference's extra all-zero row */
        if (cur[j].word < next) next = cur[j].word;
      }
      if (rows < max_rows) counts[rows * n_lists + j] = c;
    }
    if (rows < max_rows) words[rows] = word;
    rows += 1;
    word = next;
  }
  *n_rows = rows;
  return 0;
}

int
gt4o_is_union_matrix (const gt4o_list *lists, unsigned n_lists,
                      uint64_t *words, uint32_t *counts, uint64_t max_rows, uint64_t *n_rows)
{
  static cursor cur[GT4O_MAX_SETS];
  unsigned j;
  uint64_t rows = 0;
  *n_rows = 0;
  if (n_lists == 0 || n_lists > GT4O_MAX_SETS) return 1;
  for (j = 0; j < n_lists; j++) {
    if (!lists[j].n_words) return 2;
    cursor_bind (&cur[j], &lists[j]);
    cursor_first (&cur[j]);
  }
  while (cursor_live (&cur[0])) {
    uint64_t word = cur[0].word;
    if (rows < max_rows) {
      words[rows] = word;
      counts[rows * n_lists] = cur[0].count;
    }
    for (j = 1; j < n_lists; j++) {
      uint32_t c = 0;
      while (cursor_live (&cur[j]) && cur[j].word < word) cursor_next (&cur[j]);
      if (cursor_live (&cur[j]) && cur[j].word == word) c = cur[j].count;
      if (rows < max_rows) counts[rows * n_lists + j] = c;
    }
    rows += 1;
    cursor_next (&cur[0]);
  }
  *n_rows = rows;
  return 0;
}
