/*
 * gt4gpu.h -- C ABI of libgt4gpu, the B200-native (sm_100a) engine for the
 * sorted-merge set operations of GenomeTester4 k-mer lists.
 *
 * This is the drop-in boundary for ONE path of the reference: what
 * glistcompare does in compare_wordmaps / union_multi / intersect_multi
 * (/root/reference/src/glistcompare.c:789,500,605) and what glistmaker /
 * glistquery do through gt4_write_union (/root/reference/src/set-operations.c:41).
 * The reference reads its inputs one element at a time through the
 * GT4WordSList iterator vtable (src/word-list-sorted.h:42-57); a GPU cannot be
 * fed that way, so this library replaces the containers (GT4WordMap,
 * GT4WordListStream) and the merge functions TOGETHER: gt4gpu_list_* are the
 * containers (SoA u64 words / u32 counts in HBM), gt4gpu_compare2 & friends are
 * the merges.  Each entry point cites the reference interface it replaces.
 *
 * Conventions (same as the reference):
 *   - every function returns 0 on success and a small positive int on error;
 *     a message is kept per thread (gt4gpu_last_error) and nothing aborts;
 *   - lists are immutable and borrowed by the merge calls; results are owned
 *     by the library until gt4gpu_result_free;
 *   - plain pointers and sizes only; no C++ or torch types cross this boundary;
 *   - there is NO CPU fallback: without a CUDA device every compute call fails
 *     with GT4GPU_ERR_CUDA.
 */
#ifndef GT4GPU_H
#define GT4GPU_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GT4GPU_VERSION_MAJOR 4   /* list format generation handled (src/version.h:27-30) */
#define GT4GPU_VERSION_MINOR 2
#define GT4GPU_VERSION_MICRO 16

/* error codes */
enum {
  GT4GPU_OK = 0,
  GT4GPU_ERR_ARG = 1,        /* bad argument / rule rejected (the reference returns 1 too) */
  GT4GPU_ERR_IO = 2,
  GT4GPU_ERR_FORMAT = 3,     /* not a GT4C list / unsupported version / truncated */
  GT4GPU_ERR_CUDA = 4,       /* no device, launch failure, out of device memory */
  GT4GPU_ERR_CAPACITY = 5    /* caller-provided output buffer too small */
};

/* enum Rules of the reference, src/glistcompare.c:45-54 (same numeric values) */
enum {
  GT4GPU_RULE_DEFAULT = 0,
  GT4GPU_RULE_ADD = 1,
  GT4GPU_RULE_SUBTRACT = 2,
  GT4GPU_RULE_MIN = 3,
  GT4GPU_RULE_MAX = 4,
  GT4GPU_RULE_FIRST = 5,
  GT4GPU_RULE_SECOND = 6,
  GT4GPU_RULE_NUMBER = 7
};

/* output streams of the two-list merge; index = position in out[4]
 * (find_union / find_intrsec / find_diff / find_ddiff of compare_wordmaps, :66) */
enum {
  GT4GPU_OP_UNION = 1u,     /* out[0]  <o>_<k>_union.list   */
  GT4GPU_OP_INTRSEC = 2u,   /* out[1]  <o>_<k>_intrsec.list */
  GT4GPU_OP_DIFF = 4u,      /* out[2]  <o>_<k>_0_diff1.list */
  GT4GPU_OP_DDIFF = 8u      /* out[3]  <o>_<k>_0_diff2.list (the caller sets DIFF too: -dd implies -d, :334) */
};

/* GT4ListHeader (struct _GT4ListHeader_4_4, src/word-list.h:61-72): 48 bytes, little endian */
typedef struct gt4gpu_header {
  uint32_t code;            /* 'G'<<24|'T'<<16|'4'<<8|'C', src/word-list.c:31 */
  uint32_t version_major;
  uint32_t version_minor;
  uint32_t word_length;
  uint64_t n_words;
  uint64_t total_count;
  uint64_t list_start;
  uint32_t word_bytes;
  uint32_t count_bytes;
} gt4gpu_header;

/* replaces gt4_list_header_init, src/word-list.c:33-44 */
void gt4gpu_header_init (gt4gpu_header *hdr, uint32_t word_length);

/* opaque container: one sorted k-mer list resident in HBM as SoA arrays */
typedef struct gt4gpu_list gt4gpu_list;

#define GT4GPU_RESULT_CALLER_BUFFERS 1u   /* words/counts/capacity were supplied by the caller (device memory) */
#define GT4GPU_RESULT_COUNT_ONLY 2u       /* nothing materialised, only n_words/total_count */

/* one output stream of a merge */
typedef struct gt4gpu_result {
  uint64_t n_words;         /* header n_words of the output list */
  uint64_t total_count;     /* header total_count (u64 sum of the emitted u32 counts) */
  uint64_t *words;          /* DEVICE pointer, n_words valid entries, ascending */
  uint32_t *counts;         /* DEVICE pointer */
  uint64_t capacity;        /* entries allocated in words/counts */
  uint32_t word_length;     /* k to put in the output header */
  uint32_t flags;
} gt4gpu_result;

/* ---- context -------------------------------------------------------------------------- */

/* Binds the calling process to one CUDA device (one process per GPU) and creates the
 * library stream and memory pool.  device < 0: use the current device. */
/* number of CUDA devices visible to the process (0 without a driver); no context is created */
int gt4gpu_device_count (void);
int gt4gpu_init (int device);
void gt4gpu_shutdown (void);
/* Use an externally owned cudaStream_t (e.g. torch's current stream) for all launches.  NULL restores the library's own
 * non-blocking stream, which is NOT ordered with the default stream: to run on the default stream pass cudaStreamLegacy
 * ((cudaStream_t) 0x1).  Arrays handed in through gt4gpu_list_from_device must be complete on the launch stream. */
int gt4gpu_set_stream (void *cuda_stream);
const char *gt4gpu_last_error (void);
/* Kernel tile shape: threads per CTA and merged items per thread.  Unsupported pairs fail with GT4GPU_ERR_ARG. */
int gt4gpu_set_tile (int threads, int items_per_thread);
/* Tuning knobs: "stream_shape" (consumer threads per CTA * 100 + merged items per thread of the single-output kernel,
 * e.g. 51209; also settable one at a time as "stream_consumers" / "stream_items"),
 * "use_stream_kernel" (0 routes single-output merges through the multi-output tile kernel as well),
 * "use_fused" (0: several outputs take one pass of the single-output kernel each instead of the one-read kernel),
 * "use_kway" (0: N-list calls run as a tree / chain of two-list merges, 1: unions take the single-pass kernel,
 * 2: intersections too) and "stream_side" (0: sparse outputs never take the side-buffer variant of the single-output
 * kernel, 1: when a density sample of the call says so, 2 / 3: always the three-stage / the two-stage variant where it
 * exists).  None of them changes a result. */
int gt4gpu_set_option (const char *name, int value);
/* Device time of the most recent merge call on this thread, from CUDA events on the launch stream:
 * partition kernel, tile kernel, and the launch count (each may be NULL). */
int gt4gpu_last_timing (float *ms_partition, float *ms_merge, uint32_t *n_launches);

/* ---- containers (replace gt4_word_map_new, src/word-map.c:165-241, and
 *      gt4_word_list_stream_new, src/word-list-stream.c:127-186) ------------------------ */

/* Opens a GT4C .list file, applies the reference's header rules (stream_mode 0 = mmap container
 * rules, src/word-map.c:179-215; 1 = --stream rules, src/word-list-stream.c:150-168), copies the
 * 12-byte records to the device and de-interleaves them there into SoA arrays. */
int gt4gpu_list_open (const char *path, int stream_mode, gt4gpu_list **out);
/* Same, but only records [first, first + count) -- the loader of one key-range shard. */
int gt4gpu_list_open_range (const char *path, int stream_mode, uint64_t first, uint64_t count, gt4gpu_list **out);
/* Header of a list file as the reference would see it (no device needed). */
int gt4gpu_list_read_header (const char *path, int stream_mode, gt4gpu_header *out);
/* From host memory: packed 12-byte records (the file/mmap layout) or SoA arrays. */
int gt4gpu_list_from_host_aos (const void *records, uint64_t n_words, uint32_t word_length, gt4gpu_list **out);
int gt4gpu_list_from_host_soa (const uint64_t *words, const uint32_t *counts, uint64_t n_words, uint32_t word_length, gt4gpu_list **out);
/* Wraps device arrays owned by the caller (borrowed; must outlive the list). */
int gt4gpu_list_from_device (const uint64_t *d_words, const uint32_t *d_counts, uint64_t n_words, uint32_t word_length, gt4gpu_list **out);
void gt4gpu_list_close (gt4gpu_list *list);

/* GT4WordSListInstance fields, src/word-list-sorted.h:50-57 */
uint64_t gt4gpu_list_n_words (const gt4gpu_list *list);
uint32_t gt4gpu_list_word_length (const gt4gpu_list *list);
uint64_t gt4gpu_list_sum_counts (const gt4gpu_list *list);   /* header total_count (0 if built from arrays) */
const uint64_t *gt4gpu_list_device_words (const gt4gpu_list *list);
/* Copies the list's words and counts back to host arrays of gt4gpu_list_n_words entries (what walking the container
 * with get_first_word / get_next_word, src/word-list-sorted.c:59-78, would visit). */
int gt4gpu_list_to_host_soa (const gt4gpu_list *list, uint64_t *words, uint32_t *counts);
const uint32_t *gt4gpu_list_device_counts (const gt4gpu_list *list);

/* ---- merges --------------------------------------------------------------------------- */

/* Replaces compare_wordmaps (src/glistcompare.c:66,789-955).  ops = OR of GT4GPU_OP_*; rule,
 * cutoff, count_override (the global of :82), subtract (-du) as in the reference.  countonly != 0
 * materialises nothing.  `out` must be zero-initialised by the caller, except streams flagged
 * GT4GPU_RESULT_CALLER_BUFFERS whose words/counts/capacity name caller-owned device buffers.
 * Only the requested streams are written. */
int gt4gpu_compare2 (const gt4gpu_list *a, const gt4gpu_list *b, uint32_t ops, int rule, uint32_t cutoff,
                     uint32_t count_override, int subtract, int countonly, gt4gpu_result out[4]);

/* Replace union_multi / intersect_multi (src/glistcompare.c:70-71,500-717): N >= 1 lists, cutoff on
 * the COMBINED count, zero not filtered, rules {default,add,max,number} (+min for intersection),
 * anything else returns GT4GPU_ERR_ARG (=1) like the reference. */
int gt4gpu_union_multi (const gt4gpu_list *const *lists, unsigned n_lists, uint32_t cutoff, int rule,
                        uint32_t count_override, int countonly, gt4gpu_result *out);
int gt4gpu_intersect_multi (const gt4gpu_list *const *lists, unsigned n_lists, uint32_t cutoff, int rule,
                            uint32_t count_override, int countonly, gt4gpu_result *out);

/* Replaces gt4_write_union (src/set-operations.h:34, src/set-operations.c:41-129): N-way union
 * with rule add; writes header + records to `ofile` when ofile != 0 ("no actual writing will be
 * done if ofile is 0") and always fills *header.  Re-entrant like the original. */
int gt4gpu_write_union (const gt4gpu_list *const *lists, unsigned n_lists, uint32_t cutoff, int ofile, gt4gpu_header *header);

/* Replace gt4_union / gt4_is_union (src/set-operations.h:38-39): instead of one callback per
 * distinct word, fill a row-major count matrix on the host: words[r], counts[r * n_lists + j]
 * (0 = absent).  is_union != 0 restricts rows to the words of list 0.  *n_rows is always the
 * full row count; at most max_rows rows are stored.  Lists must be non-empty (the reference is
 * undefined otherwise). */
int gt4gpu_union_matrix (const gt4gpu_list *const *lists, unsigned n_lists, int is_union,
                         uint64_t *words, uint32_t *counts, uint64_t max_rows, uint64_t *n_rows);

/* ---- lookups (the step after the merge) ------------------ */

/* Replaces word_map_lookup (binary search, src/word-map.c:134-163) for a batch of words, the way glistquery's
 * search_one_word drives it without mismatches (src/glistquery.c:544-568): when canonize is non-zero every query is
 * first replaced by the smaller of itself and its reverse complement (get_reverse_complement, src/sequence.c:65-79)
 * and that word is stored in canonical_out (may be NULL); counts_out[i] = the list's count of the word, 0 when the list
 * does not hold it (the reference prints "<word>\t0" then).  queries / outputs are all HOST (on_device = 0) or all
 * DEVICE arrays.  The list-against-list zipper (search_list_zipper, src/glistquery.c:702-717) needs no entry point of
 * its own: it is gt4gpu_compare2 (query_list, list, GT4GPU_OP_INTRSEC, GT4GPU_RULE_FIRST, cutoff 0), and
 * search_lists_multi (:776-812) is gt4gpu_union_matrix with is_union = 1. */
int gt4gpu_lookup (const gt4gpu_list *list, const uint64_t *queries, uint64_t n_queries, int on_device, int canonize,
                   uint32_t *counts_out, uint64_t *canonical_out);

/* ---- list building (the step before the merge) ----------- */

/* Replaces fasta_reader_read_nwords (src/fasta.c:88-290) as glistmaker drives it (read_table, src/glistmaker.c:922;
 * canonize = 1, src/listmaker-queue.c:196): HOST-side parse of a FastA / FastQ image into the canonical word of every
 * k-mer, in file order.  words may be NULL to only count (size the buffer with a first call, or use n_bytes as the
 * bound).  GT4GPU_ERR_FORMAT where the reader gives up (bad start tag, FastQ '+' / '@' missing): *n_words then holds
 * the words read so far, as glistmaker keeps them.  No device work. */
int gt4gpu_sequence_words (const void *text, uint64_t n_bytes, uint32_t word_length, uint64_t *words, uint64_t capacity,
                           uint64_t *n_words);
/* Device form of gt4gpu_sequence_words: text is a HOST buffer holding a FastA or FastQ file image; the canonical words
 * come back as a library-owned DEVICE array in file order (release it with gt4gpu_device_free), ready for
 * gt4gpu_count_words (..., on_device = 1, ...).  FastA: same acceptance rules as the host reader.  FastQ: well-formed
 * four-line records only; an image the reference's reader would give up on returns GT4GPU_ERR_FORMAT and no words --
 * gt4gpu_sequence_words then yields the words up to that point, as glistmaker keeps them.  *d_words is NULL when the
 * image holds no k-mer. */
int gt4gpu_fasta_words_device (const void *text, uint64_t n_bytes, uint32_t word_length, uint64_t **d_words, uint64_t *n_words);
void gt4gpu_device_free (void *d_ptr);

/* Replaces the back end of glistmaker for one table of raw words: wordtable_sort (src/word-table.c, radix sort of
 * src/utils.c:127-198, called from read_table, src/glistmaker.c:893-968) followed by the run-length counting of
 * merge_tables_to_file (src/glistmaker.c:1080-1144).  words: n_words canonical k-mer words (< 4^word_length) in any
 * order, duplicates included, in HOST (on_device = 0) or DEVICE memory; the input is not modified.  out: the distinct
 * words ascending with their number of occurrences (device arrays, library-owned), total_count = n_words.  Several
 * tables are collated with gt4gpu_union_multi / gt4gpu_write_union (rule add, cutoff 1) like collate_files (:787-835).
 * gt4gpu_last_timing afterwards reports the sort as ms_partition and the run-length pass as ms_merge. */
int gt4gpu_count_words (const uint64_t *words, uint64_t n_words, int on_device, uint32_t word_length, gt4gpu_result *out);

/* ---- results -------------------------------------------------------------------------- */

int gt4gpu_result_to_host_soa (const gt4gpu_result *res, uint64_t *words, uint32_t *counts);
/* packed 12-byte records, interleaved on the device (what write_word_to_file emits, :491-496) */
int gt4gpu_result_to_host_aos (const gt4gpu_result *res, void *records);
/* header (n_words, total_count filled) + records to fd at offset 0, like compare_wordmaps' fwrite path */
int gt4gpu_write_list (const gt4gpu_result *res, int fd);
/* only this result's records, at byte offset 48 + 12 * first_record (one rank's slice of a sharded output) */
int gt4gpu_write_records_at (const gt4gpu_result *res, int fd, uint64_t first_record);
void gt4gpu_result_free (gt4gpu_result *res);

/* ---- host-to-host convenience (the end-to-end path: H2D, merge, D2H inside) ----------- */

/* Two lists given as packed 12-byte records in host memory (pinned or pageable); requested
 * outputs are returned as packed records in caller-provided host buffers out_records[k] holding
 * out_capacity[k] records (ignored when countonly).  n_out/total_out are always filled for the
 * requested streams. */
int gt4gpu_compare2_host_aos (const void *records_a, uint64_t n_a, const void *records_b, uint64_t n_b,
                              uint32_t word_length, uint32_t ops, int rule, uint32_t cutoff,
                              uint32_t count_override, int subtract, int countonly,
                              void *const out_records[4], const uint64_t out_capacity[4],
                              uint64_t n_out[4], uint64_t total_out[4]);

/* File to file: what glistcompare does with two list files (compare_wordmaps over two GT4WordMap / GT4WordListStream
 * containers, src/glistcompare.c:256-291, 789-955), as a pipeline: the key space is cut into parts, part p + 1 is read
 * out of the page cache and copied to the device while part p is merged and part p - 1 is written with pwrite.  out_fd[k]
 * (k = union, intersection, diff1, diff2; only the requested ones are used, ignored when countonly) receives a complete
 * list file: records at offset 48, the header with the totals last.  List files only (no GT4I index inputs).
 * n_out / total_out are always filled; *word_length (may be NULL) is the first list's. */
int gt4gpu_compare2_files (const char *path_a, const char *path_b, int stream_mode, uint32_t ops, int rule, uint32_t cutoff,
                           uint32_t count_override, int subtract, int countonly, const int out_fd[4],
                           uint64_t n_out[4], uint64_t total_out[4], uint32_t *word_length);

/* ---- key-range sharding (multi-GPU) ---------------------------- */

/* Chooses n_parts - 1 splitter KEY VALUES so that every part holds about the same number of
 * input records summed over all lists, and returns for every list the record index of every part
 * boundary: bounds[j * (n_parts + 1) + p], with bounds[..0] = 0 and bounds[..n_parts] = n_j.
 * Equal keys always land in the same part (lower_bound in every list).  Lists are host arrays of
 * keys with a byte stride (12 for the packed file layout, 8 for SoA). */
int gt4gpu_plan_splitters (const void *const *keys, const size_t *stride_bytes, const uint64_t *n_words,
                           unsigned n_lists, unsigned n_parts, uint64_t *bounds, uint64_t *splitters);

/* ---- device-side helpers exposed for harnesses ------------------------------------------ */

/* AoS <-> SoA on device buffers (d_records: packed 12-byte records). */
int gt4gpu_deinterleave (const void *d_records, uint64_t n, uint64_t *d_words, uint32_t *d_counts);
int gt4gpu_interleave (const uint64_t *d_words, const uint32_t *d_counts, uint64_t n, void *d_records);

#ifdef __cplusplus
}
#endif

#endif /* GT4GPU_H */
