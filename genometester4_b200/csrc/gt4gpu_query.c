/*
 * gt4gpu-query -- drop-in for the exact-match side of GenomeTester4's glistquery, with the lookups, the list-against-
 * list zipper and the multi-list dumps running on a B200 through libgt4gpu (include/gt4gpu.h).
 *
 * Flag grammar, messages, stdout text and exit codes follow main() of /root/reference/src/glistquery.c:108-440 (help
 * text :932-960).  What is carried over:
 *
 *   LIST                     dump "KMER\tcount" (print_full_map :481-494); several lists: the count matrix of
 *                            gt4_union / gt4_is_union (--is_union), optional --header (:371-389, dump_lists :95-106)
 *   -stat                    header statistics (get_statistics :814-829)
 *   -q WORD / -f FILE / -s FASTA|FASTQ      exact lookups (search_one_word :544-568): one batch through gt4gpu_lookup
 *   -l LIST                  the zipper (search_list_zipper :702-717) = intersection under rule "first";
 *                            with several lists: search_lists_multi (:776-812) = the is_union count matrix
 *   -min / -max, --3p / --5p, --all (prints found words unfiltered, :552-556)
 *
 * Not carried over (exit 1 + message): -mm / -p above 0, --median, --distribution, --gc, --files, --sequences,
 * --locations.  --bloom, --disable_scouts and -D are accepted and have nothing to steer here.
 */
#include <limits.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <fcntl.h>
#include <unistd.h>

#include "gt4gpu.h"

#define MAX_LISTS 1024

static unsigned int use_3p = 0, use_5p = 0;

static void
print_help (int exit_value)
{
  fprintf (stderr, "glistquery version %u.%u.%u (%s)\n", GT4GPU_VERSION_MAJOR, GT4GPU_VERSION_MINOR, GT4GPU_VERSION_MICRO, "stable");
  fprintf (stderr, "Usage: glistquery INPUT_LIST [OPTIONS]\n");
  fprintf (stderr, "Options:\n");
  fprintf (stderr, "    -v, --version             - print version information and exit\n");
  fprintf (stderr, "    -h, --help                - print this usage screen and exit\n");
  fprintf (stderr, "    -stat, --stats            - print statistics of the list file and exit\n");
  fprintf (stderr, "    --median                  - print min/max/median/average and exit\n");
  fprintf (stderr, "    --distribution MAX        - print distribution up to MAX\n");
  fprintf (stderr, "    --gc                      - print average GC content of all words\n");
  fprintf (stderr, "    -q, --query               - single query word\n");
  fprintf (stderr, "    -f, --queryfile           - list of query words in a file\n");
  fprintf (stderr, "    -s, --seqfile             - FastA/FastQ file\n");
  fprintf (stderr, "    -l, --listfile            - list file made by glistmaker\n");
  fprintf (stderr, "    -mm, --mismatch NUMBER    - specify number of mismatches (0-16; default 0)\n");
  fprintf (stderr, "    -p, --perfectmatch NUMBER - specify number of 3' perfect matches (0-32; default 0)\n");
  fprintf (stderr, "    -min, --minfreq NUMBER    - minimum frequency of the printed words (default 0)\n");
  fprintf (stderr, "    -max, --maxfreq NUMBER    - maximum frequency of the printed words (default MAX_UINT)\n");
  fprintf (stderr, "    --files                   - Print indexed files\n");
  fprintf (stderr, "    --sequences               - Print indexed subsequences\n");
  fprintf (stderr, "    --bloom                   - use bloom filter to speed up lookups\n");
  fprintf (stderr, "    --all                     - in case of mismatches prints all found words\n");
  fprintf (stderr, "    --locations               - in case of index print all word locations\n");
  fprintf (stderr, "    --3p                      - if query is longer than word use 3' end\n");
  fprintf (stderr, "    --5p                      - if query is longer than word use 5' end\n");
  fprintf (stderr, "    -D                        - increase debug level\n");
  exit (exit_value);
}

/* word2string, src/sequence.c:102-114 */
static void
word_text (char *b, uint64_t word, unsigned int k)
{
  for (unsigned int i = 0; i < k; i++) {
    b[k - i - 1] = "ACGT"[word & 3];
    word >>= 2;
  }
  b[k] = 0;
}

/* string_to_word with get_nucl_value's bit trick, src/sequence.c:43-52, :116-130 (also for characters outside ACGTU,
 * which only draw a complaint) */
static uint64_t
text_word (const char *s, unsigned int k)
{
  uint64_t word = 0;
  for (unsigned int i = 0; i < (k < 32 ? k : 32); i++) {
    const char c = s[i];
    if (!c || !strchr ("ACGTUacgtu", c)) fprintf (stderr, "Invalid character %c in string!\n", c);
    word = (word << 2) | (uint64_t) ((c & 4) ? (((c >> 4) | 2) & 3) : ((c & 6) >> 1));
  }
  return word;
}

/* one query string -> word; 0 ok, 1 = the reference's complaint has been printed */
static int
query_word (const char *who, const char *c, unsigned int k, uint64_t *word)
{
  const unsigned int len = (unsigned int) strlen (c);
  if (len == k) *word = text_word (c, k);
  else if (len < k) {
    fprintf (stderr, "%s: Word too short (%u < %u)\n", who, k, len);
    return 1;
  } else if (use_3p) *word = text_word (c + (len - k), k);
  else if (use_5p) *word = text_word (c, k);
  else {
    fprintf (stderr, "%s: Wrong query length (%u != %u) - use --3p or --5p\n", who, k, len);
    return 1;
  }
  return 0;
}

/* the printing rules of search_one_word, src/glistquery.c:552-566 */
static void
print_lookups (const uint64_t *canonical, const uint32_t *counts, uint64_t n, unsigned int k, unsigned int min_freq, unsigned int max_freq,
               int print_all)
{
  char b[40];
  for (uint64_t i = 0; i < n; i++) {
    if (counts[i] ? (print_all || (counts[i] >= min_freq && counts[i] <= max_freq)) : !min_freq) {
      word_text (b, canonical[i], k);
      fprintf (stdout, "%s\t%u\n", b, counts[i]);
    }
  }
}

static int
gpu_fail (void)
{
  fprintf (stderr, "Error: %s\n", gt4gpu_last_error ());
  return 1;
}

int
main (int argc, const char *argv[])
{
  const char *lists[MAX_LISTS];
  unsigned int n_lists = 0, invalid = 0, nmm = 0, pm3 = 0, minfreq = 0, maxfreq = UINT_MAX, is_union = 0, wlen = 0;
  const char *querystring = NULL, *queryfilename = NULL, *seqfilename = NULL, *querylistfilename = NULL, *unsupported = NULL;
  int printall = 0, print_header = 0, stats = 0, argidx, i;
  char *end;

  for (argidx = 1; argidx < argc; argidx++) {
    const char *a = argv[argidx];
    const char **file_opt = NULL;
    const char *missing = NULL;
    if (!strcmp (a, "-v") || !strcmp (a, "--version")) {
      fprintf (stdout, "glistquery version %u.%u.%u (%s)\n", GT4GPU_VERSION_MAJOR, GT4GPU_VERSION_MINOR, GT4GPU_VERSION_MICRO, "stable");
      return 0;
    }
    if (!strcmp (a, "-h") || !strcmp (a, "--help") || !strcmp (a, "-?")) print_help (0);
    if (!strcmp (a, "-s") || !strcmp (a, "--seqfile")) { file_opt = &seqfilename; missing = "sequence file name"; }
    else if (!strcmp (a, "-l") || !strcmp (a, "--listfile")) { file_opt = &querylistfilename; missing = "query list file name"; }
    else if (!strcmp (a, "-f") || !strcmp (a, "--queryfile")) { file_opt = &queryfilename; missing = "query file name"; }
    else if (!strcmp (a, "-q") || !strcmp (a, "--query")) { file_opt = &querystring; missing = "query"; }
    if (file_opt) {
      if (!argv[argidx + 1] || argv[argidx + 1][0] == '-') fprintf (stderr, "Warning: No %s specified!\n", missing);
      else *file_opt = argv[argidx + 1];
      argidx += 1;
      continue;
    }
    if (!strcmp (a, "-p") || !strcmp (a, "--perfectmatch") || !strcmp (a, "-mm") || !strcmp (a, "--mismatch")) {
      const int is_p = a[1] == 'p' || a[2] == 'p';
      long v;
      if (++argidx >= argc) print_help (1);
      v = strtol (argv[argidx], &end, 10);
      if (*end || v < 0 || v > (is_p ? 32 : 16)) print_help (1);
      if (is_p) pm3 = (unsigned int) v; else nmm = (unsigned int) v;
      continue;
    }
    if (!strcmp (a, "-min") || !strcmp (a, "--minfreq") || !strcmp (a, "-max") || !strcmp (a, "--maxfreq")) {
      const int is_min = a[2] == 'i' || a[3] == 'i';
      if (!argv[argidx + 1]) {
        if (is_min) fprintf (stderr, "Warning: No minimum frequency specified! Using the default value: %d.\n", minfreq);
        else fprintf (stderr, "Warning: No maximum frequency specified! Using the default value: %d.\n", maxfreq);
        argidx += 1;
        continue;
      }
      const unsigned int v = (unsigned int) strtol (argv[argidx + 1], &end, 10);
      if (*end != 0) {
        fprintf (stderr, "Error: Invalid %s frequency: %s! Must be a positive integer.\n", is_min ? "minimum" : "maximum", argv[argidx + 1]);
        print_help (1);
      }
      if (is_min) minfreq = v; else maxfreq = v;
      argidx += 1;
      continue;
    }
    if (!strcmp (a, "-D") || !strcmp (a, "--bloom") || !strcmp (a, "--disable_scouts")) continue;
    if (!strcmp (a, "--all") || !strcmp (a, "-all")) { printall = 1; continue; }
    if (!strcmp (a, "--stats") || !strcmp (a, "--stat") || !strcmp (a, "-stat")) { stats = 1; continue; }
    if (!strcmp (a, "--distribution") || !strcmp (a, "-distribution")) {
      if ((argidx + 1) >= argc) print_help (1);
      argidx += 1;
      unsupported = a;
      continue;
    }
    if (!strcmp (a, "--median") || !strcmp (a, "-median") || !strcmp (a, "-gc") || !strcmp (a, "--gc") || !strcmp (a, "--files") ||
        !strcmp (a, "--sequences") || !strcmp (a, "--locations")) { unsupported = a; continue; }
    if (!strcmp (a, "--3p")) { use_3p = 1; continue; }
    if (!strcmp (a, "--5p")) { use_5p = 1; continue; }
    if (!strcmp (a, "--header")) { print_header = 1; continue; }
    if (!strcmp (a, "--is_union")) { is_union = 1; continue; }
    if (a[0] != '-') {
      if (n_lists < MAX_LISTS) lists[n_lists++] = a;
      continue;
    }
    fprintf (stderr, "Error: Unknown argument: %s!\n", a);
    print_help (1);
  }
  if (!n_lists) {
    fprintf (stderr, "No list/index files specified!\n");
    print_help (1);
  }
  if (unsupported || nmm || pm3) {
    fprintf (stderr, "Error: %s is not supported by the GPU query tool\n", unsupported ? unsupported : "a search with mismatches");
    return 1;
  }

  /* headers first: -stat needs nothing else, and the word lengths must agree (:302-316) */
  gt4gpu_header headers[MAX_LISTS];
  for (i = 0; i < (int) n_lists; i++) {
    FILE *ifs = fopen (lists[i], "r");
    uint32_t code = 0;
    if (!ifs) {
      fprintf (stderr, "Cannot open list %s\n", lists[i]);
      exit (1);
    }
    if (fread (&code, 4, 1, ifs) != 1) code = 0;
    fclose (ifs);
    if (code != 0x47543443u && code != 0x47543449u) {          /* 'GT4C' / 'GT4I' */
      fprintf (stderr, "Error: %s is not a valid GenomeTester4 list/index file\n", lists[i]);
      invalid = 1;
    }
    if (gt4gpu_list_read_header (lists[i], 0, &headers[i])) {
      fprintf (stderr, "Error: %s is invalid or corrupted\n", lists[i]);
      invalid = 1;
      continue;
    }
    if (!wlen) wlen = headers[i].word_length;
    else if (headers[i].word_length != wlen) {
      fprintf (stderr, "Error: %s has different word length %u (first list had %u)\n", lists[i], headers[i].word_length, wlen);
      invalid = 1;
    }
  }
  if (querylistfilename) {
    gt4gpu_header qh;
    if (gt4gpu_list_read_header (querylistfilename, 1, &qh)) {
      fprintf (stderr, "Error: %s is invalid or corrupted\n", querylistfilename);
      invalid = 1;
    } else if (qh.word_length != wlen) {
      fprintf (stderr, "Error: %s has different word length %u (first list had %u)\n", querylistfilename, qh.word_length, wlen);
      invalid = 1;
    }
  }
  if (invalid) exit (1);
  if (stats) {
    for (i = 0; i < (int) n_lists; i++) {
      fprintf (stdout, "List %s: built with glistmaker version %d.%d\n", lists[i], headers[i].version_major, headers[i].version_minor);
      fprintf (stdout, "Wordlength\t%u\n", headers[i].word_length);
      fprintf (stdout, "NUnique\t%llu\n", (unsigned long long) headers[i].n_words);
      fprintf (stdout, "NTotal\t%llu\n", (unsigned long long) headers[i].total_count);
    }
    exit (0);
  }

  const int have_query = seqfilename || querylistfilename || queryfilename || querystring;
  if (have_query && !(querylistfilename && n_lists > 1) && n_lists > 1) {
    fprintf (stderr, "Error: Query is incompatible with multiple lists/indices\n");
    exit (1);
  }

  /* containers on the device */
  static gt4gpu_list *maps[MAX_LISTS + 1];
  const int stream_rules = querylistfilename && n_lists > 1;
  for (i = 0; i < (int) n_lists; i++) {
    if (gt4gpu_list_open (lists[i], stream_rules, &maps[i + 1])) return gpu_fail ();
  }
  char b[40];

  if (!have_query) {
    if (n_lists > 1) {
      /* count matrix of all lists (dump_lists) */
      uint64_t cap = n_lists + 1, n_rows = 0;
      for (i = 0; i < (int) n_lists; i++) cap += headers[i].n_words;
      uint64_t *words = (uint64_t *) malloc (cap * sizeof (uint64_t));
      uint32_t *counts = (uint32_t *) malloc (cap * n_lists * sizeof (uint32_t));
      if (!words || !counts) { fprintf (stderr, "Out of memory\n"); return 1; }
      if (print_header) {
        fprintf (stdout, "KMER");
        for (i = 0; i < (int) n_lists; i++) fprintf (stdout, "\t%s", lists[i]);
        fprintf (stdout, "\n");
      }
      if (gt4gpu_union_matrix ((const gt4gpu_list *const *) (maps + 1), n_lists, (int) is_union, words, counts, cap, &n_rows)) return gpu_fail ();
      for (uint64_t r = 0; r < n_rows; r++) {
        word_text (b, words[r], wlen);
        fprintf (stdout, "%s", b);
        for (i = 0; i < (int) n_lists; i++) fprintf (stdout, "\t%u", counts[r * n_lists + i]);
        fprintf (stdout, "\n");
      }
    } else {
      const uint64_t n = gt4gpu_list_n_words (maps[1]);
      uint64_t *words = (uint64_t *) malloc ((n + 1) * sizeof (uint64_t));
      uint32_t *counts = (uint32_t *) malloc ((n + 1) * sizeof (uint32_t));
      if (!words || !counts) { fprintf (stderr, "Out of memory\n"); return 1; }
      if (gt4gpu_list_to_host_soa (maps[1], words, counts)) return gpu_fail ();
      for (uint64_t r = 0; r < n; r++) {
        word_text (b, words[r], wlen);
        fprintf (stdout, "%s\t%u\n", b, counts[r]);
      }
    }
    exit (0);
  }

  if (querylistfilename) {
    if (gt4gpu_list_open (querylistfilename, 1, &maps[0])) return gpu_fail ();
    if (n_lists > 1) {
      /* search_lists_multi: rows = the query list's words, "\t<list>:<count>" for the lists that hold them */
      const uint64_t n = gt4gpu_list_n_words (maps[0]);
      const unsigned int cols = n_lists + 1;
      uint64_t n_rows = 0;
      uint64_t *words = (uint64_t *) malloc ((n + 2) * sizeof (uint64_t));
      uint32_t *counts = (uint32_t *) malloc ((n + 2) * cols * sizeof (uint32_t));
      if (!words || !counts) { fprintf (stderr, "Out of memory\n"); return 1; }
      if (n && gt4gpu_union_matrix ((const gt4gpu_list *const *) maps, cols, 1, words, counts, n + 2, &n_rows)) return gpu_fail ();
      for (uint64_t r = 0; r < n_rows && r < n; r++) {
        int printed = 0;
        for (unsigned int j = 1; j < cols; j++) {
          if (!counts[r * cols + j]) continue;
          if (!printed) {
            word_text (b, words[r], wlen);
            fprintf (stdout, "%s", b);
            printed = 1;
          }
          fprintf (stdout, "\t%u:%u", j - 1, counts[r * cols + j]);
        }
        if (printed) fprintf (stdout, "\n");
      }
      exit (0);
    }
    /* the zipper: the query list's records for the words the list holds too */
    gt4gpu_result res[4];
    memset (res, 0, sizeof (res));
    if (gt4gpu_compare2 (maps[0], maps[1], GT4GPU_OP_INTRSEC, GT4GPU_RULE_FIRST, 0, 1, 0, 0, res)) return gpu_fail ();
    const uint64_t n = res[1].n_words;
    uint64_t *words = (uint64_t *) malloc ((n + 1) * sizeof (uint64_t));
    uint32_t *counts = (uint32_t *) malloc ((n + 1) * sizeof (uint32_t));
    if (!words || !counts) { fprintf (stderr, "Out of memory\n"); return 1; }
    if (n && gt4gpu_result_to_host_soa (&res[1], words, counts)) return gpu_fail ();
    for (uint64_t r = 0; r < n; r++) {
      word_text (b, words[r], wlen);
      fprintf (stdout, "%s\t%u\n", b, counts[r]);
    }
    return 0;
  }

  /* collect the query words, ask once, print in order */
  uint64_t *queries = NULL, n_q = 0, cap_q = 0;
  int rc_after = 0;      /* status to leave with once the words read so far have been answered */
  if (querystring) {
    queries = (uint64_t *) malloc (sizeof (uint64_t));
    if (query_word ("search_one_query_string", querystring, wlen, &queries[0])) return 1;
    n_q = 1;
  } else if (queryfilename) {
    /* one query per line (search_n_query_strings :610-664): at most 255 characters of a line count, control characters
     * before the next word are skipped */
    FILE *ifs = fopen (queryfilename, "r");
    if (!ifs) {
      fprintf (stderr, "search_n_query_strings: Cannot open file %s.\n", queryfilename);
      return 1;
    }
    int val = fgetc (ifs);
    while (val > 0) {
      char c[256];
      unsigned int len = 0;
      uint64_t w;
      while (val > 0 && len < 255 && val != '\n') {
        c[len++] = (char) val;
        val = fgetc (ifs);
      }
      c[len] = 0;
      while (val > 0 && val != '\n') val = fgetc (ifs);
      while (val > 0 && val < 'A') val = fgetc (ifs);
      if (query_word ("search_n_query_strings", c, wlen, &w)) { rc_after = 1; break; }
      if (n_q == cap_q) {
        cap_q = cap_q ? 2 * cap_q : 1024;
        queries = (uint64_t *) realloc (queries, cap_q * sizeof (uint64_t));
        if (!queries) { fprintf (stderr, "Out of memory\n"); return 1; }
      }
      queries[n_q++] = w;
    }
    fclose (ifs);
  } else {
    struct stat s;
    int fd = open (seqfilename, O_RDONLY);
    if (fd < 0 || fstat (fd, &s)) {
      fprintf (stderr, "search_fasta: Cannot open %s\n", seqfilename);
      return 1;
    }
    if (s.st_size) {
      const void *text = mmap (NULL, s.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
      if (text == MAP_FAILED) { fprintf (stderr, "search_fasta: Cannot open %s\n", seqfilename); return 1; }
      queries = (uint64_t *) malloc ((size_t) s.st_size * sizeof (uint64_t) + 8);
      if (!queries) { fprintf (stderr, "Out of memory\n"); return 1; }
      if (gt4gpu_sequence_words (text, (uint64_t) s.st_size, wlen, queries, (uint64_t) s.st_size, &n_q)) {
        fprintf (stderr, "fasta_reader_read_nwords: Reader %s: %s\n", seqfilename, gt4gpu_last_error ());
        rc_after = 255;      /* the reference returns the reader's -1 as its exit status */
      }
      munmap ((void *) text, s.st_size);
    }
    close (fd);
  }
  if (n_q) {
    uint64_t *canonical = (uint64_t *) malloc (n_q * sizeof (uint64_t));
    uint32_t *counts = (uint32_t *) malloc (n_q * sizeof (uint32_t));
    if (!canonical || !counts) { fprintf (stderr, "Out of memory\n"); return 1; }
    if (gt4gpu_lookup (maps[1], queries, n_q, 0, 1, counts, canonical)) return gpu_fail ();
    print_lookups (canonical, counts, n_q, wlen, minfreq, maxfreq, printall);
  }
  return rc_after;
}
