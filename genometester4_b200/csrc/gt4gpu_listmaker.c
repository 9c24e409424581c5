/*
 * gt4gpu-listmaker -- drop-in for the list mode of GenomeTester4's glistmaker with the sort / count / collate
 * back end running on a B200 through libgt4gpu (include/gt4gpu.h).
 *
 * Flag grammar, validation order, messages, the "<out>_<k>.list" name, tmp + rename and exit codes follow main()
 * of /root/reference/src/glistmaker.c:138-366 (help text :1303-1326).  The pipeline mirrors the reference's tasks:
 *
 *   read_table      (:893-968)    sequence file -> a table of canonical words        gt4gpu_fasta_words_device (GPU)
 *                                                                                     gt4gpu_sequence_words (host: malformed FastQ)
 *   wordtable_sort + merge_tables_to_file (:924, :1080-1144)   table -> (word, count)  gt4gpu_count_words   (GPU)
 *   collate_files / final gt4_write_union (:787-835, :314-333)  tables -> one list     gt4gpu_union_multi   (GPU)
 *
 * Like the reference's list mode, -c/--cutoff/--min and --max are parsed and validated but do not filter the list
 * (they only act on --index, :486).  Not carried over: --index, .gz inputs and "-" (stdin); they exit with status 1
 * and a message.  --num_threads, --max_tables, --tmpdir and --stream are accepted and have nothing to steer here;
 * --table_size is the number of words sorted per GPU table (default 2^30).
 */
#include <errno.h>
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <unistd.h>

#include "gt4gpu.h"

#define MAX_INPUTS 1024          /* glistmaker.c:141 */
#define DEFAULT_CUTOFF 1         /* :49 */
#define DEFAULT_NUM_THREADS 8    /* :50 */
#define DEFAULT_NUM_TABLES (32 * 128)   /* :52 */
#define DEFAULT_TABLE_WORDS (1ULL << 30)

static int debug = 0;

static void
print_help (int exit_value)
{
  fprintf (stderr, "glistmaker version %u.%u.%u (%s)\n", GT4GPU_VERSION_MAJOR, GT4GPU_VERSION_MINOR, GT4GPU_VERSION_MICRO, "stable");
  fprintf (stderr, "Usage: glistmaker <INPUTFILES> [OPTIONS]\n");
  fprintf (stderr, "Options:\n");
  fprintf (stderr, "    -v, --version           - print version information and exit\n");
  fprintf (stderr, "    -h, --help              - print this usage screen and exit\n");
  fprintf (stderr, "    -w, --wordlength NUMBER - specify index wordsize (1-32)\n");
  fprintf (stderr, "    -o, --outputname STRING - specify output name (default \"out\")\n");
  fprintf (stderr, "    --index                 - create index instead of list\n");
  fprintf (stderr, "    --num_threads           - number of threads (default %u)\n", DEFAULT_NUM_THREADS);
  fprintf (stderr, "    --max_tables            - maximum number of temporary tables (default %u)\n", DEFAULT_NUM_TABLES);
  fprintf (stderr, "    --table_size            - maximum size of the temporary table (default %llu)\n", DEFAULT_TABLE_WORDS);
  fprintf (stderr, "    --tmpdir                - directory for temporary files (may need an order of magnitude more space than the size of the final list)\n");
  fprintf (stderr, "    --stream                - read files as streams instead of memory-mapping (slower but uses less virtual memory)\n");
  fprintf (stderr, "    --index                 - creates indexed list (larger and slower)\n");
  fprintf (stderr, "    -D                      - increase debug level\n");
  exit (exit_value);
}

static double
now (void)
{
  struct timeval tv;
  gettimeofday (&tv, NULL);
  return tv.tv_sec + tv.tv_usec * 1e-6;
}

/* the tables sorted so far, resident on the device */
static gt4gpu_result tables[4096];
static gt4gpu_list *table_lists[4096];
static unsigned n_tables = 0;
static double t_read = 0, t_sort = 0, t_collate = 0;
static uint64_t fasta_block = 1ULL << 30;     /* bytes of FastA text parsed per GPU call (GT4GPU_FASTA_BLOCK overrides) */
static uint64_t fastq_device_max = 8ULL << 30;   /* larger FastQ files take the host reader */
static unsigned long long n_read = 0;
static unsigned fold_tables = 16;                /* tables resident before they are folded into one (GT4GPU_FOLD_TABLES) */
static uint64_t fold_bytes = 32ULL << 30;          /* ... or this many bytes of table records (GT4GPU_FOLD_BYTES) */
static uint64_t resident_bytes = 0;

static int
flush_table (const uint64_t *words, uint64_t n, int on_device, unsigned wordlength)
{
  double t0 = now ();
  int rc;
  if (!n) return 0;
  if (n_tables >= fold_tables || resident_bytes >= fold_bytes) {
    /* fold what we have into one table first, so that the device only ever holds a bounded number of tables plus the
     * distinct words seen so far (the reference collates its temporary files 16 at a time, src/glistmaker.c:822-830) */
    gt4gpu_result merged;
    memset (&merged, 0, sizeof (merged));
    rc = gt4gpu_union_multi ((const gt4gpu_list *const *) table_lists, n_tables, 1, GT4GPU_RULE_ADD, 1, 0, &merged);
    if (rc) return rc;
    for (unsigned i = 0; i < n_tables; i++) { gt4gpu_list_close (table_lists[i]); gt4gpu_result_free (&tables[i]); }
    tables[0] = merged;
    rc = gt4gpu_list_from_device (merged.words, merged.counts, merged.n_words, wordlength, &table_lists[0]);
    if (rc) return rc;
    n_tables = 1;
    resident_bytes = merged.n_words * 12ULL;
    /* the folded table itself may exceed the budget: leave room for as many new words again before the next fold */
    if (resident_bytes >= fold_bytes / 2) fold_bytes = resident_bytes * 2;
    if (debug) fprintf (stderr, "Folded tables: %llu unique\n", (unsigned long long) merged.n_words);
  }
  memset (&tables[n_tables], 0, sizeof (tables[0]));
  rc = gt4gpu_count_words (words, n, on_device, wordlength, &tables[n_tables]);
  if (rc) return rc;
  rc = gt4gpu_list_from_device (tables[n_tables].words, tables[n_tables].counts, tables[n_tables].n_words, wordlength, &table_lists[n_tables]);
  if (rc) return rc;
  if (debug) fprintf (stderr, "Table %u: %llu words, %llu unique\n", n_tables, (unsigned long long) n, (unsigned long long) tables[n_tables].n_words);
  resident_bytes += tables[n_tables].n_words * 12ULL;
  n_tables += 1;
  t_sort += now () - t0;
  return 0;
}

int
main (int argc, const char *argv[])
{
  const char *inputs[MAX_INPUTS];
  unsigned int n_inputs = 0, i;
  char *end;
  unsigned int wordlength = 0, nthreads = DEFAULT_NUM_THREADS, ntables = DEFAULT_NUM_TABLES, min = DEFAULT_CUTOFF, max = 0xffffffff;
  unsigned long long tablesize = DEFAULT_TABLE_WORDS;
  const char *outputname = "out";
  int create_index = 0;
  char tmp_name[1024], out_name[1024];

  /* integer options: names, what the complaint calls them, where the value goes */
  long long v_word = 0, v_min = DEFAULT_CUTOFF, v_max = 0xffffffffLL, v_threads = DEFAULT_NUM_THREADS, v_tables = DEFAULT_NUM_TABLES,
            v_tsize = (long long) DEFAULT_TABLE_WORDS;
  const struct {
    const char *name[3];
    const char *what;
    long long *value;
    int swallow_next;      /* --table_size also skips the token after its value (src/glistmaker.c:214) */
  } int_opts[] = {
    {{"-w", "--wordlength", NULL}, "word-length", &v_word, 0},
    {{"-c", "--cutoff", "--min"}, "frequency cut-off", &v_min, 0},
    {{"--max", NULL, NULL}, "frequency cut-off", &v_max, 0},
    {{"--num_threads", NULL, NULL}, "num-threads", &v_threads, 0},
    {{"--max_tables", NULL, NULL}, "max_tables", &v_tables, 0},
    {{"--table_size", NULL, NULL}, "table-size", &v_tsize, 1},
  };
  for (i = 1; i < (unsigned int) argc; i++) {
    const char *a = argv[i];
    unsigned int o, m, hit = 0;
    for (o = 0; o < sizeof (int_opts) / sizeof (int_opts[0]) && !hit; o++) {
      for (m = 0; m < 3 && int_opts[o].name[m]; m++) {
        if (strcmp (a, int_opts[o].name[m])) continue;
        if (++i >= (unsigned int) argc) print_help (1);
        *int_opts[o].value = strtoll (argv[i], &end, 10);
        if (*end != 0) {
          fprintf (stderr, "Error: Invalid %s: %s! Must be an integer.\n", int_opts[o].what, argv[i]);
          print_help (1);
        }
        i += int_opts[o].swallow_next;
        hit = 1;
        break;
      }
    }
    if (hit) continue;
    if (!strcmp (a, "-v") || !strcmp (a, "--version")) {
      fprintf (stdout, "glistmaker version %u.%u.%u (%s)\n", GT4GPU_VERSION_MAJOR, GT4GPU_VERSION_MINOR, GT4GPU_VERSION_MICRO, "stable");
      return 0;
    }
    if (!strcmp (a, "-h") || !strcmp (a, "--help") || !strcmp (a, "-?")) print_help (0);
    if (!strcmp (a, "-o") || !strcmp (a, "--outputname") || !strcmp (a, "--tmpdir")) {
      if (++i >= (unsigned int) argc) print_help (1);
      if (strcmp (a, "--tmpdir")) outputname = argv[i];          /* --tmpdir: accepted, no temporary files here */
      continue;
    }
    if (!strcmp (a, "--stream")) continue;                        /* inputs are mapped and parsed in one go */
    if (!strcmp (a, "--index")) { create_index = 1; continue; }
    if (!strcmp (a, "-D")) { debug += 1; continue; }
    if (a[0] == '-' && a[1]) print_help (1);                      /* unknown option; a lone "-" would be stdin */
    if (n_inputs < MAX_INPUTS) inputs[n_inputs++] = a;
  }
  wordlength = (unsigned int) v_word;
  min = (unsigned int) v_min;
  max = (unsigned int) v_max;
  nthreads = (unsigned int) v_threads;
  ntables = (unsigned int) v_tables;
  tablesize = (unsigned long long) v_tsize;

  (void) nthreads;
  (void) ntables;

  if (!n_inputs) {
    fprintf (stderr, "Error: No FastA/FastQ file specified!\n");
    print_help (1);
  }
  if (wordlength < 1 || wordlength > 32) {
    fprintf (stderr, "Error: Invalid word-length %d (must be 1 - 32)!\n", wordlength);
    print_help (1);
  }
  if (min < 1) {
    fprintf (stderr, "Error: Invalid frequency cut-off: %d! Must be positive.\n", min);
    print_help (1);
  }
  if (max < min) {
    fprintf (stderr, "Error: Invalid frequency range: %u-%u!\n", min, max);
    print_help (1);
  }
  if (strlen (outputname) > 200) {
    fprintf (stderr, "Error: Output name exceeds the 200 character limit.");
    return 1;
  }
  if (create_index) {
    fprintf (stderr, "Error: --index is not supported by the GPU list maker\n");
    return 1;
  }
  if (tablesize < 1) tablesize = 1;
  if (getenv ("GT4GPU_FASTA_BLOCK")) fasta_block = strtoull (getenv ("GT4GPU_FASTA_BLOCK"), NULL, 10);
  if (getenv ("GT4GPU_FOLD_TABLES")) fold_tables = (unsigned) strtoul (getenv ("GT4GPU_FOLD_TABLES"), NULL, 10);
  if (getenv ("GT4GPU_FOLD_BYTES")) fold_bytes = strtoull (getenv ("GT4GPU_FOLD_BYTES"), NULL, 10);
  if (fold_tables < 2) fold_tables = 2;
  if (fold_tables > 4096) fold_tables = 4096;
  if (fasta_block < 1) fasta_block = 1;
  for (i = 0; i < n_inputs; i++) {
    struct stat s;
    size_t len = strlen (inputs[i]);
    if (!strcmp (inputs[i], "-") || (len > 3 && !strcmp (inputs[i] + len - 3, ".gz"))) {
      fprintf (stderr, "Error: stdin and .gz inputs are not supported by the GPU list maker: %s\n", inputs[i]);
      return 1;
    }
    if (stat (inputs[i], &s)) {
      fprintf (stderr, "main: No such file (cannot stat): %s\n", inputs[i]);
      exit (1);
    }
  }

  /* read every file into tables of at most `tablesize` words; each full table is sorted and counted on the GPU */
  uint64_t *table = NULL, table_cap = 0, table_fill = 0;
  for (i = 0; i < n_inputs; i++) {
    struct stat s;
    uint64_t n_words = 0, done = 0;
    double t0 = now ();
    int fd = open (inputs[i], O_RDONLY);
    if (fd < 0 || fstat (fd, &s)) {
      fprintf (stderr, "Cannot open %s\n", inputs[i]);
      return 1;
    }
    if (s.st_size == 0) { close (fd); continue; }
    const unsigned char *text = (const unsigned char *) mmap (NULL, s.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close (fd);
    if (text == MAP_FAILED) {
      fprintf (stderr, "Cannot map %s\n", inputs[i]);
      return 1;
    }
    if (text[0] == '>') {
      /* FastA: parsed on the GPU, block by block (blocks end where a record ends, like the reference's
       * gt4_sequence_block_split, src/sequence-block.c:149-207); the words never visit the host */
      const unsigned char *zero = (const unsigned char *) memchr (text, 0, s.st_size);
      uint64_t size = zero ? (uint64_t) (zero - text) : (uint64_t) s.st_size, off = 0;
      int rc = 0;
      while (off < size && !rc) {
        uint64_t end = size, n_block = 0, taken = 0;
        uint64_t *d_words = NULL;
        if (size - off > fasta_block) {
          const unsigned char *q = text + off + fasta_block;
          while ((q = (const unsigned char *) memchr (q, '\n', text + size - q)) != NULL) {
            if (q + 1 < text + size && q[1] == '>') { end = (uint64_t) (q + 1 - text); break; }
            q++;
          }
        }
        rc = gt4gpu_fasta_words_device (text + off, end - off, wordlength, &d_words, &n_block);
        t_read += now () - t0;
        n_read += n_block;
        while (!rc && taken < n_block) {
          uint64_t take = n_block - taken < tablesize ? n_block - taken : tablesize;
          rc = flush_table (d_words + taken, take, 1, wordlength);
          taken += take;
        }
        gt4gpu_device_free (d_words);
        t0 = now ();
        off = end;
      }
      munmap ((void *) text, s.st_size);
      if (rc) {
        fprintf (stderr, "Error: %s\n", gt4gpu_last_error ());
        return 1;
      }
      continue;
    }
    if (text[0] == '@' && (uint64_t) s.st_size <= fastq_device_max) {
      /* FastQ: well-formed four-line records are parsed on the GPU in one go; anything the reference's reader would give
       * up on (GT4GPU_ERR_FORMAT) goes to the serial reader below, which keeps the words up to that point like glistmaker */
      uint64_t n_block = 0, taken = 0;
      uint64_t *d_words = NULL;
      int rc = gt4gpu_fasta_words_device (text, (uint64_t) s.st_size, wordlength, &d_words, &n_block);
      if (rc == 0) {
        t_read += now () - t0;
        n_read += n_block;
        while (!rc && taken < n_block) {
          uint64_t take = n_block - taken < tablesize ? n_block - taken : tablesize;
          rc = flush_table (d_words + taken, take, 1, wordlength);
          taken += take;
        }
        gt4gpu_device_free (d_words);
        munmap ((void *) text, s.st_size);
        if (rc) {
          fprintf (stderr, "Error: %s\n", gt4gpu_last_error ());
          return 1;
        }
        continue;
      }
      if (rc != GT4GPU_ERR_FORMAT) {
        fprintf (stderr, "Error: %s\n", gt4gpu_last_error ());
        return 1;
      }
    }
    uint64_t *words = (uint64_t *) malloc ((size_t) s.st_size * sizeof (uint64_t) + 8);
    if (!words) {
      fprintf (stderr, "Out of memory reading %s\n", inputs[i]);
      return 1;
    }
    int rc = gt4gpu_sequence_words (text, (uint64_t) s.st_size, wordlength, words, (uint64_t) s.st_size, &n_words);
    if (rc) fprintf (stderr, "fasta_reader_read_nwords: Reader %s: %s\n", inputs[i], gt4gpu_last_error ());
    munmap ((void *) text, s.st_size);
    t_read += now () - t0;
    n_read += n_words;
    while (done < n_words) {
      uint64_t take = n_words - done;
      if (table_fill == 0 && take >= tablesize) {
        /* a whole table straight from the reader's buffer */
        if ((rc = flush_table (words + done, tablesize, 0, wordlength))) goto gpu_error;
        done += tablesize;
        continue;
      }
      if (!table) {
        table_cap = tablesize < (1ULL << 20) ? tablesize : (1ULL << 20);
        table = (uint64_t *) malloc (table_cap * sizeof (uint64_t));
      }
      if (take > tablesize - table_fill) take = tablesize - table_fill;
      if (table_fill + take > table_cap) {
        while (table_cap < table_fill + take) table_cap *= 2;
        if (table_cap > tablesize) table_cap = tablesize;
        table = (uint64_t *) realloc (table, table_cap * sizeof (uint64_t));
      }
      if (!table) {
        fprintf (stderr, "Out of memory\n");
        return 1;
      }
      memcpy (table + table_fill, words + done, take * sizeof (uint64_t));
      table_fill += take;
      done += take;
      if (table_fill == tablesize) {
        if ((rc = flush_table (table, table_fill, 0, wordlength))) goto gpu_error;
        table_fill = 0;
      }
    }
    free (words);
    continue;
gpu_error:
    fprintf (stderr, "Error: %s\n", gt4gpu_last_error ());
    free (words);
    return 1;
  }
  if (table_fill) {
    if (flush_table (table, table_fill, 0, wordlength)) {
      fprintf (stderr, "Error: %s\n", gt4gpu_last_error ());
      return 1;
    }
  }
  free (table);

  snprintf (tmp_name, sizeof (tmp_name), "%s_%u.list.tmp", outputname, wordlength);
  snprintf (out_name, sizeof (out_name), "%s_%u.list", outputname, wordlength);
  int ofile = open (tmp_name, O_WRONLY | O_CREAT | O_TRUNC, 0666);
  if (ofile < 0) {
    fprintf (stderr, "Cannot create output file %s\n", tmp_name);
    exit (1);
  }
  double t0 = now ();
  int rc = 0;
  if (n_tables > 0) {
    /* final collation (gt4_write_union with cutoff 1, :314-333) */
    gt4gpu_header header;
    rc = gt4gpu_write_union ((const gt4gpu_list *const *) table_lists, n_tables, 1, ofile, &header);
    if (debug) fprintf (stderr, "Words %llu, unique %llu\n", (unsigned long long) header.total_count, (unsigned long long) header.n_words);
  } else {
    gt4gpu_header header;
    gt4gpu_header_init (&header, wordlength);
    if (write (ofile, &header, sizeof (header)) != (ssize_t) sizeof (header)) rc = GT4GPU_ERR_IO;
  }
  close (ofile);
  t_collate = now () - t0;
  if (rc) {
    fprintf (stderr, "Error: %s\n", gt4gpu_last_error ());
    unlink (tmp_name);
    return 1;
  }
  if (rename (tmp_name, out_name)) {
    fprintf (stderr, "Cannot rename %s to %s\n", tmp_name, out_name);
  }
  if (debug) {
    fprintf (stderr, "Read %llu words at %.2f (%.0f words/s)\n", n_read, t_read, n_read / (t_read > 0 ? t_read : 1e-9));
    fprintf (stderr, "Sort %llu words at %.2f (%.0f words/s)\n", n_read, t_sort, n_read / (t_sort > 0 ? t_sort : 1e-9));
    fprintf (stderr, "Collate and write %u tables at %.2f\n", n_tables, t_collate);
  }
  for (i = 0; i < n_tables; i++) {
    gt4gpu_list_close (table_lists[i]);
    gt4gpu_result_free (&tables[i]);
  }
  gt4gpu_shutdown ();
  return 0;
}
