/*
 * gt4gpu-compare -- drop-in for the set-operation path of GenomeTester4's glistcompare, running the
 * merges on a B200 through libgt4gpu (include/gt4gpu.h).
 *
 * Flag grammar, validation order, messages, output file names, tmp+rename behaviour, stdout text
 * and exit codes follow main() of /root/reference/src/glistcompare.c:84-429 (help text :1171-1196),
 * so that scripts such as MakeUnion.pl can switch binaries.  Not carried over (SURVEY.md section 2,
 * out of scope): -mm N with N > 0 and -ss/--subset; they exit with status 1 and a message.  GT4I index files are
 * accepted as inputs like the reference does (their k-mer table read as a list, counts = number of locations).  --stream selects the stream reader's header rules; --disable_scouts is accepted (there is
 * no read-ahead thread to disable).
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
#include <sys/wait.h>
#include <unistd.h>

#include "gt4gpu.h"

#define MAX_FILES 1024   /* glistcompare.c:76 */

static int debug = 0;

static void
print_help (int exit_value)
{
  fprintf (stdout, "glistcompare version %u.%u.%u (%s)\n", GT4GPU_VERSION_MAJOR, GT4GPU_VERSION_MINOR, GT4GPU_VERSION_MICRO, "stable");
  fprintf (stdout, "Usage: glistcompare INPUTLIST1 [INPUTLIST2...] METHOD [OPTIONS]\n");
  fprintf (stdout, "Options:\n");
  fprintf (stdout, "    -v, --version            - print version information and exit\n");
  fprintf (stdout, "    -h, --help               - print this usage screen and exit\n");
  fprintf (stdout, "    -u, --union              - union of input lists\n");
  fprintf (stdout, "    -i, --intersection       - intersection of input lists\n");
  fprintf (stdout, "    -d, --difference         - difference of input lists\n");
  fprintf (stdout, "    -dd, --double_difference - double difference of input lists\n");
  fprintf (stdout, "    -du, --diff_union        - subtract first list from the second and finds difference\n");
  fprintf (stdout, "    -mm, --mismatch   NUMBER - specify number of mismatches (default 0, can be used with -diff and -ddiff)\n");
  fprintf (stdout, "    -c, --cutoff NUMBER      - specify frequency cut-off (default 1)\n");
  fprintf (stdout, "    -o, --outputname STRING  - specify output name (default \"out\")\n");
  fprintf (stdout, "    -r, --rule STRING        - specify rule how final frequencies are calculated (default, add, subtract, min, max, first, second, 1, 2)\n");
  fprintf (stdout, "                               NOTE: rules min, subtract, first and second can only be used with finding the intersection.\n");
  fprintf (stdout, "    -ss, --subset METHOD SIZE - make subset with given method (rand, rand_unique, rand_weighted_unique)\n");
  fprintf (stdout, "    --seed INTEGER           - Set seed of random number generator (default uses start time)\n");
  fprintf (stdout, "    --count_only             - output count of k-mers instead of k-mers themself\n");
  fprintf (stdout, "    --disable_scouts         - disable list read-ahead in background thread\n");
  fprintf (stdout, "    --stream                 - read input as stream (do not memory map files)\n");
  fprintf (stdout, "    -D                       - increase debug level\n");
  exit (exit_value);
}

static double
now (void)
{
  struct timeval tv;
  gettimeofday (&tv, NULL);
  return tv.tv_sec + tv.tv_usec * 1e-6;
}

/* one output of the two-list run: "<out>_<k>_<tag>.list" written through a .tmp and renamed (:814-834,:908-953) */
static int
write_output (const gt4gpu_result *res, const char *out, unsigned int wlen, const char *tag, mode_t mode)
{
  char tmp_name[2048], name[2048];
  int fd, rc;
  snprintf (tmp_name, sizeof (tmp_name), "%s_%d_%s.list.tmp", out, wlen, tag);
  snprintf (name, sizeof (name), "%s_%d_%s.list", out, wlen, tag);
  fd = open (tmp_name, O_WRONLY | O_CREAT | O_TRUNC, mode);
  if (fd < 0) {
    fprintf (stderr, "Error: Cannot create output file %s\n", tmp_name);
    return 1;
  }
  rc = gt4gpu_write_list (res, fd);
  close (fd);
  if (rc) {
    fprintf (stderr, "Error: %s\n", gt4gpu_last_error ());
    unlink (tmp_name);
    return 1;
  }
  if (debug) fprintf (stderr, "Renaming %s to %s\n", tmp_name, name);
  if (rename (tmp_name, name)) {
    fprintf (stderr, "Error: Cannot rename %s to %s\n", tmp_name, name);
    return 1;
  }
  return 0;
}


/* ---- several GPUs of one box, without Python (--gpus N, or GT4GPU_GPUS=N) -------------------------------------------
 * The path shards by key range with no payload exchange.  The parent -- which never touches
 * CUDA -- maps the list files, lets gt4gpu_plan_splitters cut the key space into N ranges holding the same number of
 * records, creates the output files and forks one child per device.  Child r loads its record range of every list,
 * merges on device r, reports {n_out, sum} of every output up a pipe, gets its record offsets back and writes its slice
 * with gt4gpu_write_records_at (48 + 12 * offset).  The parent adds up the totals, writes the headers and renames.
 * Every stream of every mode goes through the same four "slots": two-list runs use union / intersection / diff1 /
 * diff2, N-list runs use slot 0 (union) and slot 1 (intersection). */
struct shard_job {
  unsigned int nfiles, wlen, cutoff, count_override;
  const char **fnames;
  int stream, rule, subtraction, countonly;
  uint32_t ops;              /* slots requested */
  const char *outputname;
};

static ssize_t
xfer (int fd, void *buf, size_t bytes, int writing)
{
  size_t done = 0;
  while (done < bytes) {
    ssize_t r = writing ? write (fd, (char *) buf + done, bytes - done) : read (fd, (char *) buf + done, bytes - done);
    if (r < 0 && errno == EINTR) continue;
    if (r <= 0) return -1;
    done += (size_t) r;
  }
  return (ssize_t) done;
}

static int
shard_child (const struct shard_job *job, unsigned int rank, unsigned int ngpus, const uint64_t *bounds, int up, int down, char tmp_names[4][2048])
{
  static gt4gpu_list *lists[MAX_FILES];
  gt4gpu_result res[4];
  uint64_t report[9], offsets[4];
  unsigned int i;
  int s, rc = 0;
  memset (res, 0, sizeof (res));
  memset (report, 0, sizeof (report));
  {
    const int n_dev = gt4gpu_device_count ();       /* more shards than devices: they share (tests on one GPU) */
    if (gt4gpu_init (n_dev > 0 ? (int) (rank % (unsigned int) n_dev) : (int) rank)) goto fail;
  }
  for (i = 0; i < job->nfiles; i++) {
    const uint64_t lo = bounds[i * (ngpus + 1) + rank], hi = bounds[i * (ngpus + 1) + rank + 1];
    if (gt4gpu_list_open_range (job->fnames[i], job->stream, lo, hi - lo, &lists[i])) goto fail;
  }
  if (job->nfiles == 2) {
    rc = gt4gpu_compare2 (lists[0], lists[1], job->ops, job->rule, job->cutoff, job->count_override, job->subtraction, job->countonly, res);
  } else {
    /* a rejected rule (GT4GPU_ERR_ARG) is reported, not fatal: the parent reproduces the reference's handling */
    if (job->ops & 1u) rc = gt4gpu_union_multi ((const gt4gpu_list *const *) lists, job->nfiles, job->cutoff, job->rule, job->count_override, job->countonly, &res[0]);
    if (rc == GT4GPU_ERR_ARG) { report[8] |= 1u; rc = 0; }
    if (!rc && (job->ops & 2u)) rc = gt4gpu_intersect_multi ((const gt4gpu_list *const *) lists, job->nfiles, job->cutoff, job->rule, job->count_override, job->countonly, &res[1]);
    if (rc == GT4GPU_ERR_ARG) { report[8] |= 2u; rc = 0; }
  }
  if (rc) goto fail;
  for (s = 0; s < 4; s++) { report[2 * s] = res[s].n_words; report[2 * s + 1] = res[s].total_count; }
  if (xfer (up, report, sizeof (report), 1) < 0) return 1;
  if (xfer (down, offsets, sizeof (offsets), 0) < 0) return 1;
  for (s = 0; s < 4 && !job->countonly; s++) {
    int fd;
    if (!((job->ops >> s) & 1u) || ((report[8] >> s) & 1u) || !res[s].n_words) continue;
    fd = open (tmp_names[s], O_WRONLY);
    if (fd < 0) return 1;
    rc = gt4gpu_write_records_at (&res[s], fd, offsets[s]);
    close (fd);
    if (rc) goto fail;
  }
  for (s = 0; s < 4; s++) gt4gpu_result_free (&res[s]);
  for (i = 0; i < job->nfiles; i++) gt4gpu_list_close (lists[i]);
  gt4gpu_shutdown ();
  return 0;
fail:
  fprintf (stderr, "Error: (GPU %u) %s\n", rank, gt4gpu_last_error ());
  report[8] = ~0ull;
  xfer (up, report, sizeof (report), 1);
  return 1;
}

/* returns the exit status of the run, or -1 when the inputs do not allow sharding (the caller runs on one GPU) */
static int
run_sharded (const struct shard_job *job, unsigned int ngpus)
{
  static const char *tags2[4] = { "union", "intrsec", "0_diff1", "0_diff2" };
  const void *keys[MAX_FILES];
  size_t strides[MAX_FILES];
  uint64_t sizes[MAX_FILES];
  void *maps[MAX_FILES];
  size_t map_bytes[MAX_FILES];
  uint64_t *bounds;
  char tmp_names[4][2048], names[4][2048];
  int up[64][2], down[64][2];
  pid_t pids[64];
  uint64_t reports[64][9], totals[8];
  unsigned int i, r;
  int s, v = 0;
  const mode_t mode = job->nfiles == 2 ? 0666 : (S_IRUSR | S_IWUSR | S_IRGRP | S_IROTH);
  if (ngpus > 64) ngpus = 64;
  for (i = 0; i < job->nfiles; i++) {
    gt4gpu_header h;
    struct stat st;
    int fd;
    if (gt4gpu_list_read_header (job->fnames[i], job->stream, &h) || h.count_bytes == 8) return -1;      /* (GT4I index inputs: one GPU) */
    fd = open (job->fnames[i], O_RDONLY);
    if (fd < 0 || fstat (fd, &st) < 0) return -1;
    map_bytes[i] = (size_t) st.st_size;
    maps[i] = mmap (NULL, map_bytes[i], PROT_READ, MAP_PRIVATE, fd, 0);
    close (fd);
    if (maps[i] == MAP_FAILED) return -1;
    keys[i] = (const char *) maps[i] + h.list_start;
    strides[i] = 12;
    sizes[i] = h.n_words;
  }
  bounds = (uint64_t *) malloc (sizeof (uint64_t) * job->nfiles * (ngpus + 1));
  if (!bounds || gt4gpu_plan_splitters (keys, strides, sizes, job->nfiles, ngpus, bounds, NULL)) return -1;
  for (i = 0; i < job->nfiles; i++) munmap (maps[i], map_bytes[i]);
  if (job->nfiles > 2 && (job->ops & 2u)) {
    /* an empty list empties the intersection (glistcompare.c:631-636); a rank whose RANGE of a list is empty does not */
    int any_empty = 0;
    for (i = 0; i < job->nfiles; i++) any_empty |= sizes[i] == 0;
    if (any_empty) return -1;
  }
  for (s = 0; s < 4; s++) {
    if (!((job->ops >> s) & 1u)) continue;
    snprintf (tmp_names[s], sizeof (tmp_names[s]), "%s_%d_%s.list.tmp", job->outputname, job->wlen, tags2[s]);
    snprintf (names[s], sizeof (names[s]), "%s_%d_%s.list", job->outputname, job->wlen, tags2[s]);
    if (!job->countonly) {
      int fd = open (tmp_names[s], O_WRONLY | O_CREAT | O_TRUNC, mode);
      if (fd < 0) {
        fprintf (stderr, "Error: Cannot create output file %s\n", tmp_names[s]);
        return 1;
      }
      close (fd);
    }
  }
  fflush (stdout);
  fflush (stderr);
  for (r = 0; r < ngpus; r++) {
    if (pipe (up[r]) || pipe (down[r])) return 1;
    pids[r] = fork ();
    if (pids[r] < 0) return 1;
    if (pids[r] == 0) {
      close (up[r][0]);
      close (down[r][1]);
      _exit (shard_child (job, r, ngpus, bounds, up[r][1], down[r][0], tmp_names));
    }
    close (up[r][1]);
    close (down[r][0]);
  }
  memset (totals, 0, sizeof (totals));
  {
    uint64_t rejected = 0;
    int failed = 0;
    for (r = 0; r < ngpus; r++) {
      if (xfer (up[r][0], reports[r], sizeof (reports[r]), 0) < 0 || reports[r][8] == ~0ull) failed = 1;
      else rejected |= reports[r][8];
    }
    for (r = 0; r < ngpus; r++) {
      uint64_t offsets[4];
      for (s = 0; s < 4; s++) {
        offsets[s] = totals[2 * s];
        if (!failed) { totals[2 * s] += reports[r][2 * s]; totals[2 * s + 1] += reports[r][2 * s + 1]; }
      }
      if (!failed) xfer (down[r][1], offsets, sizeof (offsets), 1);
      close (down[r][1]);
      close (up[r][0]);
    }
    for (r = 0; r < ngpus; r++) {
      int status = 0;
      waitpid (pids[r], &status, 0);
      if (!WIFEXITED (status) || WEXITSTATUS (status)) failed = 1;
    }
    if (failed) {
      for (s = 0; s < 4; s++) if (((job->ops >> s) & 1u) && !job->countonly) unlink (tmp_names[s]);
      return 1;
    }
    for (s = 0; s < 4; s++) {
      if (!((job->ops >> s) & 1u)) continue;
      if ((rejected >> s) & 1u) {
        /* union_multi / intersect_multi refused the rule: no file, zero counts, status 1 (glistcompare.c:518-523, :622-627) */
        fprintf (stderr, "%s: Invalid rule %u (only %s allowed)\n", s == 0 ? "union_multi" : "intersect_multi", job->rule,
                 s == 0 ? "ADD, MAX and NUMBER" : "ADD, MIN, MAX and NUMBER");
        if (!job->countonly) unlink (tmp_names[s]);
        if (job->countonly || debug) fprintf (stdout, "NUnique\t0\nNTotal\t0\n");
        v = 1;
        continue;
      }
      v = 0;
      if (!job->countonly) {
        gt4gpu_header h;
        int fd = open (tmp_names[s], O_WRONLY);
        gt4gpu_header_init (&h, job->wlen);
        h.n_words = totals[2 * s];
        h.total_count = totals[2 * s + 1];
        if (fd < 0 || pwrite (fd, &h, sizeof (h), 0) != (ssize_t) sizeof (h) || ftruncate (fd, (off_t) (sizeof (h) + 12 * totals[2 * s]))) {
          fprintf (stderr, "Error: Cannot write %s\n", tmp_names[s]);
          return 1;
        }
        close (fd);
        if (debug) fprintf (stderr, "Renaming %s to %s\n", tmp_names[s], names[s]);
        if (rename (tmp_names[s], names[s])) {
          fprintf (stderr, "Error: Cannot rename %s to %s\n", tmp_names[s], names[s]);
          return 1;
        }
      }
      if (job->countonly || (debug && job->nfiles > 2))
        fprintf (stdout, "NUnique\t%llu\nNTotal\t%llu\n", (unsigned long long) totals[2 * s], (unsigned long long) totals[2 * s + 1]);
    }
  }
  free (bounds);
  return v ? 1 : 0;
}

int
main (int argc, const char *argv[])
{
  int arg_idx, v = 0;
  unsigned int i, nfiles = 0, n_index = 0;
  static const char *fnames[MAX_FILES];
  static gt4gpu_list *lists[MAX_FILES];
  char *end;
  int rule = GT4GPU_RULE_DEFAULT;
  unsigned int wlen = 0, err = 0;
  unsigned int cutoff = 1, nmm = 0, count_override = 1;
  int find_union = 0, find_intrsec = 0, find_diff = 0, find_ddiff = 0, subtraction = 0, countonly = 0, print_operation = 0;
  int find_subset = 0, stream = 0;
  unsigned int ngpus = getenv ("GT4GPU_GPUS") ? (unsigned int) strtoul (getenv ("GT4GPU_GPUS"), NULL, 10) : 1;
  const char *outputname = "out";

  if (argc <= 1) print_help (1);

  for (arg_idx = 1; arg_idx < argc; arg_idx++) {
    const char *a = argv[arg_idx];
    if (a[0] != '-') {
      if (nfiles >= MAX_FILES) {
        fprintf (stderr, "Too many file arguments (max %d)\n", MAX_FILES);
        print_help (1);
      }
      fnames[nfiles++] = a;
      continue;
    }
    if (!strcmp (a, "-v") || !strcmp (a, "--version")) {
      fprintf (stdout, "glistcompare version %u.%u.%u (%s)\n", GT4GPU_VERSION_MAJOR, GT4GPU_VERSION_MINOR, GT4GPU_VERSION_MICRO, "stable");
      return 0;
    } else if (!strcmp (a, "-h") || !strcmp (a, "--help") || !strcmp (a, "-?")) {
      print_help (0);
    } else if (!strcmp (a, "-o") || !strcmp (a, "--outputname")) {
      if (!argv[arg_idx + 1] || argv[arg_idx + 1][0] == '-') {
        fprintf (stderr, "Warning: No output name specified!\n");
        arg_idx += 1;
        continue;
      }
      outputname = argv[arg_idx + 1];
      arg_idx += 1;
    } else if (!strcmp (a, "-c") || !strcmp (a, "--cutoff")) {
      if (!argv[arg_idx + 1]) {
        fprintf (stderr, "Warning: No frequency cut-off specified! Using the default value: %d.\n", cutoff);
        continue;
      }
      cutoff = strtol (argv[arg_idx + 1], &end, 10);
      if (*end != 0) {
        fprintf (stderr, "Error: Invalid frequency cut-off: %s! Must be an integer.\n", argv[arg_idx + 1]);
        print_help (1);
      }
      arg_idx += 1;
    } else if (!strcmp (a, "-mm") || !strcmp (a, "--mismatch")) {
      if (!argv[arg_idx + 1]) {
        fprintf (stderr, "Warning: No number of mismatches specified!");
        continue;
      }
      nmm = strtol (argv[arg_idx + 1], &end, 10);
      if (*end != 0) {
        fprintf (stderr, "Error: Invalid number of mismatches: %s! Must be an integer.\n", argv[arg_idx + 1]);
        print_help (1);
      }
      arg_idx += 1;
    } else if (!strcmp (a, "-u") || !strcmp (a, "--union")) {
      find_union = 1;
    } else if (!strcmp (a, "-i") || !strcmp (a, "--intersection")) {
      find_intrsec = 1;
    } else if (!strcmp (a, "-d") || !strcmp (a, "--difference")) {
      find_diff = 1;
    } else if (!strcmp (a, "-dd") || !strcmp (a, "--double_difference")) {
      find_ddiff = 1;
    } else if (!strcmp (a, "-du") || !strcmp (a, "--diff_union")) {
      find_diff = 1;
      subtraction = 1;
    } else if (!strcmp (a, "--count_only")) {
      countonly = 1;
    } else if (!strcmp (a, "-r") || !strcmp (a, "--rule")) {
      arg_idx += 1;
      if (arg_idx >= argc) print_help (1);
      a = argv[arg_idx];
      if ((*a >= '1') && (*a <= '9')) {
        rule = GT4GPU_RULE_NUMBER;
        count_override = strtol (a, &end, 10);
      } else if (!strcmp (a, "default")) {
        rule = GT4GPU_RULE_DEFAULT;
      } else if (!strcmp (a, "add") || !strcmp (a, "sum")) {
        rule = GT4GPU_RULE_ADD;
      } else if (!strcmp (a, "subtract")) {
        rule = GT4GPU_RULE_SUBTRACT;
      } else if (!strcmp (a, "min")) {
        rule = GT4GPU_RULE_MIN;
      } else if (!strcmp (a, "max")) {
        rule = GT4GPU_RULE_MAX;
      } else if (!strcmp (a, "first")) {
        rule = GT4GPU_RULE_FIRST;
      } else if (!strcmp (a, "second")) {
        rule = GT4GPU_RULE_SECOND;
      }
    } else if (!strcmp (a, "-ss") || !strcmp (a, "--subset")) {
      find_subset = 1;
      arg_idx += 2;
      if (arg_idx >= argc) print_help (1);
    } else if (!strcmp (a, "--seed")) {
      arg_idx += 1;
      if (arg_idx >= argc) print_help (1);
    } else if (!strcmp (a, "--print_operation")) {
      print_operation = 1;
    } else if (!strcmp (a, "--disable_scouts")) {
      /* nothing to disable */
    } else if (!strcmp (a, "--stream")) {
      stream = 1;
    } else if (!strcmp (a, "--gpus")) {          /* (not a glistcompare flag) key-range shards on N devices, one process each */
      arg_idx += 1;
      if (arg_idx >= argc) print_help (1);
      ngpus = (unsigned int) strtoul (argv[arg_idx], &end, 10);
      if (*end != 0 || ngpus < 1) print_help (1);
    } else if (!strcmp (a, "-D")) {
      debug += 1;
    } else {
      fprintf (stderr, "Unknown argument: %s!\n", a);
      print_help (1);
    }
  }
  if (debug) fprintf (stderr, "Rule: %d\n", rule);
  if (debug) fprintf (stderr, "Num files: %d\n", nfiles);

  if (nmm || find_subset) {
    fprintf (stderr, "Error: %s is not supported by the GPU set-operation engine (use the reference glistcompare)\n",
             nmm ? "-mm/--mismatch" : "-ss/--subset");
    exit (1);
  }

  /* Build list of containers (:250-291): sniff the 4-byte tag, check word lengths */
  for (i = 0; i < nfiles; i++) {
    FILE *ifs;
    unsigned char tag[4] = { 0, 0, 0, 0 };
    gt4gpu_header hdr;
    ifs = fopen (fnames[i], "r");
    if (!ifs) {
      fprintf (stderr, "Error: Cannot open %s\n", fnames[i]);
      err = 1;
      continue;
    }
    if (fread (tag, 1, 4, ifs) != 4) memset (tag, 0, 4);
    fclose (ifs);
    if (!memcmp (tag, "I4TG", 4)) n_index += 1;
    if (!memcmp (tag, "C4TG", 4) || !memcmp (tag, "I4TG", 4)) {     /* list, or index read as a list (:264-270) */
      if (gt4gpu_list_read_header (fnames[i], stream, &hdr)) {
        fprintf (stderr, "%s\n", gt4gpu_last_error ());
        fprintf (stderr, "Error: File %s is invalid or corrupted\n", fnames[i]);
        err = 1;
        continue;
      }
    } else {
      fprintf (stderr, "Error: File %s has unknown format\n", fnames[i]);
      err = 1;
      continue;
    }
    if (!wlen) {
      wlen = hdr.word_length;
    } else if (hdr.word_length != wlen) {
      fprintf (stderr, "Error: File %s has different word length (%u != %u)\n", fnames[i], hdr.word_length, wlen);
      err = 1;
    }
  }
  if (err) {
    fprintf (stderr, "Stopping...\n");
    exit (1);
  }

  if (nfiles < 2) {
    fprintf (stderr, "Error: At least 2 list/index files are needed\n");
    exit (1);
  }
  if (nfiles > 2) {
    if (!(find_union || find_intrsec) || find_diff || find_ddiff) {
      fprintf (stderr, "Error: Algorithm incompatible with multiple files!\n");
      print_help (1);
    }
  }
  if (find_ddiff) find_diff = 1;
  if (!find_diff && subtraction) fprintf (stderr, "Warning: Subtraction is not used!\n");
  if (strlen (outputname) > 200) {
    fprintf (stderr, "Error: Output name exceeds the 200 character limit.\n");
    exit (1);
  }
  if (!find_intrsec && (rule == GT4GPU_RULE_MIN || rule == GT4GPU_RULE_FIRST || rule == GT4GPU_RULE_SECOND)) {
    fprintf (stderr, "Error: Rules min, fist and second can only be used with finding the intersection.\n");
    exit (1);
  }
  if ((!find_intrsec && !find_diff) && (rule == GT4GPU_RULE_SUBTRACT)) {
    fprintf (stderr, "Error: Rule subtract can only be used with intersection and difference.\n");
    exit (1);
  }
  if (print_operation) {
    fprintf (stdout, "Operation\t%s%s%s%s\trule\t%u\nFiles\t%u\n", (find_union) ? "U" : "", (find_intrsec) ? "I" : "",
             (find_diff) ? "D" : "", (find_ddiff) ? "X" : "", rule, nfiles);
    for (i = 0; i < nfiles; i++) fprintf (stdout, "%u\t%s\n", i, fnames[i]);
  }

  if (ngpus > 1) {
    struct shard_job job;
    int rc;
    job.nfiles = nfiles; job.wlen = wlen; job.cutoff = cutoff; job.count_override = count_override;
    job.fnames = fnames; job.stream = stream; job.rule = rule; job.subtraction = subtraction; job.countonly = countonly;
    job.outputname = outputname;
    job.ops = nfiles == 2 ? (uint32_t) ((find_union ? GT4GPU_OP_UNION : 0) | (find_intrsec ? GT4GPU_OP_INTRSEC : 0) |
                                        (find_diff ? GT4GPU_OP_DIFF : 0) | (find_ddiff ? GT4GPU_OP_DDIFF : 0))
                          : (uint32_t) ((find_union ? 1 : 0) | (find_intrsec ? 2 : 0));
    rc = job.ops ? run_sharded (&job, ngpus) : 0;
    if (rc >= 0) return rc;
    /* (inputs that cannot be sharded, e.g. GT4I index files: one GPU below) */
  }

  /* from here on the GPU is needed */
  double t_phase = now ();
  if (gt4gpu_init (-1)) {
    fprintf (stderr, "Error: %s\n", gt4gpu_last_error ());
    exit (1);
  }
  if (debug) fprintf (stderr, "gt4gpu: device init %.3f s\n", now () - t_phase);
  t_phase = now ();
  if (nfiles == 2 && n_index == 0 && !getenv ("GT4GPU_NO_FILE_PIPELINE")) {
    /* two list files: the pipelined file-to-file path (read part p + 1 | merge part p | write part p - 1) */
    static const char *tags[4] = { "union", "intrsec", "0_diff1", "0_diff2" };
    const uint32_t ops = (find_union ? GT4GPU_OP_UNION : 0) | (find_intrsec ? GT4GPU_OP_INTRSEC : 0) |
                         (find_diff ? GT4GPU_OP_DIFF : 0) | (find_ddiff ? GT4GPU_OP_DDIFF : 0);
    char tmp_name[4][2048], name[4][2048];
    int fds[4] = { -1, -1, -1, -1 }, s;
    uint64_t n_out[4], total_out[4];
    if (debug) fprintf (stderr, "compare_wordmaps: methods %u/%u/%u/%u\n", find_union, find_intrsec, find_diff, find_ddiff);
    if (!ops) {
      gt4gpu_shutdown ();
      return 0;
    }
    for (s = 0; s < 4 && !countonly; s++) {
      if (!((ops >> s) & 1u)) continue;
      snprintf (tmp_name[s], sizeof (tmp_name[s]), "%s_%d_%s.list.tmp", outputname, wlen, tags[s]);
      snprintf (name[s], sizeof (name[s]), "%s_%d_%s.list", outputname, wlen, tags[s]);
      fds[s] = open (tmp_name[s], O_WRONLY | O_CREAT | O_TRUNC, 0666);
      if (fds[s] < 0) {
        fprintf (stderr, "Error: Cannot create output file %s\n", tmp_name[s]);
        exit (1);
      }
    }
    if (gt4gpu_compare2_files (fnames[0], fnames[1], stream, ops, rule, cutoff, count_override, subtraction, countonly, fds, n_out, total_out, NULL)) {
      fprintf (stderr, "Error: %s\n", gt4gpu_last_error ());
      for (s = 0; s < 4; s++) if (fds[s] >= 0) { close (fds[s]); unlink (tmp_name[s]); }
      exit (1);
    }
    if (debug) {
      float ms_p = 0, ms_m = 0;
      gt4gpu_last_timing (&ms_p, &ms_m, NULL);
      fprintf (stderr, "gt4gpu: partition %.3f ms, merge %.3f ms on device; read + merge + write %.3f s\n", ms_p, ms_m, now () - t_phase);
    }
    for (s = 0; s < 4; s++) {
      if (!((ops >> s) & 1u)) continue;
      if (countonly) {
        fprintf (stdout, "NUnique\t%llu\nNTotal\t%llu\n", (unsigned long long) n_out[s], (unsigned long long) total_out[s]);
        continue;
      }
      close (fds[s]);
      if (debug) fprintf (stderr, "Renaming %s to %s\n", tmp_name[s], name[s]);
      if (rename (tmp_name[s], name[s])) {
        fprintf (stderr, "Error: Cannot rename %s to %s\n", tmp_name[s], name[s]);
        v = 1;
      }
    }
    gt4gpu_shutdown ();
    return v ? 1 : 0;
  }
  for (i = 0; i < nfiles; i++) {
    if (gt4gpu_list_open (fnames[i], stream, &lists[i])) {
      fprintf (stderr, "Error: %s\n", gt4gpu_last_error ());
      fprintf (stderr, "Stopping...\n");
      exit (1);
    }
  }

  if (debug) fprintf (stderr, "gt4gpu: lists loaded to the device in %.3f s\n", now () - t_phase);
  t_phase = now ();
  if (nfiles == 2) {
    /* compare_wordmaps (:789-955) */
    static const char *tags[4] = { "union", "intrsec", "0_diff1", "0_diff2" };
    gt4gpu_result res[4];
    uint32_t ops = (find_union ? GT4GPU_OP_UNION : 0) | (find_intrsec ? GT4GPU_OP_INTRSEC : 0) |
                   (find_diff ? GT4GPU_OP_DIFF : 0) | (find_ddiff ? GT4GPU_OP_DDIFF : 0);
    int s;
    memset (res, 0, sizeof (res));
    if (debug) {
      fprintf (stderr, "compare_wordmaps: methods %u/%u/%u/%u\n", find_union, find_intrsec, find_diff, find_ddiff);
      fprintf (stderr, "compare_wordmaps: List 1: %llu entries\n", (unsigned long long) gt4gpu_list_n_words (lists[0]));
      fprintf (stderr, "compare_wordmaps; List 2: %llu entries\n", (unsigned long long) gt4gpu_list_n_words (lists[1]));
    }
    if (ops) {
      if (gt4gpu_compare2 (lists[0], lists[1], ops, rule, cutoff, count_override, subtraction, countonly, res)) {
        fprintf (stderr, "Error: %s\n", gt4gpu_last_error ());
        exit (1);
      }
      if (debug) {
        float ms_p = 0, ms_m = 0;
        gt4gpu_last_timing (&ms_p, &ms_m, NULL);
        fprintf (stderr, "gt4gpu: partition %.3f ms, merge %.3f ms on device\n", ms_p, ms_m);
      }
      for (s = 0; s < 4; s++) {
        if (!((ops >> s) & 1u)) continue;
        if (!countonly) {
          if (write_output (&res[s], outputname, wlen, tags[s], 0666)) v = 1;
        } else {
          fprintf (stdout, "NUnique\t%llu\nNTotal\t%llu\n", (unsigned long long) res[s].n_words, (unsigned long long) res[s].total_count);
        }
        gt4gpu_result_free (&res[s]);
      }
    }
  } else {
    /* N-list dispatch (:366-422): union first, then intersection; a rejected rule unlinks that
     * output and the run goes on, the LAST status decides the exit code */
    int pass;
    for (pass = 0; pass < 2; pass++) {
      const int want = pass == 0 ? find_union : find_intrsec;
      const char *tag = pass == 0 ? "union" : "intrsec";
      gt4gpu_result res;
      double t_s, t_e;
      unsigned long long total = 0;
      if (!want) continue;
      memset (&res, 0, sizeof (res));
      for (i = 0; i < nfiles; i++) total += gt4gpu_list_n_words (lists[i]);
      t_s = now ();
      if (pass == 0) v = gt4gpu_union_multi ((const gt4gpu_list *const *) lists, nfiles, cutoff, rule, count_override, countonly, &res);
      else v = gt4gpu_intersect_multi ((const gt4gpu_list *const *) lists, nfiles, cutoff, rule, count_override, countonly, &res);
      t_e = now ();
      if (v && v != GT4GPU_ERR_ARG) {
        fprintf (stderr, "Error: %s\n", gt4gpu_last_error ());
        exit (1);
      }
      if (!v && debug > 0) {
        fprintf (stderr, "Combined %u maps: input %llu (%.3f Mwords/s) output %llu (%.3f Mwords/s)\n", nfiles, total,
                 total / (1000000 * (t_e - t_s)), (unsigned long long) res.n_words, res.n_words / (1000000 * (t_e - t_s)));
      }
      if (!v && !countonly) {
        if (write_output (&res, outputname, wlen, tag, S_IRUSR | S_IWUSR | S_IRGRP | S_IROTH)) exit (1);
      }
      if (countonly || debug) fprintf (stdout, "NUnique\t%llu\nNTotal\t%llu\n", (unsigned long long) res.n_words, (unsigned long long) res.total_count);
      gt4gpu_result_free (&res);
    }
  }
  if (debug) fprintf (stderr, "gt4gpu: merge + output %.3f s\n", now () - t_phase);
  for (i = 0; i < nfiles; i++) gt4gpu_list_close (lists[i]);
  gt4gpu_shutdown ();
  return v ? 1 : 0;
}
