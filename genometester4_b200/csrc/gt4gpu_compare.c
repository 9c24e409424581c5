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
#include <sys/stat.h>
#include <sys/time.h>
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

int
main (int argc, const char *argv[])
{
  int arg_idx, v = 0;
  unsigned int i, nfiles = 0;
  static const char *fnames[MAX_FILES];
  static gt4gpu_list *lists[MAX_FILES];
  char *end;
  int rule = GT4GPU_RULE_DEFAULT;
  unsigned int wlen = 0, err = 0;
  unsigned int cutoff = 1, nmm = 0, count_override = 1;
  int find_union = 0, find_intrsec = 0, find_diff = 0, find_ddiff = 0, subtraction = 0, countonly = 0, print_operation = 0;
  int find_subset = 0, stream = 0;
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

  /* from here on the GPU is needed */
  double t_phase = now ();
  if (gt4gpu_init (-1)) {
    fprintf (stderr, "Error: %s\n", gt4gpu_last_error ());
    exit (1);
  }
  if (debug) fprintf (stderr, "gt4gpu: device init %.3f s\n", now () - t_phase);
  t_phase = now ();
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
