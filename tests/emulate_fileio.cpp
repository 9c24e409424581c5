// emulate_fileio.cpp -- TEST HELPER.  Drives the host-only file output of the library
// (genometester4_b200/csrc/gt4gpu_fileio.h: write_span, write_mapped, write_all_errno) without a GPU.
// Built by tests/test_fileio_host.py with g++.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "gt4gpu_fileio.h"

using namespace gt4gpu::fileio;

// deterministic payload: byte i of a span that starts at file offset `at`
static inline unsigned char pattern (uint64_t file_pos, uint32_t seed) { return (unsigned char) (((file_pos + seed) * 2654435761ull) >> 17); }

// Writes [at, at + bytes) of `path` (opened write-only, like the CLI does) through write_all_errno; mode 1 forces the
// mapped path to be tried directly (returns 100 when it declined), mode 0 goes through write_all_errno.
extern "C" int fileio_write (const char *path, uint64_t at, uint64_t bytes, uint32_t seed, int mode, int sequential)
{
  unsigned char *buf = (unsigned char *) malloc (bytes ? bytes : 1);
  if (!buf) return 98;
  for (uint64_t i = 0; i < bytes; i++) buf[i] = pattern (at + i, seed);
  const int fd = open (path, O_WRONLY | O_CREAT, 0644);
  if (fd < 0) { free (buf); return 97; }
  int rc;
  if (mode == 1) rc = write_mapped (fd, buf, bytes, (int64_t) at) ? 0 : 100;
  else if (sequential) {
    if (lseek (fd, (off_t) at, SEEK_SET) < 0) rc = 96;
    else {
      rc = write_all_errno (fd, buf, bytes, -1);
      if (!rc && lseek (fd, 0, SEEK_CUR) != (off_t) (at + bytes)) rc = 95;      // the descriptor's position must have moved
    }
  } else rc = write_all_errno (fd, buf, bytes, (int64_t) at);
  close (fd);
  free (buf);
  return rc;
}

// 0 when [at, at + bytes) holds the pattern
extern "C" int fileio_check (const char *path, uint64_t at, uint64_t bytes, uint32_t seed)
{
  const int fd = open (path, O_RDONLY);
  if (fd < 0) return 97;
  const size_t chunk = 1u << 22;
  unsigned char *buf = (unsigned char *) malloc (chunk);
  int rc = 0;
  for (uint64_t done = 0; done < bytes && !rc;) {
    const size_t n = (size_t) (bytes - done < chunk ? bytes - done : chunk);
    const ssize_t k = pread (fd, buf, n, (off_t) (at + done));
    if (k <= 0) { rc = 94; break; }
    for (ssize_t i = 0; i < k; i++)
      if (buf[i] != pattern (at + done + (uint64_t) i, seed)) { rc = 1; break; }
    done += (uint64_t) k;
  }
  free (buf);
  close (fd);
  return rc;
}
