// gt4gpu_fileio.h -- host-only file output shared by the C-ABI layer (gt4gpu_api.cu) and the CPU tests
// (tests/emulate_fileio.cpp): writing a large span of records into a list file fast.  No CUDA here.
#pragma once

#include <errno.h>
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/statvfs.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

namespace gt4gpu {
namespace fileio {

// returns 0 or an errno value
inline int write_span (int fd, const unsigned char *p, size_t bytes, int64_t offset)
{
  while (bytes) {
    ssize_t w = (offset >= 0) ? pwrite (fd, p, bytes, offset) : write (fd, p, bytes);
    if (w < 0) {
      if (errno == EINTR) continue;
      return errno ? errno : EIO;
    }
    p += w;
    bytes -= (size_t) w;
    if (offset >= 0) offset += w;
  }
  return 0;
}

// A large span of a regular file, written through a shared mapping.  write / pwrite hold the inode's lock for the whole
// call, so any number of threads writing ONE file share a single core's page-cache speed (3.6 GB/s measured on the GPU
// boxes' tmpfs: 18 GB of union records took 5 s of a 5.4 s gt4gpu-compare pipeline).  Page faults on a shared mapping take
// no inode lock, so the copy scales with the threads (6.8 GB/s with 8; reserving the blocks with fallocate first was slower,
// 4.7 GB/s: profiles/r02_cli_file_pipeline.txt).  A page fault that cannot get a block raises SIGBUS instead of returning
// ENOSPC, so the mapping is only used when the file system reports at least twice the span free; otherwise, and on any
// failure, the caller falls back to pwrite.  The file is extended with a one-byte fallocate at the end of the span, which
// never shrinks it (several processes write disjoint ranges of one file in the sharded path).
// GT4GPU_NO_MAPPED_WRITES=1 disables the path.
inline bool write_mapped (int fd, const unsigned char *p, size_t bytes, int64_t at)
{
  static const bool disabled = getenv ("GT4GPU_NO_MAPPED_WRITES") != nullptr;
  if (disabled) return false;
  struct stat sb;
  if (fstat (fd, &sb) != 0 || !S_ISREG (sb.st_mode)) return false;
  struct statvfs vfs;
  if (fstatvfs (fd, &vfs) != 0 || (double) vfs.f_bavail * (double) vfs.f_frsize < 2.0 * (double) bytes) return false;
  char path[64];
  snprintf (path, sizeof (path), "/proc/self/fd/%d", fd);
  const int rw = open (path, O_RDWR);            // the caller's descriptor is usually write-only: a mapping needs read access
  if (rw < 0) return false;
  int r;
  do r = fallocate (rw, 0, at + (int64_t) bytes - 1, 1); while (r != 0 && errno == EINTR);
  if (r != 0) { close (rw); return false; }
  const size_t page = (size_t) sysconf (_SC_PAGESIZE);
  const int64_t map_lo = at & ~(int64_t) (page - 1);
  const size_t lead = (size_t) (at - map_lo);
  unsigned char *map = static_cast<unsigned char *> (mmap (NULL, lead + bytes, PROT_READ | PROT_WRITE, MAP_SHARED, rw, map_lo));
  if (map == MAP_FAILED) { close (rw); return false; }
  constexpr size_t PIECE = 4u << 20;
  const unsigned n_copy = std::min (12u, std::max (2u, std::thread::hardware_concurrency () * 3 / 4));
  const size_t n_pieces = (bytes + PIECE - 1) / PIECE;
  std::atomic<size_t> next{0};
  std::vector<std::thread> pool;
  for (unsigned t = 0; t < n_copy && t < n_pieces; t++)
    pool.emplace_back ([&] {
      for (size_t k = next.fetch_add (1); k < n_pieces; k = next.fetch_add (1)) {
        const size_t off = k * PIECE;
        memcpy (map + lead + off, p + off, std::min (PIECE, bytes - off));
      }
    });
  for (auto &th : pool) th.join ();
  munmap (map, lead + bytes);
  close (rw);
  return true;
}

// offset < 0: sequential write at the descriptor's position.  Large spans to a regular file go through write_mapped;
// where that is not possible they are cut into pieces written with pwrite from a few threads.  Returns 0 or an errno
// value; *what (optional) names the call that failed.
inline int write_all_errno (int fd, const void *buf, size_t bytes, int64_t offset, const char **what = nullptr)
{
  const unsigned char *p = static_cast<const unsigned char *> (buf);
  constexpr size_t PARALLEL_MIN = 32u << 20;
  constexpr unsigned N_THREADS = 8;
  if (what) *what = "write";
  if (bytes >= PARALLEL_MIN) {
    const int64_t at = (offset >= 0) ? offset : (int64_t) lseek (fd, 0, SEEK_CUR);
    if (at >= 0 && write_mapped (fd, p, bytes, at)) {
      if (offset < 0 && lseek (fd, at + (int64_t) bytes, SEEK_SET) < 0) { if (what) *what = "lseek"; return errno ? errno : EIO; }
      return 0;
    }
    if (at >= 0) {
      int err[N_THREADS] = {};
      std::vector<std::thread> pool;
      const size_t piece = ((bytes / N_THREADS) + 4095) & ~(size_t) 4095;
      for (unsigned t = 0; t < N_THREADS; t++) {
        const size_t lo = (size_t) t * piece;
        if (lo >= bytes) break;
        const size_t len = std::min (piece, bytes - lo);
        pool.emplace_back ([=, &err] { err[t] = write_span (fd, p + lo, len, at + (int64_t) lo); });
      }
      for (auto &th : pool) th.join ();
      for (unsigned t = 0; t < N_THREADS; t++)
        if (err[t]) return err[t];
      if (offset < 0 && lseek (fd, at + (int64_t) bytes, SEEK_SET) < 0) { if (what) *what = "lseek"; return errno ? errno : EIO; }
      return 0;
    }
  }
  return write_span (fd, p, bytes, offset);
}

}  // namespace fileio
}  // namespace gt4gpu
