// gt4gpu_api.cu -- the C ABI declared in include/gt4gpu.h: context, containers (the gt4gpu
// loader that turns 12-byte AoS records into SoA arrays in HBM), the merge entry points that
// replace compare_wordmaps / union_multi / intersect_multi / gt4_write_union, result handling
// and the host-side key-range sharding plan.  Host language: C++ compiled by nvcc, exported as
// extern "C".  There is no CPU implementation of the merges in here: every compute entry point
// fails with GT4GPU_ERR_CUDA when no device is available.
#include <cuda_runtime.h>

#include <errno.h>
#include <fcntl.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/statvfs.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/gt4gpu.h"
#include "gt4gpu_fileio.h"
#include "gt4gpu_internal.h"

using namespace gt4gpu;

static_assert (sizeof (gt4gpu_header) == 48, "GT4ListHeader is 48 bytes (src/word-list.h:61-72)");

struct gt4gpu_list {
  uint64_t *words;
  uint32_t *counts;
  uint64_t n_words;
  uint64_t sum_counts;
  uint32_t word_length;
  int owned;
};

namespace gt4gpu {

int debug_flags ()
{
  const char *env = getenv ("GT4GPU_DEBUG");
  if (!env) return 0;
#ifdef GT4GPU_UNSAFE_EXPERIMENTS
  return atoi (env);
#else
  return atoi (env) & (2 | 4 | 32);
#endif
}

}  // namespace gt4gpu

namespace {

constexpr uint32_t LIST_CODE = (uint32_t) ('G' << 24 | 'T' << 16 | '4' << 8 | 'C');   // src/word-list.c:31
constexpr uint32_t INDEX_CODE = (uint32_t) ('G' << 24 | 'T' << 16 | '4' << 8 | 'I');  // src/index-map.c:38
constexpr uint64_t STAGE_RECORDS = 32ull << 20;   // records per AoS staging chunk (384 MiB)

struct Context {
  bool ready = false;
  int device = -1;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  cudaStream_t in_stream = nullptr;    // host-to-device copies of the pipelined host path
  cudaStream_t out_stream = nullptr;   // device-to-host copies of the pipelined host path
  TileShape shape = {256, 9};   // tile shape of the multi-output kernel (setop2_tile_kernel)
  int stream_consumers = 512;   // consumer threads per CTA of the single-output kernel (setop2_stream_kernel)
  int stream_items = 9;         // its merged items per thread
  int use_stream = 1;           // 0: run single-output merges through setop2_tile_kernel too
  int use_fused = 1;            // 0: several outputs of one merge take one pass of the single-output kernel each
  int use_kway = 1;             // 0: N-list calls go through the tree / chain of two-list merges; 1: unions take the single pass; 2: intersections too
  int sm_count = 0;
};
Context g_ctx;
std::mutex g_init_mutex;      // gt4gpu_init / gt4gpu_shutdown; the merge entry points only read the context

thread_local char tl_error[512] = "";
thread_local float tl_ms_partition = 0.f, tl_ms_merge = 0.f;
thread_local uint32_t tl_launches = 0;
thread_local cudaEvent_t tl_ev[3] = {nullptr, nullptr, nullptr};

int fail (int code, const char *fmt, ...)
{
  va_list ap;
  va_start (ap, fmt);
  vsnprintf (tl_error, sizeof (tl_error), fmt, ap);
  va_end (ap);
  return code;
}

#define CU(expr)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (expr);                                                                       \
    if (e_ != cudaSuccess) return fail (GT4GPU_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString (e_)); \
  } while (0)

int ensure_ready ()
{
  if (g_ctx.ready) return 0;
  return gt4gpu_init (-1);
}

// Pinned bounce buffers for the file <-> device paths.  cudaMallocHost costs tens of milliseconds per call (it pins pages),
// so the buffers are kept and handed out again; concurrent callers each get their own.
constexpr size_t BOUNCE_BYTES = (8ull << 20) * 12;      // 8 Mi records
struct PinnedPool {
  std::mutex mutex;
  std::vector<void *> free_list;
  void *take ()
  {
    {
      std::lock_guard<std::mutex> lock (mutex);
      if (!free_list.empty ()) {
        void *p = free_list.back ();
        free_list.pop_back ();
        return p;
      }
    }
    void *p = nullptr;
    if (cudaMallocHost (&p, BOUNCE_BYTES) != cudaSuccess) { cudaGetLastError (); return nullptr; }
    return p;
  }
  void give (void *p)
  {
    if (!p) return;
    std::lock_guard<std::mutex> lock (mutex);
    free_list.push_back (p);
  }
  void release_all ()
  {
    std::lock_guard<std::mutex> lock (mutex);
    for (void *p : free_list) cudaFreeHost (p);
    free_list.clear ();
  }
};
PinnedPool g_pinned;

int dev_alloc (void **p, size_t bytes)
{
  CU (cudaMallocAsync (p, bytes ? bytes : 16, g_ctx.stream));
  return 0;
}

void dev_free (void *p)
{
  if (p) cudaFreeAsync (p, g_ctx.stream);
}

// One output stream of an internal merge.
struct MergeOut {
  uint64_t *words = nullptr;
  uint32_t *counts = nullptr;
  uint64_t capacity = 0;
  uint64_t n = 0;
  uint64_t sum = 0;
  bool caller = false;
};

struct DevList {
  const uint64_t *words;
  const uint32_t *counts;
  uint64_t n;
};

uint64_t worst_case (const SetOpParams &p, int stream, uint64_t na, uint64_t nb)
{
  if (p.sem == SEM_NISECT_PARTIAL || p.sem == SEM_NISECT_FINAL) return std::min (na, nb);
  if (p.sem != SEM_PAIR) return na + nb;
  switch (stream) {
  case 0: return na + nb;
  case 1: return std::min (na, nb);
  case 2: return na;
  default: return nb;
  }
}

// The engine: partition + tile kernel over two device-resident SoA lists.
int merge2_device (const DevList &a, const DevList &b, const SetOpParams &p, uint32_t stream_mask, bool countonly, MergeOut out[4])
{
  const uint64_t total = a.n + b.n;
  const int n_req = __builtin_popcount (stream_mask);
  if (n_req == 0) return fail (GT4GPU_ERR_ARG, "no output stream requested");
  // Several outputs: one pass of the single-output kernel per output beats the fused multi-output kernel by ~2x
  // (measured, profiles/README.md), so the fused kernel is only used when the stream kernel is switched off.
  const TileShape shape = g_ctx.shape;
  const bool use_stream = g_ctx.use_stream != 0;
  // several outputs: ONE pass of the fused kernel (one read of the lists); -du and count-only runs take one pass per output
  const bool use_fused = use_stream && g_ctx.use_fused && !countonly && fused_applicable (p, stream_mask);
  const int ns = (n_req == 1 || use_stream) ? 1 : 4;
  const uint64_t tile = use_fused ? (uint64_t) fused_tile_slots (stream_mask)
                      : use_stream ? (uint64_t) g_ctx.stream_consumers * g_ctx.stream_items : (uint64_t) shape.threads * shape.items;
  const uint64_t n_tiles = (total + tile - 1) / tile;
  cudaStream_t st = g_ctx.stream;

  for (int s = 0; s < 4; s++) {
    if (!((stream_mask >> s) & 1u)) continue;
    out[s].n = out[s].sum = 0;
    if (countonly) continue;
    if (!out[s].caller) {
      out[s].capacity = worst_case (p, s, a.n, b.n);
      int rc = dev_alloc ((void **) &out[s].words, out[s].capacity * sizeof (uint64_t));
      if (!rc) rc = dev_alloc ((void **) &out[s].counts, out[s].capacity * sizeof (uint32_t));
      if (rc) {           // hand back what this call allocated so far; the caller's buffers stay untouched
        for (int q = 0; q <= s; q++) if (((stream_mask >> q) & 1u) && !out[q].caller) { dev_free (out[q].words); dev_free (out[q].counts); out[q].words = nullptr; out[q].counts = nullptr; }
        return rc;
      }
    }
  }
  if (total == 0) return 0;

  // scratch: [CallHeader | descriptors | partition]
  const size_t hdr_bytes = (sizeof (CallHeader) + 255) & ~(size_t) 255;
  const size_t desc_bytes = countonly ? 0 : use_fused ? fused_desc_bytes (n_tiles) : (size_t) ns * n_tiles * sizeof (uint64_t);
  const size_t part_bytes = (n_tiles + 1) * sizeof (uint64_t);
  unsigned char *ws = nullptr;
  int rc = dev_alloc ((void **) &ws, hdr_bytes + desc_bytes + part_bytes);
  if (rc) return rc;
  struct ScratchGuard {      // the scratch goes back to the pool on every exit path
    unsigned char *&p;
    ~ScratchGuard () { if (p) { dev_free (p); p = nullptr; } }
  } guard{ws};
  CU (cudaMemsetAsync (ws, 0, hdr_bytes + desc_bytes, st));

  TileArgs args;
  memset (&args, 0, sizeof (args));
  args.a_words = a.words; args.a_counts = a.counts; args.na = a.n;
  args.b_words = b.words; args.b_counts = b.counts; args.nb = b.n;
  args.hdr = reinterpret_cast<CallHeader *> (ws);
  args.desc = reinterpret_cast<uint64_t *> (ws + hdr_bytes);
  uint64_t *part = reinterpret_cast<uint64_t *> (ws + hdr_bytes + desc_bytes);
  args.part = part;
  args.n_tiles = n_tiles;
  args.p = p;
  args.p.ops = stream_mask;
  args.stream0 = __builtin_ctz (stream_mask);
  args.debug = debug_flags ();
  for (int s = 0; s < 4; s++) {
    args.out_words[s] = out[s].words;
    args.out_counts[s] = out[s].counts;
    args.out_capacity[s] = out[s].capacity;
  }

  for (int i = 0; i < 3; i++) if (!tl_ev[i]) CU (cudaEventCreate (&tl_ev[i]));
  CU (cudaEventRecord (tl_ev[0], st));
  CU (launch_partition (a.words, a.n, b.words, b.n, (uint32_t) tile, n_tiles, part, st));
  CU (cudaEventRecord (tl_ev[1], st));
  uint32_t n_launches = 1;
  if (use_fused) {
    CU (launch_setop2_fused (args, g_ctx.sm_count, st));
    n_launches += 1;
  } else if (use_stream) {
    bool first = true;
    for (int s = 0; s < 4; s++) {
      if (!((stream_mask >> s) & 1u)) continue;
      if (!first) {       // same co-ranks, fresh ticket and look-back descriptors
        CU (cudaMemsetAsync (&args.hdr->ticket, 0, sizeof (uint32_t), st));
        if (desc_bytes) CU (cudaMemsetAsync (args.desc, 0, desc_bytes, st));
      }
      first = false;
      args.stream0 = s;
      args.p.ops = 1u << s;
      // Sparse outputs merge 12 % faster through the side-buffer variant of the kernel, dense ones 20 % slower, and the
      // density is a property of the data: count the survivors of every 64th tile first (1.6 % of the merge's time)
      // when the call is big enough for that to pay.
      args.side_hint = 0;
      constexpr uint64_t SAMPLE_STRIDE = 64, SAMPLE_MIN_TILES = 8192;
      if (!countonly && n_tiles >= SAMPLE_MIN_TILES && stream_side_capable (args.p, s, g_ctx.stream_consumers, g_ctx.stream_items)) {
        CallHeader *h2 = nullptr;
        const size_t h2_bytes = (sizeof (CallHeader) + 255) & ~(size_t) 255;
        int rc2 = dev_alloc ((void **) &h2, h2_bytes);
        if (rc2) return rc2;
        struct H2Guard { CallHeader *p; ~H2Guard () { dev_free (p); } } h2_guard{h2};
        CU (cudaMemsetAsync (h2, 0, h2_bytes, st));
        TileArgs sample = args;
        sample.hdr = h2;
        sample.tile_stride = (uint32_t) SAMPLE_STRIDE;
        CU (launch_setop2_stream (sample, g_ctx.stream_consumers, g_ctx.stream_items, true, g_ctx.sm_count, st));
        n_launches += 1;
        CallHeader hs;
        CU (cudaMemcpyAsync (&hs, h2, sizeof (hs), cudaMemcpyDeviceToHost, st));
        CU (cudaStreamSynchronize (st));
        uint64_t kept = 0;
        for (int k = 0; k < TOTAL_SLOTS; k++) kept += hs.totals[s][k][0];
        const uint64_t sampled_slots = ((n_tiles + SAMPLE_STRIDE - 1) / SAMPLE_STRIDE) * tile;
        // (a side buffer holds a third of a tile's slots in the three-stage variant, 61 % in the two-stage one)
        args.side_hint = (double) kept <= 0.27 * (double) sampled_slots ? 1 : (double) kept <= 0.56 * (double) sampled_slots ? 2 : 0;
      }
      CU (launch_setop2_stream (args, g_ctx.stream_consumers, g_ctx.stream_items, countonly, g_ctx.sm_count, st));
      n_launches += 1;
    }
  } else {
    CU (launch_setop2 (args, shape, ns, countonly, st));
    n_launches += 1;
  }
  CU (cudaEventRecord (tl_ev[2], st));

  CallHeader h;
  CU (cudaMemcpyAsync (&h, ws, sizeof (h), cudaMemcpyDeviceToHost, st));
  CU (cudaStreamSynchronize (st));
  float ms = 0.f;
  CU (cudaEventElapsedTime (&ms, tl_ev[0], tl_ev[1]));
  tl_ms_partition += ms;
  CU (cudaEventElapsedTime (&ms, tl_ev[1], tl_ev[2]));
  tl_ms_merge += ms;
  tl_launches += n_launches;

  // statistics of the LAST pass only make sense with one switch at a time (bits 2 and 32 share dbg[0..1])
  if ((args.debug & 2) && !(args.debug & 32) && h.dbg[3])
    fprintf (stderr, "gt4gpu debug: look-backs %llu, mean %.0f cycles, %.2f polls, %.2f extra hops\n", h.dbg[3],
             (double) h.dbg[0] / h.dbg[3], (double) h.dbg[1] / h.dbg[3], (double) h.dbg[2] / h.dbg[3] - 1.0);
  if ((args.debug & 32) && h.dbg[7])
    fprintf (stderr, "gt4gpu debug: consumer warp cycles per tile: wait %.0f search %.0f merge %.0f scan+barrier %.0f scatter %.0f (warp-tiles %llu); "
                     "splitter warp per tile: waits for the TMA %.0f, searches %.0f\n",
             (double) h.dbg[0] / h.dbg[7], (double) h.dbg[1] / h.dbg[7], (double) h.dbg[4] / h.dbg[7], (double) h.dbg[5] / h.dbg[7],
             (double) h.dbg[6] / h.dbg[7], h.dbg[7], (double) h.dbg[2] / n_tiles, (double) h.dbg[3] / n_tiles);
  if (h.overflow == 2u) return fail (GT4GPU_ERR_ARG, "input lists are not strictly ascending (the merge result is undefined)");
  if (h.overflow) return fail (GT4GPU_ERR_CAPACITY, "output buffer too small for the merge result");
  for (int s = 0; s < 4; s++) {
    if (!((stream_mask >> s) & 1u)) continue;
    for (int k = 0; k < TOTAL_SLOTS; k++) {
      out[s].n += h.totals[s][k][0];
      out[s].sum += h.totals[s][k][1];
    }
  }
  return 0;
}

void free_out (MergeOut &o)
{
  if (!o.caller) {
    dev_free (o.words);
    dev_free (o.counts);
  }
  o.words = nullptr;
  o.counts = nullptr;
}

void reset_timing ()
{
  tl_ms_partition = tl_ms_merge = 0.f;
  tl_launches = 0;
}

// Header acceptance, restated from src/word-map.c:179-215 (mode 0) and
// src/word-list-stream.c:150-168 (mode 1).
int parse_header (const unsigned char *file, uint64_t size, int mode, const char *path, gt4gpu_header *out)
{
  gt4gpu_header h;
  memset (&h, 0, sizeof (h));
  if (mode == 0) {
    if (size < 12) return fail (GT4GPU_ERR_FORMAT, "%s: file too small for a list header", path);
    memcpy (&h, file, 12);
    if (h.code != LIST_CODE) return fail (GT4GPU_ERR_FORMAT, "%s: invalid file tag (%x, should be %x)", path, h.code, LIST_CODE);
    if (h.version_major != GT4GPU_VERSION_MAJOR)
      return fail (GT4GPU_ERR_FORMAT, "%s: incompatible major version %u (required %u)", path, h.version_major, GT4GPU_VERSION_MAJOR);
    const size_t take = (h.version_minor <= 2) ? 40 : 48;
    if (size < take) return fail (GT4GPU_ERR_FORMAT, "%s: truncated header", path);
    memcpy (&h, file, take);
    if (h.version_minor == 0) h.list_start = 40;
    if (h.version_minor <= 2) { h.word_bytes = 8; h.count_bytes = 4; }
    // (division form: n_words comes from the file and must not be able to wrap the products)
    const uint64_t rec_bytes = (uint64_t) h.word_bytes + h.count_bytes;
    if (h.list_start > size || (rec_bytes && (size - h.list_start) / rec_bytes < h.n_words)) {
      const uint64_t need = h.list_start + h.n_words * rec_bytes;
      return fail (GT4GPU_ERR_FORMAT, "%s: file size too small (%llu, should be at least %llu)", path,
                   (unsigned long long) size, (unsigned long long) need);
    }
    // the accessors always stride 12 bytes (src/word-map.h:89-99); make sure that stays in bounds
    if ((size - h.list_start) / 12ull < h.n_words) return fail (GT4GPU_ERR_FORMAT, "%s: records exceed the file", path);
  } else {
    if (size < 48) return fail (GT4GPU_ERR_FORMAT, "%s: could not read list header", path);
    memcpy (&h, file, 48);
    if (h.code != LIST_CODE) return fail (GT4GPU_ERR_FORMAT, "%s: invalid file tag (%x, should be %x)", path, h.code, LIST_CODE);
    if (h.version_major > GT4GPU_VERSION_MAJOR)
      return fail (GT4GPU_ERR_FORMAT, "%s: incompatible major version %u (required %u)", path, h.version_major, GT4GPU_VERSION_MAJOR);
    if (h.version_major == 4 && h.version_minor == 0) h.list_start = 48;
    if (h.list_start > size || (size - h.list_start) / 12ull < h.n_words) return fail (GT4GPU_ERR_FORMAT, "%s: records exceed the file", path);
  }
  *out = h;
  return 0;
}

// struct _GT4IndexHeader, src/index-map.h:69-83
struct IndexHeader {
  uint32_t code, version_major, version_minor, word_length;
  uint64_t num_words, num_locations;
  uint32_t n_file_bits, n_subseq_bits, n_pos_bits, filler0;
  uint64_t files_start, kmers_start, locations_start;
};

// GT4I acceptance as in gt4_index_map_new (src/index-map.c:316-350); reported through a list header whose
// count_bytes = 8 marks the 16-byte record layout (word + location offset)
int parse_index_header (const unsigned char *file, uint64_t size, const char *path, gt4gpu_header *out, uint64_t *num_locations)
{
  IndexHeader ih;
  if (size < sizeof (ih)) return fail (GT4GPU_ERR_FORMAT, "%s: file too small for an index header", path);
  memcpy (&ih, file, sizeof (ih));
  if (ih.code != INDEX_CODE) return fail (GT4GPU_ERR_FORMAT, "%s: invalid file tag (%x, should be %x)", path, ih.code, INDEX_CODE);
  if (ih.version_major != GT4GPU_VERSION_MAJOR)
    return fail (GT4GPU_ERR_FORMAT, "%s: incompatible major version %u (required %u)", path, ih.version_major, GT4GPU_VERSION_MAJOR);
  if (ih.kmers_start > size || (size - ih.kmers_start) / 16 < ih.num_words) return fail (GT4GPU_ERR_FORMAT, "%s: k-mer table exceeds the file", path);
  memset (out, 0, sizeof (*out));
  out->code = ih.code;
  out->version_major = ih.version_major;
  out->version_minor = ih.version_minor;
  out->word_length = ih.word_length;
  out->n_words = ih.num_words;
  out->total_count = ih.num_locations;
  out->list_start = ih.kmers_start;
  out->word_bytes = 8;
  out->count_bytes = 8;
  if (num_locations) *num_locations = ih.num_locations;
  return 0;
}

bool is_index_file (const unsigned char *file, uint64_t size)
{
  uint32_t code = 0;
  if (size >= 4) memcpy (&code, file, 4);
  return code == INDEX_CODE;
}

// index records [first, first + n) of a mapped GT4I file -> device SoA, chunked; every chunk carries one extra record
// (the next word's offset) except the last one of the file, which is closed by num_locations
int upload_index (const unsigned char *kmers, uint64_t n_total, uint64_t num_locations, uint64_t first, uint64_t n, uint64_t *d_words, uint32_t *d_counts)
{
  if (n == 0) return 0;
  const uint64_t chunk = std::min<uint64_t> (n, 16ull << 20);
  void *stage = nullptr;
  int rc = dev_alloc (&stage, (chunk + 1) * 16);
  if (rc) return rc;
  cudaError_t e = cudaSuccess;
  for (uint64_t done = 0; done < n && e == cudaSuccess; done += chunk) {
    const uint64_t m = std::min (chunk, n - done);
    const int has_next = (first + done + m < n_total) ? 1 : 0;
    e = cudaMemcpyAsync (stage, kmers + (first + done) * 16, (m + has_next) * 16, cudaMemcpyHostToDevice, g_ctx.stream);
    if (e == cudaSuccess) e = launch_index16 (stage, m, has_next, num_locations, d_words + done, d_counts + done, g_ctx.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize (g_ctx.stream);     // one staging buffer, pageable source
  }
  dev_free (stage);
  if (e != cudaSuccess) return fail (GT4GPU_ERR_CUDA, "index upload: %s", cudaGetErrorString (e));
  return 0;
}

struct Mapping {
  const unsigned char *data = nullptr;
  uint64_t size = 0;
  ~Mapping () { if (data) munmap ((void *) data, size); }
};

int map_file (const char *path, Mapping &m)
{
  int fd = open (path, O_RDONLY);
  if (fd < 0) return fail (GT4GPU_ERR_IO, "cannot open %s: %s", path, strerror (errno));
  struct stat st;
  if (fstat (fd, &st) < 0) { close (fd); return fail (GT4GPU_ERR_IO, "cannot stat %s", path); }
  m.size = (uint64_t) st.st_size;
  if (m.size == 0) { close (fd); return fail (GT4GPU_ERR_FORMAT, "%s: empty file", path); }
  void *p = mmap (NULL, m.size, PROT_READ, MAP_PRIVATE, fd, 0);
  close (fd);
  if (p == MAP_FAILED) return fail (GT4GPU_ERR_IO, "cannot mmap %s: %s", path, strerror (errno));
  m.data = static_cast<const unsigned char *> (p);
  madvise (p, m.size, MADV_SEQUENTIAL);
  return 0;
}

int new_list (uint64_t n, uint32_t k, gt4gpu_list **out)
{
  gt4gpu_list *l = static_cast<gt4gpu_list *> (calloc (1, sizeof (gt4gpu_list)));
  if (!l) return fail (GT4GPU_ERR_ARG, "out of host memory");
  l->n_words = n;
  l->word_length = k;
  l->owned = 1;
  int rc = dev_alloc ((void **) &l->words, n * sizeof (uint64_t));
  if (!rc) rc = dev_alloc ((void **) &l->counts, n * sizeof (uint32_t));
  if (rc) { dev_free (l->words); free (l); return rc; }
  *out = l;
  return 0;
}

// parallel memcpy (page-cache-hot mmaps are limited by one core's page-fault + copy rate)
static void copy_parallel (void *dst, const void *src, size_t bytes, unsigned n_threads)
{
  if (bytes < (8u << 20) || n_threads < 2) { memcpy (dst, src, bytes); return; }
  std::vector<std::thread> pool;
  const size_t piece = ((bytes / n_threads) + 4095) & ~(size_t) 4095;
  for (unsigned t = 0; t < n_threads; t++) {
    const size_t lo = (size_t) t * piece;
    if (lo >= bytes) break;
    const size_t len = std::min (piece, bytes - lo);
    pool.emplace_back ([=] { memcpy (static_cast<unsigned char *> (dst) + lo, static_cast<const unsigned char *> (src) + lo, len); });
  }
  for (auto &th : pool) th.join ();
}

// host AoS -> device SoA, chunked through a device staging buffer.  Pinned sources are copied directly; pageable
// ones (the mmap of a list file) go through two pinned bounce buffers filled by a few threads, so the copy into
// pinned memory of chunk i+1 overlaps the H2D + de-interleave of chunk i.
int upload_aos (const void *records, uint64_t n, uint64_t *d_words, uint32_t *d_counts)
{
  if (n == 0) return 0;
  cudaPointerAttributes attr;
  bool pinned_src = false;
  if (cudaPointerGetAttributes (&attr, records) == cudaSuccess) pinned_src = (attr.type == cudaMemoryTypeHost);
  else cudaGetLastError ();
  const unsigned char *src = static_cast<const unsigned char *> (records);
  cudaError_t e = cudaSuccess;
  if (pinned_src || n < (1u << 20)) {
    const uint64_t chunk = std::min (n, STAGE_RECORDS);
    void *stage = nullptr;
    int rc = dev_alloc (&stage, chunk * 12);
    if (rc) return rc;
    for (uint64_t done = 0; done < n && e == cudaSuccess; done += chunk) {
      const uint64_t m = std::min (chunk, n - done);
      e = cudaMemcpyAsync (stage, src + done * 12, m * 12, cudaMemcpyHostToDevice, g_ctx.stream);
      if (e == cudaSuccess) e = launch_deinterleave (stage, m, d_words + done, d_counts + done, g_ctx.stream);
    }
    dev_free (stage);
    if (e == cudaSuccess) e = cudaStreamSynchronize (g_ctx.stream);
  } else {
    const uint64_t chunk = std::min<uint64_t> (n, 8ull << 20);        // 96 MiB bounce buffers
    void *stage[2] = {nullptr, nullptr}, *bounce[2] = {nullptr, nullptr};
    cudaEvent_t done_ev[2] = {nullptr, nullptr};
    int rc = 0;
    for (int b = 0; b < 2 && !rc; b++) {
      rc = dev_alloc (&stage[b], chunk * 12);
      if (!rc && !(bounce[b] = g_pinned.take ())) rc = fail (GT4GPU_ERR_CUDA, "cudaMallocHost failed");
      if (!rc && cudaEventCreateWithFlags (&done_ev[b], cudaEventDisableTiming) != cudaSuccess) rc = fail (GT4GPU_ERR_CUDA, "cudaEventCreate failed");
    }
    const unsigned n_threads = std::min (8u, std::max (2u, std::thread::hardware_concurrency () / 2));
    int b = 0;
    for (uint64_t done = 0; done < n && !rc && e == cudaSuccess; done += chunk, b ^= 1) {
      const uint64_t m = std::min (chunk, n - done);
      if (done >= 2 * chunk) e = cudaEventSynchronize (done_ev[b]);     // the H2D that last read this bounce buffer
      if (e != cudaSuccess) break;
      copy_parallel (bounce[b], src + done * 12, m * 12, n_threads);
      e = cudaMemcpyAsync (stage[b], bounce[b], m * 12, cudaMemcpyHostToDevice, g_ctx.stream);
      if (e == cudaSuccess) e = launch_deinterleave (stage[b], m, d_words + done, d_counts + done, g_ctx.stream);
      if (e == cudaSuccess) e = cudaEventRecord (done_ev[b], g_ctx.stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize (g_ctx.stream);
    for (int k = 0; k < 2; k++) {
      dev_free (stage[k]);
      g_pinned.give (bounce[k]);
      if (done_ev[k]) cudaEventDestroy (done_ev[k]);
    }
    if (rc) return rc;
  }
  if (e != cudaSuccess) return fail (GT4GPU_ERR_CUDA, "upload: %s", cudaGetErrorString (e));
  return 0;
}

// device SoA -> host AoS, chunked; sink (ptr, bytes, first_record) consumes each chunk when given,
// otherwise the chunks land contiguously in `records`
template <typename Sink>
int download_aos (const uint64_t *d_words, const uint32_t *d_counts, uint64_t n, void *records, Sink sink, bool use_sink)
{
  if (n == 0) return 0;
  cudaError_t e = cudaSuccess;
  int rc = 0;
  if (!use_sink) {
    const uint64_t chunk = std::min (n, STAGE_RECORDS);
    void *stage = nullptr;
    rc = dev_alloc (&stage, chunk * 12);
    if (rc) return rc;
    for (uint64_t done = 0; done < n && e == cudaSuccess; done += chunk) {
      const uint64_t m = std::min (chunk, n - done);
      e = launch_interleave (d_words + done, d_counts + done, m, stage, g_ctx.stream);
      if (e == cudaSuccess) e = cudaMemcpyAsync (static_cast<unsigned char *> (records) + done * 12, stage, m * 12, cudaMemcpyDeviceToHost, g_ctx.stream);
    }
    dev_free (stage);
    if (e == cudaSuccess) e = cudaStreamSynchronize (g_ctx.stream);
    if (e != cudaSuccess) return fail (GT4GPU_ERR_CUDA, "download: %s", cudaGetErrorString (e));
    return 0;
  }
  // with a sink (file writes): chunk i+1 is interleaved and copied out while the sink consumes chunk i
  const uint64_t chunk = std::min<uint64_t> (n, 8ull << 20);
  void *stage[2] = {nullptr, nullptr}, *pinned[2] = {nullptr, nullptr};
  cudaEvent_t ready[2] = {nullptr, nullptr};
  for (int b = 0; b < 2 && !rc; b++) {
    rc = dev_alloc (&stage[b], chunk * 12);
    if (!rc && !(pinned[b] = g_pinned.take ())) rc = fail (GT4GPU_ERR_CUDA, "cudaMallocHost failed");
    if (!rc && cudaEventCreateWithFlags (&ready[b], cudaEventDisableTiming) != cudaSuccess) rc = fail (GT4GPU_ERR_CUDA, "cudaEventCreate failed");
  }
  auto enqueue = [&] (uint64_t done, int b) {
    const uint64_t m = std::min (chunk, n - done);
    e = launch_interleave (d_words + done, d_counts + done, m, stage[b], g_ctx.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync (pinned[b], stage[b], m * 12, cudaMemcpyDeviceToHost, g_ctx.stream);
    if (e == cudaSuccess) e = cudaEventRecord (ready[b], g_ctx.stream);
  };
  if (!rc) enqueue (0, 0);
  int b = 0;
  for (uint64_t done = 0; done < n && !rc && e == cudaSuccess; done += chunk, b ^= 1) {
    if (done + chunk < n) enqueue (done + chunk, b ^ 1);        // its bounce buffer was consumed by the sink two rounds ago
    if (e != cudaSuccess) break;
    e = cudaEventSynchronize (ready[b]);
    if (e != cudaSuccess) break;
    rc = sink (pinned[b], std::min (chunk, n - done) * 12, done);
  }
  cudaStreamSynchronize (g_ctx.stream);
  for (int k = 0; k < 2; k++) {
    dev_free (stage[k]);
    g_pinned.give (pinned[k]);
    if (ready[k]) cudaEventDestroy (ready[k]);
  }
  if (!rc && e != cudaSuccess) rc = fail (GT4GPU_ERR_CUDA, "download: %s", cudaGetErrorString (e));
  return rc;
}

// file output: gt4gpu_fileio.h (write_span, write_mapped, write_all_errno), shared with the CPU tests
int write_all (int fd, const void *buf, size_t bytes, int64_t offset)
{
  const char *what = "write";
  const int e = fileio::write_all_errno (fd, buf, bytes, offset, &what);
  if (e) return fail (GT4GPU_ERR_IO, "%s failed: %s", what, strerror (e));
  return 0;
}

void fill_result (gt4gpu_result *r, const MergeOut &o, uint32_t k, bool countonly)
{
  r->n_words = o.n;
  r->total_count = o.sum;
  r->word_length = k;
  if (countonly) {
    r->flags |= GT4GPU_RESULT_COUNT_ONLY;
    return;
  }
  r->words = o.words;
  r->counts = o.counts;
  r->capacity = o.capacity;
}

SetOpParams nlist_params (int sem, int rule, uint32_t cutoff, uint32_t ov)
{
  SetOpParams p;
  memset (&p, 0, sizeof (p));
  p.ops = OP_UNION;
  p.cutoff = cutoff;
  p.count_override = ov;
  p.rule[0] = p.rule[1] = p.rule[2] = p.rule[3] = rule;
  p.sem = sem;
  return p;
}

// N-list union as a balanced tree of two-list merges (add / max / number are associative and
// commutative, so any tree reproduces union_multi's left-to-right fold, u32 wrap-around included);
// inner nodes keep every key, the cut-off is applied by the root only.
struct TreeNode {
  DevList list;
  MergeOut owned;      // valid when is_owned: an intermediate result this node must free
  bool is_owned;
};

int union_tree (const std::vector<DevList> &leaves, int rule, uint32_t cutoff, uint32_t ov, bool countonly, MergeOut *root)
{
  std::vector<TreeNode> level;
  for (const DevList &l : leaves) level.push_back (TreeNode{l, MergeOut (), false});
  while (level.size () < 2) level.push_back (TreeNode{DevList{nullptr, nullptr, 0}, MergeOut (), false});
  auto release = [] (TreeNode &n) { if (n.is_owned) { free_out (n.owned); n.is_owned = false; } };
  while (level.size () > 2) {
    std::vector<TreeNode> next;
    for (size_t i = 0; i + 1 < level.size (); i += 2) {
      MergeOut out[4];
      SetOpParams p = nlist_params (SEM_NUNION_PARTIAL, rule, cutoff, ov);
      int rc = merge2_device (level[i].list, level[i + 1].list, p, OP_UNION, false, out);
      release (level[i]);
      release (level[i + 1]);
      if (rc) {
        free_out (out[0]);
        for (size_t k = i + 2; k < level.size (); k++) release (level[k]);
        for (auto &n : next) release (n);
        return rc;
      }
      next.push_back (TreeNode{DevList{out[0].words, out[0].counts, out[0].n}, out[0], true});
    }
    if (level.size () & 1) next.push_back (level.back ());   // odd one out moves up unchanged
    level.swap (next);
  }
  MergeOut out[4];
  if (root->caller) out[0] = *root;
  SetOpParams p = nlist_params (SEM_NUNION_FINAL, rule, cutoff, ov);
  int rc = merge2_device (level[0].list, level[1].list, p, OP_UNION, countonly, out);
  release (level[0]);
  release (level[1]);
  if (rc) { free_out (out[0]); return rc; }
  *root = out[0];
  return 0;
}


// ---- single-pass N-list union / intersection (gt4gpu_kway_kernel.cu) ----------------------------------------------

enum { KWAY_UNION = 0, KWAY_ISECT = 1 };

// One pass over 1..KWAY_MAX_LISTS device lists.  *fallback is set (and nothing else returned) when the pass does not
// apply: list arrays that are not 16-byte aligned, or a tile beyond the capacity the sampling bound covers.
int kway_pass (const std::vector<DevList> &lists, int op, int rule, uint32_t cutoff, uint32_t ov, bool final_pass, bool countonly,
               MergeOut *out, bool *fallback)
{
  *fallback = false;
  const int n_lists = (int) lists.size ();
  if (n_lists < 1 || n_lists > KWAY_MAX_LISTS) return fail (GT4GPU_ERR_ARG, "kway_pass: %d lists", n_lists);
  uint64_t total = 0, smallest = UINT64_MAX, n_samples = 0;
  KwayArgs args;
  memset (&args, 0, sizeof (args));
  for (int j = 0; j < n_lists; j++) {
    const DevList &l = lists[j];
    if (l.n && ((((uintptr_t) l.words) | ((uintptr_t) l.counts)) & 15u)) { *fallback = true; return 0; }
    args.words[j] = l.words; args.counts[j] = l.counts; args.n[j] = l.n;
    args.sample_off[j] = n_samples;
    n_samples += l.n / KWAY_SAMPLE;
    total += l.n;
    smallest = std::min (smallest, l.n);
  }
  const int nl = n_lists <= 4 ? 4 : 8;
  // samples per tile: a tile holds every * KWAY_SAMPLE records on average and at most (every + 2 n_lists) * KWAY_SAMPLE; lists
  // whose words interleave evenly stay within +- n_lists / 2 samples of the average, anything beyond the capacity is
  // caught by the kernel (overflow 3) and the call falls back to the tree of two-list merges
  const uint64_t every = (uint64_t) (KWAY_TILE_CAP / KWAY_SAMPLE - std::max (2, n_lists / 2) - 2);
  const uint64_t n_tiles = n_samples ? (n_samples - 1) / every + 1 : 1;
  cudaStream_t st = g_ctx.stream;

  out->n = out->sum = 0;
  const uint64_t worst = (op == KWAY_ISECT) ? smallest : total;
  const bool own_out = !countonly && !out->caller;
  if (own_out) {
    out->capacity = worst;
    out->words = nullptr; out->counts = nullptr;
    int rc = dev_alloc ((void **) &out->words, worst * sizeof (uint64_t));
    if (!rc) rc = dev_alloc ((void **) &out->counts, worst * sizeof (uint32_t));
    if (rc) { free_out (*out); return rc; }
  }
  if (total == 0 || (op == KWAY_ISECT && smallest == 0)) return 0;

  // scratch: [CallHeader | descriptors | cuts | bounds | samples | sort buffer | sort scratch]
  const size_t hdr_bytes = (sizeof (CallHeader) + 255) & ~(size_t) 255;
  const size_t desc_bytes = (countonly ? 0 : (size_t) n_tiles * sizeof (uint64_t) + 255) & ~(size_t) 255;
  const size_t cuts_bytes = ((size_t) (n_tiles + 1) * nl * sizeof (uint64_t) + 255) & ~(size_t) 255;
  const size_t bounds_bytes = ((size_t) (n_tiles + 1) * sizeof (uint64_t) + 255) & ~(size_t) 255;
  const size_t smp_bytes = ((size_t) n_samples * sizeof (uint64_t) + 255) & ~(size_t) 255;
  const size_t sort_bytes = n_samples ? sort_scratch_bytes (n_samples) : 0;
  unsigned char *ws = nullptr;
  int rc = dev_alloc ((void **) &ws, hdr_bytes + desc_bytes + cuts_bytes + bounds_bytes + 2 * smp_bytes + sort_bytes);
  if (rc) { if (own_out) free_out (*out); return rc; }
  struct ScratchGuard {
    unsigned char *&p;
    ~ScratchGuard () { if (p) { dev_free (p); p = nullptr; } }
  } guard{ws};
  auto bail = [&] (cudaError_t e, const char *what) {
    if (own_out) free_out (*out);
    return fail (GT4GPU_ERR_CUDA, "%s: %s", what, cudaGetErrorString (e));
  };
  cudaError_t e = cudaMemsetAsync (ws, 0, hdr_bytes + desc_bytes, st);
  if (e != cudaSuccess) return bail (e, "kway scratch");
  uint64_t *cuts = reinterpret_cast<uint64_t *> (ws + hdr_bytes + desc_bytes);
  uint64_t *bounds = reinterpret_cast<uint64_t *> (ws + hdr_bytes + desc_bytes + cuts_bytes);
  uint64_t *smp = reinterpret_cast<uint64_t *> (ws + hdr_bytes + desc_bytes + cuts_bytes + bounds_bytes);
  uint64_t *smp_alt = reinterpret_cast<uint64_t *> (ws + hdr_bytes + desc_bytes + cuts_bytes + bounds_bytes + smp_bytes);
  unsigned char *sort_ws = ws + hdr_bytes + desc_bytes + cuts_bytes + bounds_bytes + 2 * smp_bytes;

  args.n_lists = n_lists;
  args.n_real = n_lists;
  args.cuts = cuts;
  args.bounds = bounds;
  args.n_tiles = n_tiles;
  args.out_words = out->words;
  args.out_counts = out->counts;
  args.out_capacity = countonly ? 0 : out->capacity;
  args.hdr = reinterpret_cast<CallHeader *> (ws);
  args.desc = reinterpret_cast<uint64_t *> (ws + hdr_bytes);
  args.op = op;
  args.rule = rule;
  args.cutoff = cutoff;
  args.count_override = ov;
  args.final_pass = final_pass ? 1 : 0;
  args.debug = debug_flags ();

  for (int i = 0; i < 3; i++) if (!tl_ev[i]) { e = cudaEventCreate (&tl_ev[i]); if (e != cudaSuccess) return bail (e, "cudaEventCreate"); }
  cudaEventRecord (tl_ev[0], st);
  uint64_t *sorted = smp;
  uint32_t n_launches = 2;
  if (n_samples) {
    e = launch_kway_samples (args, n_samples, smp, st);
    // all 64 key bits: the boundaries must be in key order whatever the lists' word length says
    if (e == cudaSuccess) e = launch_radix_sort (smp, smp, smp_alt, n_samples, SORT_MAX_PASSES, sort_ws, g_ctx.sm_count, &sorted, st);
    if (e != cudaSuccess) return bail (e, "kway samples");
    n_launches += 3 + SORT_MAX_PASSES;
  }
  e = launch_kway_cuts (args, sorted, every, nl, cuts, bounds, st);
  if (e != cudaSuccess) return bail (e, "kway cuts");
  cudaEventRecord (tl_ev[1], st);
  e = launch_kway_tiles (args, nl, countonly, g_ctx.sm_count, st);
  if (e != cudaSuccess) return bail (e, "kway tiles");
  cudaEventRecord (tl_ev[2], st);
  CallHeader h;
  e = cudaMemcpyAsync (&h, ws, sizeof (h), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize (st);
  if (e != cudaSuccess) return bail (e, "kway pass");
  float ms = 0.f;
  cudaEventElapsedTime (&ms, tl_ev[0], tl_ev[1]);
  tl_ms_partition += ms;
  cudaEventElapsedTime (&ms, tl_ev[1], tl_ev[2]);
  tl_ms_merge += ms;
  tl_launches += n_launches;
  if ((args.debug & 32) && h.dbg[7])
    fprintf (stderr, "gt4gpu debug: k-way consumer warp cycles per tile: wait %.0f tables %.0f merge %.0f scan+barrier %.0f compaction %.0f (warp-tiles %llu, tiles %llu)\n",
             (double) h.dbg[0] / h.dbg[7], (double) h.dbg[1] / h.dbg[7], (double) h.dbg[4] / h.dbg[7], (double) h.dbg[5] / h.dbg[7],
             (double) h.dbg[6] / h.dbg[7], h.dbg[7], (unsigned long long) n_tiles);
  if (h.overflow) {
    if (own_out) free_out (*out);
    if (h.overflow == 3u) { *fallback = true; return 0; }
    if (h.overflow == 2u) return fail (GT4GPU_ERR_ARG, "input lists are not strictly ascending (the merge result is undefined)");
    return fail (GT4GPU_ERR_CAPACITY, "output buffer too small for the merge result");
  }
  for (int k = 0; k < TOTAL_SLOTS; k++) {
    out->n += h.totals[0][k][0];
    out->sum += h.totals[0][k][1];
  }
  return 0;
}

// N-list union through single passes over groups of up to KWAY_MAX_LISTS lists (add / max / number are associative and
// commutative, so any grouping reproduces union_multi's fold); inner passes keep every word, the last one applies the
// cut-off.  *fallback: the caller runs the tree of two-list merges instead.
int union_kway (const std::vector<DevList> &leaves, int rule, uint32_t cutoff, uint32_t ov, bool countonly, MergeOut *root, bool *fallback)
{
  std::vector<TreeNode> level;
  for (const DevList &l : leaves) level.push_back (TreeNode{l, MergeOut (), false});
  auto release_all = [] (std::vector<TreeNode> &v) { for (auto &n : v) if (n.is_owned) { free_out (n.owned); n.is_owned = false; } };
  while (level.size () > (size_t) KWAY_MAX_LISTS) {
    const size_t n_groups = (level.size () + KWAY_MAX_LISTS - 1) / KWAY_MAX_LISTS;
    std::vector<TreeNode> next;
    size_t at = 0;
    for (size_t g = 0; g < n_groups; g++) {
      const size_t take = (level.size () - at + (n_groups - g) - 1) / (n_groups - g);      // near-equal groups
      if (take == 1) {
        next.push_back (level[at]);
        level[at].is_owned = false;
        at += 1;
        continue;
      }
      std::vector<DevList> group;
      for (size_t k = 0; k < take; k++) group.push_back (level[at + k].list);
      MergeOut part;
      int rc = kway_pass (group, KWAY_UNION, rule, cutoff, ov, false, false, &part, fallback);
      if (rc || *fallback) { release_all (level); release_all (next); return rc; }
      for (size_t k = 0; k < take; k++) if (level[at + k].is_owned) { free_out (level[at + k].owned); level[at + k].is_owned = false; }
      next.push_back (TreeNode{DevList{part.words, part.counts, part.n}, part, true});
      at += take;
    }
    level.swap (next);
  }
  std::vector<DevList> last;
  for (auto &n : level) last.push_back (n.list);
  MergeOut out;
  if (root->caller) out = *root;
  int rc = kway_pass (last, KWAY_UNION, rule, cutoff, ov, true, countonly, &out, fallback);
  release_all (level);
  if (rc || *fallback) return rc;
  *root = out;
  return 0;
}

// N-list intersection: the reference folds the counts list by list (glistcompare.c:668-677) and rule min's "!freq ||"
// guard makes that fold order-dependent, so many lists run as a chain of passes whose first list is the running result.
int intersect_kway (const std::vector<DevList> &lists, int rule, uint32_t cutoff, uint32_t ov, bool countonly, MergeOut *root, bool *fallback)
{
  MergeOut acc;
  bool have_acc = false;
  size_t at = 0;
  while (at < lists.size ()) {
    std::vector<DevList> group;
    if (have_acc) group.push_back (DevList{acc.words, acc.counts, acc.n});
    while (at < lists.size () && group.size () < (size_t) KWAY_MAX_LISTS) group.push_back (lists[at++]);
    const bool last = at == lists.size ();
    MergeOut out;
    if (last && root->caller) out = *root;
    int rc = kway_pass (group, KWAY_ISECT, rule, cutoff, ov, last, last && countonly, &out, fallback);
    if (have_acc) free_out (acc);
    if (rc || *fallback) return rc;
    acc = out;
    have_acc = true;
  }
  *root = acc;
  return 0;
}

}  // namespace

// ============================================================================ ABI

extern "C" {

void gt4gpu_header_init (gt4gpu_header *hdr, uint32_t word_length)
{
  memset (hdr, 0, sizeof (*hdr));
  hdr->code = LIST_CODE;
  hdr->version_major = GT4GPU_VERSION_MAJOR;
  hdr->version_minor = GT4GPU_VERSION_MINOR;
  hdr->word_length = word_length;
  hdr->list_start = sizeof (gt4gpu_header);
  hdr->word_bytes = 8;
  hdr->count_bytes = 4;
}

int gt4gpu_device_count (void)
{
  int count = 0;
  if (cudaGetDeviceCount (&count) != cudaSuccess) { cudaGetLastError (); return 0; }
  return count;
}

int gt4gpu_init (int device)
{
  std::lock_guard<std::mutex> lock (g_init_mutex);
  int count = 0;
  cudaError_t e = cudaGetDeviceCount (&count);
  if (e != cudaSuccess || count == 0)
    return fail (GT4GPU_ERR_CUDA, "no CUDA device available (%s); libgt4gpu has no CPU fallback",
                 e != cudaSuccess ? cudaGetErrorString (e) : "device count is 0");
  if (device >= 0) CU (cudaSetDevice (device));
  CU (cudaGetDevice (&g_ctx.device));
  if (!g_ctx.own_stream) CU (cudaStreamCreateWithFlags (&g_ctx.own_stream, cudaStreamNonBlocking));
  if (!g_ctx.stream) g_ctx.stream = g_ctx.own_stream;
  if (!g_ctx.in_stream) CU (cudaStreamCreateWithFlags (&g_ctx.in_stream, cudaStreamNonBlocking));
  if (!g_ctx.out_stream) CU (cudaStreamCreateWithFlags (&g_ctx.out_stream, cudaStreamNonBlocking));
  // keep freed blocks cached in the stream-ordered pool: per-call scratch and results then cost no driver call
  cudaMemPool_t pool;
  CU (cudaDeviceGetDefaultMemPool (&pool, g_ctx.device));
  uint64_t keep = UINT64_MAX;
  CU (cudaMemPoolSetAttribute (pool, cudaMemPoolAttrReleaseThreshold, &keep));
  CU (cudaDeviceGetAttribute (&g_ctx.sm_count, cudaDevAttrMultiProcessorCount, g_ctx.device));
  const char *env = getenv ("GT4GPU_STREAM_SHAPE");   // single-output kernel, e.g. GT4GPU_STREAM_SHAPE=512x11
  if (env) {
    int nc = 0, vt = 0;
    if (sscanf (env, "%dx%d", &nc, &vt) != 2 || !stream_shape_supported (nc, vt))
      return fail (GT4GPU_ERR_ARG, "GT4GPU_STREAM_SHAPE=%s is not supported", env);
    g_ctx.stream_consumers = nc;
    g_ctx.stream_items = vt;
  }
  env = getenv ("GT4GPU_USE_STREAM_KERNEL");
  if (env) g_ctx.use_stream = atoi (env) != 0;
  env = getenv ("GT4GPU_USE_KWAY");
  if (env) g_ctx.use_kway = atoi (env) < 0 ? 0 : atoi (env) > 2 ? 2 : atoi (env);
  env = getenv ("GT4GPU_TILE");   // multi-output kernel, e.g. GT4GPU_TILE=256x11
  if (env) {
    int nt = 0, vt = 0;
    if (sscanf (env, "%dx%d", &nt, &vt) == 2 && tile_shape_supported (nt, vt)) g_ctx.shape = TileShape{nt, vt};
    else return fail (GT4GPU_ERR_ARG, "GT4GPU_TILE=%s is not a supported tile shape", env);
  }
  g_ctx.ready = true;
  return 0;
}

void gt4gpu_shutdown (void)
{
  std::lock_guard<std::mutex> lock (g_init_mutex);
  if (!g_ctx.ready) return;
  cudaStreamSynchronize (g_ctx.stream);
  g_pinned.release_all ();
  if (g_ctx.own_stream) cudaStreamDestroy (g_ctx.own_stream);
  if (g_ctx.in_stream) cudaStreamDestroy (g_ctx.in_stream);
  if (g_ctx.out_stream) cudaStreamDestroy (g_ctx.out_stream);
  g_ctx = Context ();
}

int gt4gpu_set_stream (void *cuda_stream)
{
  int rc = ensure_ready ();
  if (rc) return rc;
  g_ctx.stream = cuda_stream ? static_cast<cudaStream_t> (cuda_stream) : g_ctx.own_stream;
  return 0;
}

const char *gt4gpu_last_error (void) { return tl_error; }

int gt4gpu_set_tile (int threads, int items_per_thread)
{
  if (!tile_shape_supported (threads, items_per_thread))
    return fail (GT4GPU_ERR_ARG, "unsupported tile shape %dx%d", threads, items_per_thread);
  g_ctx.shape = TileShape{threads, items_per_thread};
  return 0;
}

int gt4gpu_set_option (const char *name, int value)
{
  if (!name) return fail (GT4GPU_ERR_ARG, "null option name");
  if (!strcmp (name, "stream_shape")) {        // consumers * 100 + items, e.g. 51209
    if (!stream_shape_supported (value / 100, value % 100)) return fail (GT4GPU_ERR_ARG, "stream shape %dx%d is not supported", value / 100, value % 100);
    g_ctx.stream_consumers = value / 100;
    g_ctx.stream_items = value % 100;
    return 0;
  }
  if (!strcmp (name, "stream_items")) {
    if (!stream_shape_supported (g_ctx.stream_consumers, value)) return fail (GT4GPU_ERR_ARG, "stream shape %dx%d is not supported", g_ctx.stream_consumers, value);
    g_ctx.stream_items = value;
    return 0;
  }
  if (!strcmp (name, "stream_consumers")) {
    if (!stream_shape_supported (value, g_ctx.stream_items)) return fail (GT4GPU_ERR_ARG, "stream shape %dx%d is not supported", value, g_ctx.stream_items);
    g_ctx.stream_consumers = value;
    return 0;
  }
  if (!strcmp (name, "use_kway")) {            // 0: N-list calls run as a tree / chain of two-list merges; 1: unions take the single pass; 2: intersections too
    g_ctx.use_kway = value < 0 ? 0 : value > 2 ? 2 : value;
    return 0;
  }
  if (!strcmp (name, "use_fused")) {
    g_ctx.use_fused = value != 0;
    return 0;
  }
  if (!strcmp (name, "stream_side")) {       // 0: intersections / differences take the plain stream kernel too
    g_stream_side = value < 0 ? 0 : value > 3 ? 3 : value;
    return 0;
  }
  if (!strcmp (name, "use_stream_kernel")) {
    g_ctx.use_stream = value != 0;
    return 0;
  }
  return fail (GT4GPU_ERR_ARG, "unknown option %s", name);
}

int gt4gpu_last_timing (float *ms_partition, float *ms_merge, uint32_t *n_launches)
{
  if (ms_partition) *ms_partition = tl_ms_partition;
  if (ms_merge) *ms_merge = tl_ms_merge;
  if (n_launches) *n_launches = tl_launches;
  return 0;
}

// ------------------------------------------------------------------ containers

int gt4gpu_list_read_header (const char *path, int stream_mode, gt4gpu_header *out)
{
  if (!path || !out) return fail (GT4GPU_ERR_ARG, "null argument");
  Mapping m;
  int rc = map_file (path, m);
  if (rc) return rc;
  if (is_index_file (m.data, m.size)) return parse_index_header (m.data, m.size, path, out, nullptr);
  return parse_header (m.data, m.size, stream_mode, path, out);
}

int gt4gpu_list_open_range (const char *path, int stream_mode, uint64_t first, uint64_t count, gt4gpu_list **out)
{
  if (!path || !out) return fail (GT4GPU_ERR_ARG, "null argument");
  Mapping m;
  int rc = map_file (path, m);
  if (rc) return rc;
  gt4gpu_header h;
  const bool index = is_index_file (m.data, m.size);     // glistcompare sniffs the 4-byte tag the same way (src/glistcompare.c:256-274)
  uint64_t num_locations = 0;
  rc = index ? parse_index_header (m.data, m.size, path, &h, &num_locations) : parse_header (m.data, m.size, stream_mode, path, &h);
  if (rc) return rc;
  if (first > h.n_words) first = h.n_words;
  if (count > h.n_words - first) count = h.n_words - first;
  rc = ensure_ready ();
  if (rc) return rc;
  gt4gpu_list *l = nullptr;
  rc = new_list (count, h.word_length, &l);
  if (rc) return rc;
  l->sum_counts = h.total_count;
  if (index) rc = upload_index (m.data + h.list_start, h.n_words, num_locations, first, count, l->words, l->counts);
  else rc = upload_aos (m.data + h.list_start + first * 12, count, l->words, l->counts);
  if (rc) { gt4gpu_list_close (l); return rc; }
  *out = l;
  return 0;
}

int gt4gpu_list_open (const char *path, int stream_mode, gt4gpu_list **out)
{
  return gt4gpu_list_open_range (path, stream_mode, 0, UINT64_MAX, out);
}

int gt4gpu_list_from_host_aos (const void *records, uint64_t n_words, uint32_t word_length, gt4gpu_list **out)
{
  if (!out || (n_words && !records)) return fail (GT4GPU_ERR_ARG, "null argument");
  int rc = ensure_ready ();
  if (rc) return rc;
  gt4gpu_list *l = nullptr;
  rc = new_list (n_words, word_length, &l);
  if (rc) return rc;
  rc = upload_aos (records, n_words, l->words, l->counts);
  if (rc) { gt4gpu_list_close (l); return rc; }
  *out = l;
  return 0;
}

int gt4gpu_list_from_host_soa (const uint64_t *words, const uint32_t *counts, uint64_t n_words, uint32_t word_length, gt4gpu_list **out)
{
  if (!out || (n_words && (!words || !counts))) return fail (GT4GPU_ERR_ARG, "null argument");
  int rc = ensure_ready ();
  if (rc) return rc;
  gt4gpu_list *l = nullptr;
  rc = new_list (n_words, word_length, &l);
  if (rc) return rc;
  if (n_words) {
    CU (cudaMemcpyAsync (l->words, words, n_words * sizeof (uint64_t), cudaMemcpyHostToDevice, g_ctx.stream));
    CU (cudaMemcpyAsync (l->counts, counts, n_words * sizeof (uint32_t), cudaMemcpyHostToDevice, g_ctx.stream));
    CU (cudaStreamSynchronize (g_ctx.stream));
  }
  *out = l;
  return 0;
}

int gt4gpu_list_from_device (const uint64_t *d_words, const uint32_t *d_counts, uint64_t n_words, uint32_t word_length, gt4gpu_list **out)
{
  if (!out || (n_words && (!d_words || !d_counts))) return fail (GT4GPU_ERR_ARG, "null argument");
  gt4gpu_list *l = static_cast<gt4gpu_list *> (calloc (1, sizeof (gt4gpu_list)));
  if (!l) return fail (GT4GPU_ERR_ARG, "out of host memory");
  l->words = const_cast<uint64_t *> (d_words);
  l->counts = const_cast<uint32_t *> (d_counts);
  l->n_words = n_words;
  l->word_length = word_length;
  l->owned = 0;
  *out = l;
  return 0;
}

void gt4gpu_list_close (gt4gpu_list *list)
{
  if (!list) return;
  if (list->owned) {
    dev_free (list->words);
    dev_free (list->counts);
  }
  free (list);
}

int gt4gpu_list_to_host_soa (const gt4gpu_list *list, uint64_t *words, uint32_t *counts)
{
  if (!list || (list->n_words && (!words || !counts))) return fail (GT4GPU_ERR_ARG, "null argument");
  int rc = ensure_ready ();
  if (rc) return rc;
  if (!list->n_words) return 0;
  CU (cudaMemcpyAsync (words, list->words, list->n_words * sizeof (uint64_t), cudaMemcpyDeviceToHost, g_ctx.stream));
  CU (cudaMemcpyAsync (counts, list->counts, list->n_words * sizeof (uint32_t), cudaMemcpyDeviceToHost, g_ctx.stream));
  CU (cudaStreamSynchronize (g_ctx.stream));
  return 0;
}

uint64_t gt4gpu_list_n_words (const gt4gpu_list *l) { return l ? l->n_words : 0; }
uint32_t gt4gpu_list_word_length (const gt4gpu_list *l) { return l ? l->word_length : 0; }
uint64_t gt4gpu_list_sum_counts (const gt4gpu_list *l) { return l ? l->sum_counts : 0; }
const uint64_t *gt4gpu_list_device_words (const gt4gpu_list *l) { return l ? l->words : nullptr; }
const uint32_t *gt4gpu_list_device_counts (const gt4gpu_list *l) { return l ? l->counts : nullptr; }

// ------------------------------------------------------------------ merges

int gt4gpu_compare2 (const gt4gpu_list *a, const gt4gpu_list *b, uint32_t ops, int rule, uint32_t cutoff,
                     uint32_t count_override, int subtract, int countonly, gt4gpu_result out[4])
{
  if (!a || !b || !out) return fail (GT4GPU_ERR_ARG, "null argument");
  if (!ops || (ops & ~15u)) return fail (GT4GPU_ERR_ARG, "ops must be a non-empty OR of GT4GPU_OP_*");
  if (rule < GT4GPU_RULE_DEFAULT || rule > GT4GPU_RULE_NUMBER) return fail (GT4GPU_ERR_ARG, "unknown rule %d", rule);
  int rc = ensure_ready ();
  if (rc) return rc;
  reset_timing ();

  SetOpParams p;
  memset (&p, 0, sizeof (p));
  p.ops = ops;
  p.cutoff = cutoff;
  p.count_override = count_override;
  p.subtract = subtract ? 1 : 0;
  p.sem = SEM_PAIR;
  for (int s = 0; s < 4; s++) p.rule[s] = resolve_rule (rule, s);

  MergeOut mo[4];
  for (int s = 0; s < 4; s++) {
    if (!((ops >> s) & 1u) || countonly) continue;
    if (out[s].flags & GT4GPU_RESULT_CALLER_BUFFERS) {
      mo[s].caller = true;
      mo[s].words = out[s].words;
      mo[s].counts = out[s].counts;
      mo[s].capacity = out[s].capacity;
    }
  }
  const DevList da{a->words, a->counts, a->n_words}, db{b->words, b->counts, b->n_words};
  rc = merge2_device (da, db, p, ops, countonly != 0, mo);
  if (rc) {
    for (int s = 0; s < 4; s++) free_out (mo[s]);
    return rc;
  }
  // output header word length = first list's (src/glistcompare.c:814)
  for (int s = 0; s < 4; s++) if ((ops >> s) & 1u) fill_result (&out[s], mo[s], a->word_length, countonly != 0);
  return 0;
}

int gt4gpu_union_multi (const gt4gpu_list *const *lists, unsigned n_lists, uint32_t cutoff, int rule,
                        uint32_t count_override, int countonly, gt4gpu_result *out)
{
  if (!lists || !out || n_lists == 0) return fail (GT4GPU_ERR_ARG, "null argument");
  // allowed rules, src/glistcompare.c:518-523
  if (rule == GT4GPU_RULE_DEFAULT) rule = GT4GPU_RULE_ADD;
  else if (rule != GT4GPU_RULE_ADD && rule != GT4GPU_RULE_MAX && rule != GT4GPU_RULE_NUMBER) {
    fprintf (stderr, "union_multi: Invalid rule %u (only ADD, MAX and NUMBER allowed)\n", rule);
    return fail (GT4GPU_ERR_ARG, "union_multi: invalid rule %d", rule);
  }
  int rc = ensure_ready ();
  if (rc) return rc;
  reset_timing ();
  // only non-empty lists take part (:526-533); the header word length is the first non-empty
  // list's, or the last list's when all are empty (:535)
  std::vector<DevList> level;
  uint32_t k = lists[n_lists - 1]->word_length;
  bool have_k = false;
  for (unsigned j = 0; j < n_lists; j++) {
    if (!lists[j]) return fail (GT4GPU_ERR_ARG, "null list");
    if (lists[j]->n_words) {
      if (!have_k) { k = lists[j]->word_length; have_k = true; }
      level.push_back (DevList{lists[j]->words, lists[j]->counts, lists[j]->n_words});
    }
  }
  MergeOut root;
  if (!countonly && (out->flags & GT4GPU_RESULT_CALLER_BUFFERS)) {
    root.caller = true;
    root.words = out->words;
    root.counts = out->counts;
    root.capacity = out->capacity;
  }
  bool fallback = true;
  if (g_ctx.use_kway && level.size () >= 3) {
    rc = union_kway (level, rule, cutoff, count_override, countonly != 0, &root, &fallback);
    if (rc) return rc;
    if (fallback) reset_timing ();
  }
  if (fallback) rc = union_tree (level, rule, cutoff, count_override, countonly != 0, &root);
  if (rc) return rc;
  fill_result (out, root, k, countonly != 0);
  return 0;
}

int gt4gpu_intersect_multi (const gt4gpu_list *const *lists, unsigned n_lists, uint32_t cutoff, int rule,
                            uint32_t count_override, int countonly, gt4gpu_result *out)
{
  if (!lists || !out || n_lists == 0) return fail (GT4GPU_ERR_ARG, "null argument");
  // allowed rules, src/glistcompare.c:622-627
  if (rule == GT4GPU_RULE_DEFAULT) rule = GT4GPU_RULE_MIN;
  else if (rule != GT4GPU_RULE_ADD && rule != GT4GPU_RULE_MIN && rule != GT4GPU_RULE_MAX && rule != GT4GPU_RULE_NUMBER) {
    fprintf (stderr, "intersect_multi: Invalid rule %u (only ADD, MIN, MAX and NUMBER allowed)\n", rule);
    return fail (GT4GPU_ERR_ARG, "intersect_multi: invalid rule %d", rule);
  }
  int rc = ensure_ready ();
  if (rc) return rc;
  reset_timing ();
  for (unsigned j = 0; j < n_lists; j++) if (!lists[j]) return fail (GT4GPU_ERR_ARG, "null list");
  const uint32_t k = lists[0]->word_length;   // :639
  bool any_empty = false;
  for (unsigned j = 0; j < n_lists; j++) any_empty |= lists[j]->n_words == 0;

  MergeOut cur;       // running fold; starts as list 0 itself
  const bool caller = !countonly && (out->flags & GT4GPU_RESULT_CALLER_BUFFERS);
  if (any_empty) {
    // any empty list ends the reference loop before the first comparison (:631-636)
    MergeOut empty;
    if (caller) { empty.caller = true; empty.words = out->words; empty.counts = out->counts; empty.capacity = out->capacity; }
    fill_result (out, empty, k, countonly != 0);
    return 0;
  }
  if (n_lists == 1) {
    // one list: every word is "in all lists" and the fold of a single count is that count
    // (number: the override); expressed as a final union node against an empty list
    const int r1 = (rule == GT4GPU_RULE_NUMBER) ? GT4GPU_RULE_NUMBER : GT4GPU_RULE_ADD;
    MergeOut mo[4];
    if (caller) { mo[0].caller = true; mo[0].words = out->words; mo[0].counts = out->counts; mo[0].capacity = out->capacity; }
    SetOpParams p = nlist_params (SEM_NUNION_FINAL, r1, cutoff, count_override);
    rc = merge2_device (DevList{lists[0]->words, lists[0]->counts, lists[0]->n_words}, DevList{nullptr, nullptr, 0}, p, OP_UNION, countonly != 0, mo);
    if (rc) { free_out (mo[0]); return rc; }
    fill_result (out, mo[0], k, countonly != 0);
    return 0;
  }
  // The chain below already reads every list once and its intermediates only shrink, so its traffic is the algorithmic
  // one; measured on 8 lists sharing a third of a universe it beats the single pass (2.8 vs 4.7 ms, profiles/
  // r02_config5_one_gpu.jsonl).  The single pass is kept selectable (use_kway = 2) and tested.
  if (g_ctx.use_kway >= 2 && n_lists >= 3) {
    std::vector<DevList> all;
    for (unsigned j = 0; j < n_lists; j++) all.push_back (DevList{lists[j]->words, lists[j]->counts, lists[j]->n_words});
    MergeOut root;
    if (caller) { root.caller = true; root.words = out->words; root.counts = out->counts; root.capacity = out->capacity; }
    bool fallback = false;
    rc = intersect_kway (all, rule, cutoff, count_override, countonly != 0, &root, &fallback);
    if (rc) return rc;
    if (!fallback) {
      fill_result (out, root, k, countonly != 0);
      return 0;
    }
    reset_timing ();
  }
  // left chain ((L0 ^ L1) ^ L2) ...: exactly the reference's in-order fold (:668-677), including
  // the "!freq ||" guard of rule min; only the last link applies the cut-off (:682)
  DevList acc{lists[0]->words, lists[0]->counts, lists[0]->n_words};
  bool cur_owned = false;
  for (unsigned j = 1; j < n_lists; j++) {
    const bool last = (j + 1 == n_lists);
    MergeOut mo[4];
    if (last && caller) { mo[0].caller = true; mo[0].words = out->words; mo[0].counts = out->counts; mo[0].capacity = out->capacity; }
    SetOpParams p = nlist_params (last ? SEM_NISECT_FINAL : SEM_NISECT_PARTIAL, rule, cutoff, count_override);
    rc = merge2_device (acc, DevList{lists[j]->words, lists[j]->counts, lists[j]->n_words}, p, OP_UNION, last && countonly, mo);
    if (cur_owned) free_out (cur);
    if (rc) { free_out (mo[0]); return rc; }
    cur = mo[0];
    cur_owned = true;
    acc = DevList{cur.words, cur.counts, cur.n};
  }
  fill_result (out, cur, k, countonly != 0);
  return 0;
}

int gt4gpu_write_union (const gt4gpu_list *const *lists, unsigned n_lists, uint32_t cutoff, int ofile, gt4gpu_header *header)
{
  // preconditions of src/set-operations.c:49-50
  if (!lists || !header || n_lists == 0 || n_lists > 4096) return fail (GT4GPU_ERR_ARG, "gt4gpu_write_union: bad arguments");
  gt4gpu_result res;
  memset (&res, 0, sizeof (res));
  int rc = gt4gpu_union_multi (lists, n_lists, cutoff, GT4GPU_RULE_ADD, 0, ofile == 0, &res);
  if (rc) return rc;
  gt4gpu_header_init (header, res.word_length);
  header->n_words = res.n_words;
  header->total_count = res.total_count;
  if (ofile) {
    // header at the current position, records after it, final header at offset 0 (:75,:117-119)
    rc = write_all (ofile, header, sizeof (*header), -1);
    if (!rc) rc = download_aos (res.words, res.counts, res.n_words, nullptr,
                                [&] (const void *p, size_t bytes, uint64_t) { return write_all (ofile, p, bytes, -1); }, true);
    if (!rc) rc = write_all (ofile, header, sizeof (*header), 0);
  }
  gt4gpu_result_free (&res);
  return rc;
}

int gt4gpu_union_matrix (const gt4gpu_list *const *lists, unsigned n_lists, int is_union,
                         uint64_t *words, uint32_t *counts, uint64_t max_rows, uint64_t *n_rows)
{
  if (!lists || !n_rows || n_lists == 0 || n_lists > 4096) return fail (GT4GPU_ERR_ARG, "gt4gpu_union_matrix: bad arguments");
  for (unsigned j = 0; j < n_lists; j++)
    if (!lists[j] || lists[j]->n_words == 0) return fail (GT4GPU_ERR_ARG, "gt4gpu_union_matrix: list %u is empty (undefined in the reference)", j);
  int rc = ensure_ready ();
  if (rc) return rc;
  // row keys: list 0 (is_union) or the N-list union with nothing filtered
  gt4gpu_result rows;
  memset (&rows, 0, sizeof (rows));
  const uint64_t *d_rows;
  uint64_t n;
  if (is_union) {
    d_rows = lists[0]->words;
    n = lists[0]->n_words;
  } else {
    rc = gt4gpu_union_multi (lists, n_lists, 0, GT4GPU_RULE_MAX, 0, 0, &rows);   // max never wraps: every key survives cutoff 0
    if (rc) return rc;
    d_rows = rows.words;
    n = rows.n_words;
  }
  uint32_t *d_matrix = nullptr;
  rc = dev_alloc ((void **) &d_matrix, n * n_lists * sizeof (uint32_t));
  if (rc) { gt4gpu_result_free (&rows); return rc; }
  cudaError_t e = cudaMemsetAsync (d_matrix, 0, n * n_lists * sizeof (uint32_t), g_ctx.stream);
  for (unsigned j = 0; j < n_lists && e == cudaSuccess; j++)
    e = launch_scatter_counts (d_rows, n, lists[j]->words, lists[j]->counts, lists[j]->n_words, j, n_lists, d_matrix, g_ctx.stream);
  std::vector<uint64_t> h_rows (n);
  std::vector<uint32_t> h_mat (n * n_lists);
  if (e == cudaSuccess) e = cudaMemcpyAsync (h_rows.data (), d_rows, n * sizeof (uint64_t), cudaMemcpyDeviceToHost, g_ctx.stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync (h_mat.data (), d_matrix, n * n_lists * sizeof (uint32_t), cudaMemcpyDeviceToHost, g_ctx.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize (g_ctx.stream);
  dev_free (d_matrix);
  gt4gpu_result_free (&rows);
  if (e != cudaSuccess) return fail (GT4GPU_ERR_CUDA, "gt4gpu_union_matrix: %s", cudaGetErrorString (e));

  // gt4_union revisits the stale last word of a list that runs out while others are still live
  // and calls back once more with all-zero counts (src/set-operations.c:159-172); reproduce
  // those rows: one after each word that ends some list, unless it is the overall last word.
  std::vector<uint64_t> last_words;
  if (!is_union) {
    last_words.resize (n_lists);
    for (unsigned j = 0; j < n_lists; j++)
      CU (cudaMemcpyAsync (&last_words[j], lists[j]->words + lists[j]->n_words - 1, sizeof (uint64_t), cudaMemcpyDeviceToHost, g_ctx.stream));
    CU (cudaStreamSynchronize (g_ctx.stream));
    std::sort (last_words.begin (), last_words.end ());
    last_words.erase (std::unique (last_words.begin (), last_words.end ()), last_words.end ());
  }
  uint64_t r = 0;
  size_t lw = 0;
  auto put = [&] (uint64_t w, const uint32_t *c) {
    if (r < max_rows && words && counts) {
      words[r] = w;
      for (unsigned j = 0; j < n_lists; j++) counts[r * n_lists + j] = c ? c[j] : 0u;
    }
    r += 1;
  };
  for (uint64_t i = 0; i < n; i++) {
    put (h_rows[i], &h_mat[i * n_lists]);
    if (!is_union) {
      while (lw < last_words.size () && last_words[lw] < h_rows[i]) lw++;
      if (lw < last_words.size () && last_words[lw] == h_rows[i] && i + 1 < n) put (h_rows[i], nullptr);
    }
  }
  *n_rows = r;
  return 0;
}

// ------------------------------------------------------------------ results

// ------------------------------------------------------------------ lookups

// Device-side body of gt4gpu_lookup.  Large batches are sorted first (key-value radix sort of the canonical words with
// their positions): neighbouring threads then share their search paths and the probes hit cache instead of DRAM.
static int lookup_on_device (const gt4gpu_list *list, const uint64_t *d_queries, uint64_t n, int canonize, uint32_t *d_counts,
                             uint64_t *d_canonical)
{
  cudaStream_t st = g_ctx.stream;
  const char *env = getenv ("GT4GPU_LOOKUP_SORT_MIN");
  const uint64_t sort_min = env ? strtoull (env, nullptr, 10) : (1ull << 20);
  if (n < sort_min || n > 0xffffffffull || list->n_words < 2) {
    CU (launch_lookup (list->words, list->counts, list->n_words, list->word_length, canonize, d_queries, n, d_canonical, d_counts, st));
    return 0;
  }
  struct Scratch {
    void *p[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    ~Scratch () { for (void *q : p) dev_free (q); }
  } tmp;
  int rc;
  if ((rc = dev_alloc (&tmp.p[0], n * sizeof (uint64_t)))) return rc;
  if ((rc = dev_alloc (&tmp.p[1], n * sizeof (uint64_t)))) return rc;
  if ((rc = dev_alloc (&tmp.p[2], n * sizeof (uint32_t)))) return rc;
  if ((rc = dev_alloc (&tmp.p[3], n * sizeof (uint32_t)))) return rc;
  if ((rc = dev_alloc (&tmp.p[4], sort_scratch_bytes (n)))) return rc;
  uint64_t *keys = (uint64_t *) tmp.p[0], *alt = (uint64_t *) tmp.p[1];
  CU (launch_canonize (d_queries, n, list->word_length, canonize, keys, st));
  if (d_canonical) CU (cudaMemcpyAsync (d_canonical, keys, n * sizeof (uint64_t), cudaMemcpyDeviceToDevice, st));
  // a batch that arrives in order (a list's own words, a sorted query file) needs no sorting
  uint32_t unsorted = 1;
  CU (launch_check_sorted (keys, n, (uint32_t *) tmp.p[2], st));
  CU (cudaMemcpyAsync (&unsorted, tmp.p[2], sizeof (unsorted), cudaMemcpyDeviceToHost, st));
  CU (cudaStreamSynchronize (st));
  if (!unsorted) {
    CU (launch_lookup (list->words, list->counts, list->n_words, list->word_length, 0, keys, n, nullptr, d_counts, st));
    return 0;
  }
  const int n_pass = (int) ((2 * list->word_length + 7) / 8);
  uint64_t *sorted = nullptr;
  uint32_t *perm = nullptr;
  CU (launch_radix_sort_pairs (keys, alt, (uint32_t *) tmp.p[2], (uint32_t *) tmp.p[3], n, n_pass, (unsigned char *) tmp.p[4],
                               g_ctx.sm_count, &sorted, &perm, st));
  CU (launch_lookup_sorted (list->words, list->counts, list->n_words, sorted, perm, n, d_counts, st));
  return 0;
}

int gt4gpu_lookup (const gt4gpu_list *list, const uint64_t *queries, uint64_t n_queries, int on_device, int canonize,
                   uint32_t *counts_out, uint64_t *canonical_out)
{
  if (!list || !counts_out || (!queries && n_queries)) return fail (GT4GPU_ERR_ARG, "null argument");
  int rc = ensure_ready ();
  if (rc) return rc;
  if (n_queries == 0) return 0;
  cudaStream_t st = g_ctx.stream;
  if (on_device) {
    if ((rc = lookup_on_device (list, queries, n_queries, canonize, counts_out, canonical_out))) return rc;
    CU (cudaStreamSynchronize (st));
    return 0;
  }
  struct Scratch {
    void *p[3] = {nullptr, nullptr, nullptr};
    ~Scratch () { for (void *q : p) dev_free (q); }
  } tmp;
  if ((rc = dev_alloc (&tmp.p[0], n_queries * sizeof (uint64_t)))) return rc;
  if ((rc = dev_alloc (&tmp.p[1], n_queries * sizeof (uint32_t)))) return rc;
  if (canonical_out && (rc = dev_alloc (&tmp.p[2], n_queries * sizeof (uint64_t)))) return rc;
  CU (cudaMemcpyAsync (tmp.p[0], queries, n_queries * sizeof (uint64_t), cudaMemcpyHostToDevice, st));
  if ((rc = lookup_on_device (list, (const uint64_t *) tmp.p[0], n_queries, canonize, (uint32_t *) tmp.p[1], (uint64_t *) tmp.p[2]))) return rc;
  CU (cudaMemcpyAsync (counts_out, tmp.p[1], n_queries * sizeof (uint32_t), cudaMemcpyDeviceToHost, st));
  if (canonical_out) CU (cudaMemcpyAsync (canonical_out, tmp.p[2], n_queries * sizeof (uint64_t), cudaMemcpyDeviceToHost, st));
  CU (cudaStreamSynchronize (st));
  return 0;
}

// ------------------------------------------------------------------ list building

// Host-side reader of FastA / FastQ images (fasta_reader_read_nwords, src/fasta.c:88-290, with canonize = 1 as
// glistmaker sets it, src/listmaker-queue.c:196).  Table driven: every byte is a nucleotide (0..3), a word break
// (printable, not ACGTU) or skipped (control characters such as line ends, :263-269).
namespace {

enum : uint8_t { CH_BREAK = 4, CH_SKIP = 5 };

struct CharClass {
  uint8_t v[256];
  CharClass ()
  {
    for (int c = 0; c < 256; c++) v[c] = c < ' ' ? CH_SKIP : CH_BREAK;
    v['A'] = v['a'] = 0;
    v['C'] = v['c'] = 1;
    v['G'] = v['g'] = 2;
    v['T'] = v['t'] = v['U'] = v['u'] = 3;
  }
};
const CharClass g_chars;

struct WordWindow {       // the sliding forward / reverse-complement pair of one reader
  uint64_t fw = 0, rc = 0, mask;
  unsigned have = 0, k, top;
  explicit WordWindow (unsigned k_) : mask (k_ >= 32 ? ~0ull : (1ull << (2 * k_)) - 1), k (k_), top (2 * (k_ - 1)) {}
  void reset () { fw = rc = 0; have = 0; }
  bool push (unsigned nucl, uint64_t *word)
  {
    fw = ((fw << 2) | nucl) & mask;
    rc = (rc >> 2) | ((uint64_t) (3u - nucl) << top);
    if (have < k) have++;
    if (have < k) return false;
    *word = fw < rc ? fw : rc;
    return true;
  }
};

}  // namespace

int gt4gpu_sequence_words (const void *text, uint64_t n_bytes, uint32_t word_length, uint64_t *words, uint64_t capacity,
                           uint64_t *n_words)
{
  if ((!text && n_bytes) || !n_words) return fail (GT4GPU_ERR_ARG, "null argument");
  if (word_length < 1 || word_length > 32) return fail (GT4GPU_ERR_ARG, "word length %u not in 1..32", word_length);
  const unsigned char *p = static_cast<const unsigned char *> (text), *const end = p + n_bytes;
  *n_words = 0;
  if (p == end || *p == 0) return 0;
  const bool fastq = (*p == '@');
  if (!fastq && *p != '>') return fail (GT4GPU_ERR_FORMAT, "invalid start tag '%c'", *p);    // :136-139
  WordWindow win (word_length);
  uint64_t n = 0;
  auto emit_run = [&] (const unsigned char *q, const unsigned char *stop, unsigned char terminator) -> const unsigned char * {
    // nucleotides up to `terminator` (or a zero byte / the end of the image)
    for (; q < stop && *q != terminator && *q != 0; q++) {
      const uint8_t cls = g_chars.v[*q];
      if (cls < CH_BREAK) {
        uint64_t w;
        if (win.push (cls, &w)) {
          if (words) {
            if (n >= capacity) return nullptr;
            words[n] = w;
          }
          n++;
        }
      } else if (cls == CH_BREAK) {
        win.reset ();
      }
    }
    return q;
  };
  int rc = 0;
  while (p < end && *p != 0) {
    // p is on a record tag ('>' or '@'): the name runs to the end of the line
    const unsigned char *nl = static_cast<const unsigned char *> (memchr (p, '\n', (size_t) (end - p)));
    const unsigned char *zero = static_cast<const unsigned char *> (memchr (p, 0, (size_t) ((nl ? nl : end) - p)));
    if (zero || !nl) break;                    // the image ends inside a name
    p = nl + 1;
    win.reset ();
    if (!fastq) {
      p = emit_run (p, end, '>');              // a '>' anywhere in the sequence starts the next name (:177-190)
      if (!p) { rc = GT4GPU_ERR_CAPACITY; break; }
    } else {
      p = emit_run (p, end, '\n');             // one line of sequence (:191)
      if (!p) { rc = GT4GPU_ERR_CAPACITY; break; }
      if (p >= end || *p == 0) break;          // the image ends inside the sequence
      p++;
      if (p >= end || *p != '+') { rc = GT4GPU_ERR_FORMAT; break; }     // :203-206
      const unsigned char *q = static_cast<const unsigned char *> (memchr (p, '\n', (size_t) (end - p)));
      if (!q || memchr (p, 0, (size_t) (q - p))) { rc = GT4GPU_ERR_FORMAT; break; }   // :210-214
      p = q + 1;                               // quality line
      q = static_cast<const unsigned char *> (memchr (p, '\n', (size_t) (end - p)));
      if (!q || memchr (p, 0, (size_t) (q - p))) { p = end; break; }     // EOF inside the quality: nothing more to read
      p = q + 1;
      if (p >= end || *p == 0) break;
      if (*p != '@') { rc = GT4GPU_ERR_FORMAT; break; }                 // :284-287
    }
  }
  *n_words = n;
  if (rc == GT4GPU_ERR_CAPACITY) return fail (rc, "word buffer too small");
  if (rc) return fail (rc, "malformed FastQ record");
  return 0;
}

int gt4gpu_fasta_words_device (const void *text, uint64_t n_bytes, uint32_t word_length, uint64_t **d_words, uint64_t *n_words)
{
  if ((!text && n_bytes) || !d_words || !n_words) return fail (GT4GPU_ERR_ARG, "null argument");
  if (word_length < 1 || word_length > 32) return fail (GT4GPU_ERR_ARG, "word length %u not in 1..32", word_length);
  *d_words = nullptr;
  *n_words = 0;
  const unsigned char *p = static_cast<const unsigned char *> (text);
  if (n_bytes == 0 || p[0] == 0) return 0;
  const bool fastq = p[0] == '@';
  if (!fastq && p[0] != '>') return fail (GT4GPU_ERR_FORMAT, "invalid start tag '%c'", p[0]);          // src/fasta.c:136-139
  if (const void *z = memchr (p, 0, (size_t) n_bytes)) n_bytes = (uint64_t) (static_cast<const unsigned char *> (z) - p);   // :107-118
  int rc = ensure_ready ();
  if (rc) return rc;
  cudaStream_t st = g_ctx.stream;
  struct Scratch {
    void *p[3] = {nullptr, nullptr, nullptr};
    ~Scratch () { for (void *q : p) dev_free (q); }
  } tmp;
  if ((rc = dev_alloc (&tmp.p[0], n_bytes))) return rc;
  if ((rc = dev_alloc (&tmp.p[1], n_bytes))) return rc;
  if ((rc = dev_alloc (&tmp.p[2], fasta_scratch_bytes (fasta_chunks (n_bytes))))) return rc;
  uint8_t *d_text = (uint8_t *) tmp.p[0], *d_codes = (uint8_t *) tmp.p[1];
  unsigned char *ws = (unsigned char *) tmp.p[2];
  CU (cudaMemcpyAsync (d_text, text, n_bytes, cudaMemcpyHostToDevice, st));
  const uint64_t *d_n = nullptr;
  const uint32_t *d_bad = nullptr;
  const uint64_t *d_lines = nullptr;
  uint32_t malformed = 0;
  uint64_t n_lines = 0;
  if (fastq) CU (launch_fastq_codes (d_text, n_bytes, ws, d_codes, &d_n, &d_bad, &d_lines, st));
  else CU (launch_fasta_codes (d_text, n_bytes, ws, d_codes, &d_n, st));
  uint64_t n_codes = 0;
  CU (cudaMemcpyAsync (&n_codes, d_n, sizeof (n_codes), cudaMemcpyDeviceToHost, st));
  if (fastq) {
    CU (cudaMemcpyAsync (&malformed, d_bad, sizeof (malformed), cudaMemcpyDeviceToHost, st));
    CU (cudaMemcpyAsync (&n_lines, d_lines, sizeof (n_lines), cudaMemcpyDeviceToHost, st));
  }
  CU (cudaStreamSynchronize (st));
  if (fastq && (n_lines & 3) == 2) malformed = 1;        // the image stops on or inside a '+' line (src/fasta.c:203-214)
  if (malformed) return fail (GT4GPU_ERR_FORMAT, "malformed FastQ record (gt4gpu_sequence_words reads the image up to it, like the reference)");
  if (n_codes > n_bytes) return fail (GT4GPU_ERR_CUDA, "sequence pass kept %llu codes of %llu bytes", (unsigned long long) n_codes, (unsigned long long) n_bytes);
  CU (launch_fasta_word_counts (d_codes, n_codes, word_length, ws, &d_n, st));
  uint64_t n = 0;
  CU (cudaMemcpyAsync (&n, d_n, sizeof (n), cudaMemcpyDeviceToHost, st));
  CU (cudaStreamSynchronize (st));
  if (n > n_codes) return fail (GT4GPU_ERR_CUDA, "word pass counted %llu words in %llu codes", (unsigned long long) n, (unsigned long long) n_codes);
  if (n == 0) return 0;
  uint64_t *words = nullptr;
  if ((rc = dev_alloc ((void **) &words, n * sizeof (uint64_t)))) return rc;
  cudaError_t e = launch_fasta_words (d_codes, n_codes, word_length, ws, words, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize (st);
  if (e != cudaSuccess) {
    dev_free (words);
    return fail (GT4GPU_ERR_CUDA, "word pass: %s", cudaGetErrorString (e));
  }
  *d_words = words;
  *n_words = n;
  return 0;
}

void gt4gpu_device_free (void *d_ptr)
{
  if (d_ptr && g_ctx.ready) dev_free (d_ptr);
}

int gt4gpu_count_words (const uint64_t *words, uint64_t n_words, int on_device, uint32_t word_length, gt4gpu_result *out)
{
  if (!out || (!words && n_words)) return fail (GT4GPU_ERR_ARG, "null argument");
  if (word_length < 1 || word_length > 32) return fail (GT4GPU_ERR_ARG, "word length %u not in 1..32", word_length);
  int rc = ensure_ready ();
  if (rc) return rc;
  reset_timing ();
  cudaStream_t st = g_ctx.stream;
  memset (out, 0, sizeof (*out));
  out->word_length = word_length;
  if (n_words == 0) return 0;

  // words are < 4^k: only the low 2k bits need sorting
  const int n_pass = (int) ((2 * word_length + 7) / 8);
  const uint64_t n = n_words;
  struct Scratch {          // every temporary goes back to the pool on every exit path
    void *p[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    ~Scratch () { for (void *q : p) dev_free (q); }
  } tmp;
  uint64_t *keys = nullptr, *alt = nullptr, *first = nullptr;
  unsigned char *ws_sort = nullptr, *ws_rle = nullptr;
  if ((rc = dev_alloc (&tmp.p[0], n * sizeof (uint64_t)))) return rc;
  if ((rc = dev_alloc (&tmp.p[1], n * sizeof (uint64_t)))) return rc;
  if ((rc = dev_alloc (&tmp.p[2], sort_scratch_bytes (n)))) return rc;
  keys = (uint64_t *) tmp.p[0]; alt = (uint64_t *) tmp.p[1]; ws_sort = (unsigned char *) tmp.p[2];
  const uint64_t *input = keys;
  if (on_device) input = words;          // sorted straight out of the caller's array (read only)
  else CU (cudaMemcpyAsync (keys, words, n * sizeof (uint64_t), cudaMemcpyHostToDevice, st));

  for (int i = 0; i < 3; i++) if (!tl_ev[i]) CU (cudaEventCreate (&tl_ev[i]));
  CU (cudaEventRecord (tl_ev[0], st));
  uint64_t *sorted = nullptr;
  CU (launch_radix_sort (input, keys, alt, n, n_pass, ws_sort, g_ctx.sm_count, &sorted, st));
  CU (cudaEventRecord (tl_ev[1], st));
  uint64_t *words_tmp = (sorted == keys) ? alt : keys;      // the buffer the sort no longer needs
  if ((rc = dev_alloc (&tmp.p[3], n * sizeof (uint64_t)))) return rc;
  if ((rc = dev_alloc (&tmp.p[4], rle_scratch_bytes (n)))) return rc;
  first = (uint64_t *) tmp.p[3]; ws_rle = (unsigned char *) tmp.p[4];
  unsigned long long *d_unique = nullptr;
  CU (launch_rle_heads (sorted, n, words_tmp, first, ws_rle, g_ctx.sm_count, &d_unique, st));
  unsigned long long n_unique = 0, or_all = 0;
  CU (cudaMemcpyAsync (&n_unique, d_unique, sizeof (n_unique), cudaMemcpyDeviceToHost, st));
  CU (cudaMemcpyAsync (&or_all, ws_sort + SORT_OR_OFFSET, sizeof (or_all), cudaMemcpyDeviceToHost, st));
  CU (cudaStreamSynchronize (st));
  // only the low 2k bits were sorted: a word >= 4^k would have left the table unsorted (and the list malformed)
  if (word_length < 32 && (or_all >> (2 * word_length)) != 0)
    return fail (GT4GPU_ERR_ARG, "gt4gpu_count_words: a word does not fit %u nucleotides (OR of all words = %llx)", word_length, or_all);
  if (n_unique == 0 || n_unique > n) return fail (GT4GPU_ERR_CUDA, "run-length pass returned %llu runs for %llu words", n_unique, (unsigned long long) n);

  uint64_t *res_words = nullptr;
  uint32_t *res_counts = nullptr;
  if ((rc = dev_alloc ((void **) &res_words, n_unique * sizeof (uint64_t)))) return rc;
  if ((rc = dev_alloc ((void **) &res_counts, n_unique * sizeof (uint32_t)))) { dev_free (res_words); return rc; }
  cudaError_t e = launch_rle_counts (words_tmp, first, n_unique, n, res_words, res_counts, st);
  if (e == cudaSuccess) e = cudaEventRecord (tl_ev[2], st);
  if (e == cudaSuccess) e = cudaStreamSynchronize (st);
  if (e != cudaSuccess) {
    dev_free (res_words); dev_free (res_counts);
    return fail (GT4GPU_ERR_CUDA, "run-length pass: %s", cudaGetErrorString (e));
  }
  cudaEventElapsedTime (&tl_ms_partition, tl_ev[0], tl_ev[1]);     // sort
  cudaEventElapsedTime (&tl_ms_merge, tl_ev[1], tl_ev[2]);         // run-length encoding
  tl_launches = 2 + n_pass + 2;
  out->n_words = n_unique;
  out->total_count = n;          // every word occurrence is counted once (merge_tables_to_file, :1139)
  out->words = res_words;
  out->counts = res_counts;
  out->capacity = n_unique;
  return 0;
}

int gt4gpu_result_to_host_soa (const gt4gpu_result *res, uint64_t *words, uint32_t *counts)
{
  if (!res) return fail (GT4GPU_ERR_ARG, "null argument");
  if (res->flags & GT4GPU_RESULT_COUNT_ONLY) return fail (GT4GPU_ERR_ARG, "count-only result holds no records");
  if (res->n_words == 0) return 0;
  if (!words || !counts) return fail (GT4GPU_ERR_ARG, "null argument");
  CU (cudaMemcpyAsync (words, res->words, res->n_words * sizeof (uint64_t), cudaMemcpyDeviceToHost, g_ctx.stream));
  CU (cudaMemcpyAsync (counts, res->counts, res->n_words * sizeof (uint32_t), cudaMemcpyDeviceToHost, g_ctx.stream));
  CU (cudaStreamSynchronize (g_ctx.stream));
  return 0;
}

int gt4gpu_result_to_host_aos (const gt4gpu_result *res, void *records)
{
  if (!res) return fail (GT4GPU_ERR_ARG, "null argument");
  if (res->flags & GT4GPU_RESULT_COUNT_ONLY) return fail (GT4GPU_ERR_ARG, "count-only result holds no records");
  if (res->n_words == 0) return 0;
  if (!records) return fail (GT4GPU_ERR_ARG, "null argument");
  return download_aos (res->words, res->counts, res->n_words, records, [] (const void *, size_t, uint64_t) { return 0; }, false);
}

int gt4gpu_write_records_at (const gt4gpu_result *res, int fd, uint64_t first_record)
{
  if (!res || fd < 0) return fail (GT4GPU_ERR_ARG, "bad argument");
  if (res->flags & GT4GPU_RESULT_COUNT_ONLY) return fail (GT4GPU_ERR_ARG, "count-only result holds no records");
  const int64_t base = (int64_t) (sizeof (gt4gpu_header) + 12ull * first_record);
  return download_aos (res->words, res->counts, res->n_words, nullptr,
                       [&] (const void *p, size_t bytes, uint64_t first) { return write_all (fd, p, bytes, base + (int64_t) (first * 12)); }, true);
}

int gt4gpu_write_list (const gt4gpu_result *res, int fd)
{
  if (!res || fd < 0) return fail (GT4GPU_ERR_ARG, "bad argument");
  gt4gpu_header h;
  gt4gpu_header_init (&h, res->word_length);
  h.n_words = res->n_words;
  h.total_count = res->total_count;
  int rc = write_all (fd, &h, sizeof (h), 0);
  if (rc) return rc;
  return gt4gpu_write_records_at (res, fd, 0);
}

void gt4gpu_result_free (gt4gpu_result *res)
{
  if (!res) return;
  if (!(res->flags & GT4GPU_RESULT_CALLER_BUFFERS)) {
    dev_free (res->words);
    dev_free (res->counts);
  }
  res->words = nullptr;
  res->counts = nullptr;
  res->capacity = 0;
}

// ------------------------------------------------------------------ host-to-host

// Small inputs: upload everything, merge once, download.
static int compare2_host_simple (const void *records_a, uint64_t n_a, const void *records_b, uint64_t n_b,
                                 uint32_t word_length, uint32_t ops, int rule, uint32_t cutoff,
                                 uint32_t count_override, int subtract, int countonly,
                                 void *const out_records[4], const uint64_t out_capacity[4],
                                 uint64_t n_out[4], uint64_t total_out[4])
{
  gt4gpu_list *a = nullptr, *b = nullptr;
  int rc = gt4gpu_list_from_host_aos (records_a, n_a, word_length, &a);
  if (!rc) rc = gt4gpu_list_from_host_aos (records_b, n_b, word_length, &b);
  gt4gpu_result res[4];
  memset (res, 0, sizeof (res));
  if (!rc) rc = gt4gpu_compare2 (a, b, ops, rule, cutoff, count_override, subtract, countonly, res);
  float ms_p = tl_ms_partition, ms_m = tl_ms_merge;
  uint32_t nl = tl_launches;
  for (int s = 0; s < 4 && !rc; s++) {
    if (!((ops >> s) & 1u)) continue;
    n_out[s] = res[s].n_words;
    total_out[s] = res[s].total_count;
    if (countonly) continue;
    if (!out_records || !out_capacity || (res[s].n_words && !out_records[s])) rc = fail (GT4GPU_ERR_ARG, "missing output buffer for stream %d", s);
    else if (res[s].n_words > out_capacity[s]) rc = fail (GT4GPU_ERR_CAPACITY, "stream %d needs %llu records", s, (unsigned long long) res[s].n_words);
    else rc = gt4gpu_result_to_host_aos (&res[s], out_records[s]);
  }
  for (int s = 0; s < 4; s++) gt4gpu_result_free (&res[s]);
  gt4gpu_list_close (a);
  gt4gpu_list_close (b);
  tl_ms_partition = ms_p; tl_ms_merge = ms_m; tl_launches = nl;
  return rc;
}

// Large inputs: the key space is cut into parts (the same exact splitters the multi-GPU path uses, so equal keys
// share a part and the concatenation of the per-part results IS the result); part p+1 is copied in and
// de-interleaved on its own stream while part p is merged and part p-1 is interleaved and copied out on a third
// stream.  PCIe is full duplex, so the whole call costs about max (H2D, D2H) instead of their sum.
static int compare2_host_pipelined (const unsigned char *rec_a, uint64_t n_a, const unsigned char *rec_b, uint64_t n_b,
                                    uint32_t word_length, uint32_t ops, int rule, uint32_t cutoff,
                                    uint32_t count_override, int subtract, int countonly,
                                    void *const out_records[4], const uint64_t out_capacity[4],
                                    uint64_t n_out[4], uint64_t total_out[4], unsigned n_parts)
{
  (void) word_length;
  std::vector<uint64_t> bounds (2 * (n_parts + 1));
  {
    const void *keys[2] = {rec_a, rec_b};
    const size_t strides[2] = {12, 12};
    const uint64_t sizes[2] = {n_a, n_b};
    int rc = gt4gpu_plan_splitters (keys, strides, sizes, 2, n_parts, bounds.data (), nullptr);
    if (rc) return rc;
  }
  const uint64_t *ba = bounds.data (), *bb = bounds.data () + n_parts + 1;
  uint64_t max_a = 0, max_b = 0;
  for (unsigned p = 0; p < n_parts; p++) {
    max_a = std::max (max_a, ba[p + 1] - ba[p]);
    max_b = std::max (max_b, bb[p + 1] - bb[p]);
  }
  SetOpParams prm;
  memset (&prm, 0, sizeof (prm));
  prm.ops = ops; prm.cutoff = cutoff; prm.count_override = count_override; prm.subtract = subtract ? 1 : 0; prm.sem = SEM_PAIR;
  for (int s = 0; s < 4; s++) prm.rule[s] = resolve_rule (rule, s);

  // two sets of device buffers
  struct Set {
    void *aos_a = nullptr, *aos_b = nullptr;
    uint64_t *wa = nullptr, *wb = nullptr;
    uint32_t *ca = nullptr, *cb = nullptr;
    uint64_t *ow[4] = {nullptr, nullptr, nullptr, nullptr};
    uint32_t *oc[4] = {nullptr, nullptr, nullptr, nullptr};
    void *oaos[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t in_done = nullptr, merged = nullptr, out_done = nullptr;
  } set[2];
  uint64_t cap[4];
  for (int s = 0; s < 4; s++) cap[s] = worst_case (prm, s, max_a, max_b);
  int rc = 0;
  auto alloc = [&] (void **p, size_t bytes) { if (!rc) rc = dev_alloc (p, bytes); };
  for (int b = 0; b < 2; b++) {
    alloc (&set[b].aos_a, max_a * 12); alloc (&set[b].aos_b, max_b * 12);
    alloc ((void **) &set[b].wa, max_a * 8); alloc ((void **) &set[b].ca, max_a * 4);
    alloc ((void **) &set[b].wb, max_b * 8); alloc ((void **) &set[b].cb, max_b * 4);
    for (int s = 0; s < 4; s++) {
      if (!((ops >> s) & 1u) || countonly) continue;
      alloc ((void **) &set[b].ow[s], cap[s] * 8); alloc ((void **) &set[b].oc[s], cap[s] * 4); alloc (&set[b].oaos[s], cap[s] * 12);
    }
    if (!rc && (cudaEventCreateWithFlags (&set[b].in_done, cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags (&set[b].merged, cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags (&set[b].out_done, cudaEventDisableTiming) != cudaSuccess))
      rc = fail (GT4GPU_ERR_CUDA, "cudaEventCreate failed");
  }
  cudaError_t e = cudaSuccess;
  if (!rc) e = cudaStreamSynchronize (g_ctx.stream);     // the allocations above are ordered on the compute stream

  auto enqueue_in = [&] (unsigned p) {
    Set &st = set[p & 1];
    const uint64_t na = ba[p + 1] - ba[p], nb = bb[p + 1] - bb[p];
    // the merge that last read this set (part p - 2) has completed: merge2_device synchronises
    if (na) {
      e = cudaMemcpyAsync (st.aos_a, rec_a + ba[p] * 12, na * 12, cudaMemcpyHostToDevice, g_ctx.in_stream);
      if (e == cudaSuccess) e = launch_deinterleave (st.aos_a, na, st.wa, st.ca, g_ctx.in_stream);
    }
    if (e == cudaSuccess && nb) {
      e = cudaMemcpyAsync (st.aos_b, rec_b + bb[p] * 12, nb * 12, cudaMemcpyHostToDevice, g_ctx.in_stream);
      if (e == cudaSuccess) e = launch_deinterleave (st.aos_b, nb, st.wb, st.cb, g_ctx.in_stream);
    }
    if (e == cudaSuccess) e = cudaEventRecord (st.in_done, g_ctx.in_stream);
  };

  float ms_p = 0.f, ms_m = 0.f;
  uint32_t nl = 0;
  for (int s = 0; s < 4; s++) n_out[s] = total_out[s] = 0;
  if (!rc && e == cudaSuccess) enqueue_in (0);
  for (unsigned p = 0; p < n_parts && !rc && e == cudaSuccess; p++) {
    Set &st = set[p & 1];
    if (p + 1 < n_parts) enqueue_in (p + 1);
    if (e != cudaSuccess) break;
    e = cudaStreamWaitEvent (g_ctx.stream, st.in_done, 0);
    if (e == cudaSuccess && p >= 2) e = cudaStreamWaitEvent (g_ctx.stream, st.out_done, 0);   // its output buffers were copied out
    if (e != cudaSuccess) break;
    MergeOut mo[4];
    for (int s = 0; s < 4; s++) {
      if (!((ops >> s) & 1u) || countonly) continue;
      mo[s].caller = true; mo[s].words = st.ow[s]; mo[s].counts = st.oc[s]; mo[s].capacity = cap[s];
    }
    reset_timing ();
    rc = merge2_device (DevList{st.wa, st.ca, ba[p + 1] - ba[p]}, DevList{st.wb, st.cb, bb[p + 1] - bb[p]}, prm, ops, countonly != 0, mo);
    ms_p += tl_ms_partition; ms_m += tl_ms_merge; nl += tl_launches;
    if (rc) break;
    for (int s = 0; s < 4; s++) {
      if (!((ops >> s) & 1u)) continue;
      if (!countonly && mo[s].n) {
        if (!out_records || !out_capacity || !out_records[s]) { rc = fail (GT4GPU_ERR_ARG, "missing output buffer for stream %d", s); break; }
        if (n_out[s] + mo[s].n > out_capacity[s]) { rc = fail (GT4GPU_ERR_CAPACITY, "stream %d needs more than %llu records", s, (unsigned long long) out_capacity[s]); break; }
        e = launch_interleave (st.ow[s], st.oc[s], mo[s].n, st.oaos[s], g_ctx.out_stream);     // merge p is complete (synchronised)
        if (e == cudaSuccess)
          e = cudaMemcpyAsync (static_cast<unsigned char *> (out_records[s]) + n_out[s] * 12, st.oaos[s], mo[s].n * 12, cudaMemcpyDeviceToHost, g_ctx.out_stream);
        if (e != cudaSuccess) break;
      }
      n_out[s] += mo[s].n;
      total_out[s] += mo[s].sum;
    }
    if (e == cudaSuccess) e = cudaEventRecord (st.out_done, g_ctx.out_stream);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize (g_ctx.out_stream);
  cudaStreamSynchronize (g_ctx.in_stream);
  for (int b = 0; b < 2; b++) {
    dev_free (set[b].aos_a); dev_free (set[b].aos_b); dev_free (set[b].wa); dev_free (set[b].ca); dev_free (set[b].wb); dev_free (set[b].cb);
    for (int s = 0; s < 4; s++) { dev_free (set[b].ow[s]); dev_free (set[b].oc[s]); dev_free (set[b].oaos[s]); }
    if (set[b].in_done) cudaEventDestroy (set[b].in_done);
    if (set[b].merged) cudaEventDestroy (set[b].merged);
    if (set[b].out_done) cudaEventDestroy (set[b].out_done);
  }
  tl_ms_partition = ms_p; tl_ms_merge = ms_m; tl_launches = nl;
  if (!rc && e != cudaSuccess) rc = fail (GT4GPU_ERR_CUDA, "pipelined host path: %s", cudaGetErrorString (e));
  return rc;
}

int gt4gpu_compare2_host_aos (const void *records_a, uint64_t n_a, const void *records_b, uint64_t n_b,
                              uint32_t word_length, uint32_t ops, int rule, uint32_t cutoff,
                              uint32_t count_override, int subtract, int countonly,
                              void *const out_records[4], const uint64_t out_capacity[4],
                              uint64_t n_out[4], uint64_t total_out[4])
{
  if (!n_out || !total_out) return fail (GT4GPU_ERR_ARG, "null argument");
  if (!ops || (ops & ~15u)) return fail (GT4GPU_ERR_ARG, "ops must be a non-empty OR of GT4GPU_OP_*");
  if (rule < GT4GPU_RULE_DEFAULT || rule > GT4GPU_RULE_NUMBER) return fail (GT4GPU_ERR_ARG, "unknown rule %d", rule);
  if ((n_a && !records_a) || (n_b && !records_b)) return fail (GT4GPU_ERR_ARG, "null argument");
  int rc = ensure_ready ();
  if (rc) return rc;
  const uint64_t total = n_a + n_b;
  uint64_t part_records = 64ull << 20;             // records per part (768 MiB of input)
  if (const char *env = getenv ("GT4GPU_HOST_PART_RECORDS")) part_records = strtoull (env, nullptr, 10);
  if (part_records < 1024) part_records = 1024;
  const unsigned n_parts = (unsigned) std::min<uint64_t> ((total + part_records - 1) / part_records, 256);
  if (n_parts < 2)
    return compare2_host_simple (records_a, n_a, records_b, n_b, word_length, ops, rule, cutoff, count_override, subtract, countonly,
                                 out_records, out_capacity, n_out, total_out);
  return compare2_host_pipelined (static_cast<const unsigned char *> (records_a), n_a, static_cast<const unsigned char *> (records_b), n_b,
                                  word_length, ops, rule, cutoff, count_override, subtract, countonly,
                                  out_records, out_capacity, n_out, total_out, n_parts);
}


// ------------------------------------------------------------------ file-to-file, pipelined

namespace {

// One direction of the file pipeline: records [first, first + n) of a mapped list file -> device AoS buffer, through two
// pinned bounce buffers (filled by a few threads) so that the copy out of the page cache overlaps the H2D transfer.
struct FileCopyBuffers {
  void *pinned[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  int init ()
  {
    for (int b = 0; b < 2; b++) {
      if (!(pinned[b] = g_pinned.take ())) return fail (GT4GPU_ERR_CUDA, "cudaMallocHost failed");
      if (cudaEventCreateWithFlags (&ev[b], cudaEventDisableTiming) != cudaSuccess) return fail (GT4GPU_ERR_CUDA, "cudaEventCreate failed");
    }
    return 0;
  }
  ~FileCopyBuffers ()
  {
    for (int b = 0; b < 2; b++) {
      g_pinned.give (pinned[b]);
      if (ev[b]) cudaEventDestroy (ev[b]);
    }
  }
};

constexpr uint64_t FILE_CHUNK = BOUNCE_BYTES / 12;

// GT4GPU_FILE_TIMING=1: wall-clock seconds of the pipeline's phases on stderr (each counter has one writer thread)
struct FileTiming {
  bool on = false;
  double plan = 0, up_copy = 0, up_wait = 0, up_sync = 0, down_wait = 0, down_write = 0, merge = 0, join_up = 0, join_down = 0, total = 0;
};
FileTiming g_ft;
inline double now_s ()
{
  struct timespec ts;
  clock_gettime (CLOCK_MONOTONIC, &ts);
  return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

cudaError_t upload_records (const unsigned char *src, uint64_t n, void *d_aos, uint64_t *d_words, uint32_t *d_counts, FileCopyBuffers &buf,
                            unsigned n_threads, cudaStream_t st)
{
  cudaError_t e = cudaSuccess;
  int b = 0;
  for (uint64_t done = 0; done < n && e == cudaSuccess; done += FILE_CHUNK, b ^= 1) {
    const uint64_t m = std::min (FILE_CHUNK, n - done);
    const double t0 = g_ft.on ? now_s () : 0;
    if (done >= 2 * FILE_CHUNK) e = cudaEventSynchronize (buf.ev[b]);
    if (e != cudaSuccess) break;
    const double t1 = g_ft.on ? now_s () : 0;
    copy_parallel (buf.pinned[b], src + done * 12, m * 12, n_threads);
    if (g_ft.on) { const double t2 = now_s (); g_ft.up_wait += t1 - t0; g_ft.up_copy += t2 - t1; }
    e = cudaMemcpyAsync (static_cast<unsigned char *> (d_aos) + done * 12, buf.pinned[b], m * 12, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaEventRecord (buf.ev[b], st);
  }
  const double t3 = g_ft.on ? now_s () : 0;
  if (e == cudaSuccess && n) e = launch_deinterleave (d_aos, n, d_words, d_counts, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize (st);
  if (g_ft.on) g_ft.up_sync += now_s () - t3;
  return e;
}

// device SoA result -> records in the file at byte offset `at`, chunk i + 1 copied out while chunk i is written
int download_records (const uint64_t *d_words, const uint32_t *d_counts, uint64_t n, void *d_aos, int fd, int64_t at, FileCopyBuffers &buf, cudaStream_t st)
{
  if (n == 0) return 0;
  cudaError_t e = launch_interleave (d_words, d_counts, n, d_aos, st);
  auto enqueue = [&] (uint64_t done, int b) {
    const uint64_t m = std::min (FILE_CHUNK, n - done);
    e = cudaMemcpyAsync (buf.pinned[b], static_cast<unsigned char *> (d_aos) + done * 12, m * 12, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaEventRecord (buf.ev[b], st);
  };
  int rc = 0, b = 0;
  if (e == cudaSuccess) enqueue (0, 0);
  for (uint64_t done = 0; done < n && !rc && e == cudaSuccess; done += FILE_CHUNK, b ^= 1) {
    if (done + FILE_CHUNK < n) enqueue (done + FILE_CHUNK, b ^ 1);
    if (e != cudaSuccess) break;
    const double t0 = g_ft.on ? now_s () : 0;
    e = cudaEventSynchronize (buf.ev[b]);
    if (e != cudaSuccess) break;
    const double t1 = g_ft.on ? now_s () : 0;
    rc = write_all (fd, buf.pinned[b], std::min (FILE_CHUNK, n - done) * 12, at + (int64_t) (done * 12));
    if (g_ft.on) { g_ft.down_wait += t1 - t0; g_ft.down_write += now_s () - t1; }
  }
  cudaStreamSynchronize (st);
  if (!rc && e != cudaSuccess) rc = fail (GT4GPU_ERR_CUDA, "file output: %s", cudaGetErrorString (e));
  return rc;
}

}  // namespace

int gt4gpu_compare2_files (const char *path_a, const char *path_b, int stream_mode, uint32_t ops, int rule, uint32_t cutoff,
                           uint32_t count_override, int subtract, int countonly, const int out_fd[4],
                           uint64_t n_out[4], uint64_t total_out[4], uint32_t *word_length)
{
  if (!path_a || !path_b || !n_out || !total_out) return fail (GT4GPU_ERR_ARG, "null argument");
  if (!ops || (ops & ~15u)) return fail (GT4GPU_ERR_ARG, "ops must be a non-empty OR of GT4GPU_OP_*");
  if (rule < GT4GPU_RULE_DEFAULT || rule > GT4GPU_RULE_NUMBER) return fail (GT4GPU_ERR_ARG, "unknown rule %d", rule);
  if (!countonly && !out_fd) return fail (GT4GPU_ERR_ARG, "missing output descriptors");
  Mapping ma, mb;
  int rc = map_file (path_a, ma);
  if (!rc) rc = map_file (path_b, mb);
  if (rc) return rc;
  if (is_index_file (ma.data, ma.size) || is_index_file (mb.data, mb.size))
    return fail (GT4GPU_ERR_ARG, "gt4gpu_compare2_files reads list files only (open index files with gt4gpu_list_open)");
  gt4gpu_header ha, hb;
  if ((rc = parse_header (ma.data, ma.size, stream_mode, path_a, &ha))) return rc;
  if ((rc = parse_header (mb.data, mb.size, stream_mode, path_b, &hb))) return rc;
  if ((rc = ensure_ready ())) return rc;
  if (word_length) *word_length = ha.word_length;           // output header word length = first list's (src/glistcompare.c:814)
  const unsigned char *rec_a = ma.data + ha.list_start, *rec_b = mb.data + hb.list_start;
  const uint64_t n_a = ha.n_words, n_b = hb.n_words, total = n_a + n_b;
  for (int s = 0; s < 4; s++) n_out[s] = total_out[s] = 0;

  g_ft = FileTiming ();
  g_ft.on = getenv ("GT4GPU_FILE_TIMING") != nullptr;
  const double t_begin = g_ft.on ? now_s () : 0;
  uint64_t part_records = 64ull << 20;
  if (const char *env = getenv ("GT4GPU_HOST_PART_RECORDS")) part_records = strtoull (env, nullptr, 10);
  if (part_records < 1024) part_records = 1024;
  const unsigned n_parts = (unsigned) std::max<uint64_t> (1, std::min<uint64_t> ((total + part_records - 1) / part_records, 1024));
  std::vector<uint64_t> bounds (2 * (n_parts + 1));
  {
    const void *keys[2] = {rec_a, rec_b};
    const size_t strides[2] = {12, 12};
    const uint64_t sizes[2] = {n_a, n_b};
    if ((rc = gt4gpu_plan_splitters (keys, strides, sizes, 2, n_parts, bounds.data (), nullptr))) return rc;
  }
  if (g_ft.on) g_ft.plan = now_s () - t_begin;
  const uint64_t *ba = bounds.data (), *bb = bounds.data () + n_parts + 1;
  uint64_t max_a = 0, max_b = 0;
  for (unsigned p = 0; p < n_parts; p++) {
    max_a = std::max (max_a, ba[p + 1] - ba[p]);
    max_b = std::max (max_b, bb[p + 1] - bb[p]);
  }
  SetOpParams prm;
  memset (&prm, 0, sizeof (prm));
  prm.ops = ops; prm.cutoff = cutoff; prm.count_override = count_override; prm.subtract = subtract ? 1 : 0; prm.sem = SEM_PAIR;
  for (int s = 0; s < 4; s++) prm.rule[s] = resolve_rule (rule, s);

  struct Set {
    void *aos_a = nullptr, *aos_b = nullptr;
    uint64_t *wa = nullptr, *wb = nullptr;
    uint32_t *ca = nullptr, *cb = nullptr;
    uint64_t *ow[4] = {nullptr, nullptr, nullptr, nullptr};
    uint32_t *oc[4] = {nullptr, nullptr, nullptr, nullptr};
    uint64_t n[4] = {0, 0, 0, 0};
  } set[2];
  void *oaos = nullptr;
  uint64_t cap[4], cap_max = 0;
  for (int s = 0; s < 4; s++) { cap[s] = worst_case (prm, s, max_a, max_b); if ((ops >> s) & 1u) cap_max = std::max (cap_max, cap[s]); }
  auto alloc = [&] (void **p, size_t bytes) { if (!rc) rc = dev_alloc (p, bytes); };
  for (int b = 0; b < 2; b++) {
    alloc (&set[b].aos_a, max_a * 12); alloc (&set[b].aos_b, max_b * 12);
    alloc ((void **) &set[b].wa, max_a * 8); alloc ((void **) &set[b].ca, max_a * 4);
    alloc ((void **) &set[b].wb, max_b * 8); alloc ((void **) &set[b].cb, max_b * 4);
    for (int s = 0; s < 4; s++) {
      if (!((ops >> s) & 1u) || countonly) continue;
      alloc ((void **) &set[b].ow[s], cap[s] * 8); alloc ((void **) &set[b].oc[s], cap[s] * 4);
    }
  }
  if (!countonly) alloc (&oaos, cap_max * 12);
  FileCopyBuffers in_buf, out_buf;
  if (!rc) rc = in_buf.init ();
  if (!rc && !countonly) rc = out_buf.init ();
  cudaError_t e = cudaSuccess;
  if (!rc) e = cudaStreamSynchronize (g_ctx.stream);     // the allocations above are ordered on the compute stream

  const int device = g_ctx.device;
  const unsigned copy_threads = std::min (16u, std::max (2u, std::thread::hardware_concurrency () / 2));
  cudaError_t e_up = cudaSuccess;
  int rc_down = 0;
  auto upload = [&] (unsigned p) {
    cudaSetDevice (device);
    Set &st = set[p & 1];
    e_up = upload_records (rec_a + ba[p] * 12, ba[p + 1] - ba[p], st.aos_a, st.wa, st.ca, in_buf, copy_threads, g_ctx.in_stream);
    if (e_up == cudaSuccess) e_up = upload_records (rec_b + bb[p] * 12, bb[p + 1] - bb[p], st.aos_b, st.wb, st.cb, in_buf, copy_threads, g_ctx.in_stream);
  };
  uint64_t written[4] = {0, 0, 0, 0};
  auto download = [&] (unsigned p) {
    cudaSetDevice (device);
    Set &st = set[p & 1];
    for (int s = 0; s < 4 && !rc_down; s++) {
      if (!((ops >> s) & 1u) || countonly) continue;
      rc_down = download_records (st.ow[s], st.oc[s], st.n[s], oaos, out_fd[s], (int64_t) (sizeof (gt4gpu_header) + 12 * written[s]), out_buf, g_ctx.out_stream);
      written[s] += st.n[s];
    }
  };
  float ms_p = 0.f, ms_m = 0.f;
  uint32_t nl = 0;
  if (!rc && e == cudaSuccess) { upload (0); e = e_up; }
  for (unsigned p = 0; p < n_parts && !rc && e == cudaSuccess; p++) {
    // three things at once: part p + 1 comes in, part p is merged, part p - 1 goes out
    std::thread t_up, t_down;
    if (p + 1 < n_parts) t_up = std::thread (upload, p + 1);
    if (p >= 1 && !countonly) t_down = std::thread (download, p - 1);
    Set &st = set[p & 1];
    MergeOut mo[4];
    for (int s = 0; s < 4; s++) {
      if (!((ops >> s) & 1u) || countonly) continue;
      mo[s].caller = true; mo[s].words = st.ow[s]; mo[s].counts = st.oc[s]; mo[s].capacity = cap[s];
    }
    reset_timing ();
    const double tm0 = g_ft.on ? now_s () : 0;
    rc = merge2_device (DevList{st.wa, st.ca, ba[p + 1] - ba[p]}, DevList{st.wb, st.cb, bb[p + 1] - bb[p]}, prm, ops, countonly != 0, mo);
    const double tm1 = g_ft.on ? now_s () : 0;
    ms_p += tl_ms_partition; ms_m += tl_ms_merge; nl += tl_launches;
    for (int s = 0; s < 4 && !rc; s++) {
      if (!((ops >> s) & 1u)) continue;
      st.n[s] = mo[s].n;
      n_out[s] += mo[s].n;
      total_out[s] += mo[s].sum;
    }
    if (t_up.joinable ()) t_up.join ();
    const double tm2 = g_ft.on ? now_s () : 0;
    if (t_down.joinable ()) t_down.join ();
    if (g_ft.on) { g_ft.merge += tm1 - tm0; g_ft.join_up += tm2 - tm1; g_ft.join_down += now_s () - tm2; }
    if (!rc && rc_down) rc = rc_down;
    if (e_up != cudaSuccess) e = e_up;
  }
  if (!rc && e == cudaSuccess && !countonly) { download (n_parts - 1); rc = rc_down; }
  for (int b = 0; b < 2; b++) {
    dev_free (set[b].aos_a); dev_free (set[b].aos_b); dev_free (set[b].wa); dev_free (set[b].ca); dev_free (set[b].wb); dev_free (set[b].cb);
    for (int s = 0; s < 4; s++) { dev_free (set[b].ow[s]); dev_free (set[b].oc[s]); }
  }
  dev_free (oaos);
  tl_ms_partition = ms_p; tl_ms_merge = ms_m; tl_launches = nl;
  if (!rc && e != cudaSuccess) rc = fail (GT4GPU_ERR_CUDA, "file pipeline: %s", cudaGetErrorString (e));
  if (rc) return rc;
  // the headers last, like the reference (header with totals rewritten at offset 0, src/glistcompare.c:907-953)
  for (int s = 0; s < 4 && !countonly; s++) {
    if (!((ops >> s) & 1u)) continue;
    gt4gpu_header h;
    gt4gpu_header_init (&h, ha.word_length);
    h.n_words = n_out[s];
    h.total_count = total_out[s];
    if ((rc = write_all (out_fd[s], &h, sizeof (h), 0))) return rc;
  }
  if (g_ft.on)
    fprintf (stderr, "gt4gpu file pipeline: %u parts, total %.3f s: plan %.3f | main thread: merge %.3f, waits for the upload %.3f, for the download %.3f | "
                     "upload thread: page cache -> pinned %.3f, waits for H2D %.3f, last H2D + de-interleave %.3f | download thread: waits for D2H %.3f, pwrite %.3f\n",
             n_parts, now_s () - t_begin, g_ft.plan, g_ft.merge, g_ft.join_up, g_ft.join_down, g_ft.up_copy, g_ft.up_wait, g_ft.up_sync, g_ft.down_wait, g_ft.down_write);
  return 0;
}

// ------------------------------------------------------------------ sharding plan (host only)

int gt4gpu_plan_splitters (const void *const *keys, const size_t *stride_bytes, const uint64_t *n_words,
                           unsigned n_lists, unsigned n_parts, uint64_t *bounds, uint64_t *splitters)
{
  if (!keys || !stride_bytes || !n_words || !bounds || n_lists == 0 || n_parts == 0) return fail (GT4GPU_ERR_ARG, "bad argument");
  auto key_at = [&] (unsigned j, uint64_t i) {
    uint64_t v;
    memcpy (&v, static_cast<const unsigned char *> (keys[j]) + i * stride_bytes[j], sizeof (v));
    return v;
  };
  // number of records of list j with key < K
  auto below = [&] (unsigned j, uint64_t K) {
    uint64_t lo = 0, hi = n_words[j];
    while (lo < hi) {
      const uint64_t mid = lo + ((hi - lo) >> 1);
      if (key_at (j, mid) < K) lo = mid + 1;
      else hi = mid;
    }
    return lo;
  };
  uint64_t total = 0;
  for (unsigned j = 0; j < n_lists; j++) total += n_words[j];
  const unsigned stride = n_parts + 1;
  for (unsigned j = 0; j < n_lists; j++) {
    bounds[j * stride] = 0;
    bounds[j * stride + n_parts] = n_words[j];
  }
  for (unsigned p = 1; p < n_parts; p++) {
    // exact co-rank on key VALUES: the smallest K with  sum_j |{x in L_j : x < K}| >= target
    const uint64_t target = (uint64_t) (((unsigned __int128) total * p) / n_parts);
    uint64_t lo = 0, hi = UINT64_MAX;
    while (lo < hi) {
      const uint64_t mid = lo + ((hi - lo) >> 1);
      uint64_t cnt = 0;
      for (unsigned j = 0; j < n_lists; j++) cnt += below (j, mid);
      if (cnt >= target) hi = mid;
      else lo = mid + 1;
    }
    if (splitters) splitters[p - 1] = lo;
    for (unsigned j = 0; j < n_lists; j++) bounds[j * stride + p] = below (j, lo);
  }
  return 0;
}

// ------------------------------------------------------------------ helpers for harnesses

int gt4gpu_deinterleave (const void *d_records, uint64_t n, uint64_t *d_words, uint32_t *d_counts)
{
  int rc = ensure_ready ();
  if (rc) return rc;
  CU (launch_deinterleave (d_records, n, d_words, d_counts, g_ctx.stream));
  CU (cudaStreamSynchronize (g_ctx.stream));
  return 0;
}

int gt4gpu_interleave (const uint64_t *d_words, const uint32_t *d_counts, uint64_t n, void *d_records)
{
  int rc = ensure_ready ();
  if (rc) return rc;
  CU (launch_interleave (d_words, d_counts, n, d_records, g_ctx.stream));
  CU (cudaStreamSynchronize (g_ctx.stream));
  return 0;
}

}  // extern "C"
