// emulate_tile.cpp -- TEST HELPER.  Runs the tile decomposition of setop2_tile_kernel on the
// CPU, thread by thread, using the very same __host__ __device__ functions the kernel uses
// (genometester4_b200/csrc/gt4gpu_core.cuh): merge_path, merge_slots, eval_stream.  What it
// does NOT cover is the CUDA glue (staging loads, block scan, look-back, stores) -- that is
// what the -m gpu tests are for.  Built by tests/test_core_emulation.py with g++.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "gt4gpu_core.cuh"

using namespace gt4gpu;

template <int VT>
static void run (const uint64_t *aw, const uint32_t *ac, uint64_t na, const uint64_t *bw, const uint32_t *bc, uint64_t nb,
                 int nt, const SetOpParams &p, uint32_t mask, uint64_t *const ow[4], uint32_t *const oc[4],
                 uint64_t n_out[4], uint64_t sum_out[4])
{
  const uint64_t tile = (uint64_t) nt * VT;
  const uint64_t total = na + nb;
  const uint64_t n_tiles = (total + tile - 1) / tile;
  std::vector<uint64_t> part (n_tiles + 1);
  for (uint64_t t = 0; t <= n_tiles; t++) {
    uint64_t diag = t * tile;
    if (diag > total) diag = total;
    part[t] = merge_path<uint64_t> (aw, na, bw, nb, diag);
  }
  const int slots = (int) tile + VT + 5;
  std::vector<uint64_t> sk (slots);
  std::vector<uint32_t> sc (slots);
  for (uint64_t t = 0; t < n_tiles; t++) {
    const uint64_t d_lo = t * tile, d_hi = (d_lo + tile < total) ? d_lo + tile : total;
    const uint64_t a_lo = part[t], a_hi = part[t + 1];
    const uint64_t b_lo = d_lo - a_lo, b_hi = d_hi - a_hi;
    const int tna = (int) (a_hi - a_lo), tnb = (int) (b_hi - b_lo);
    const bool has_halo = a_lo > 0, has_peek = b_hi < nb, has_peek_a = a_hi < na;
    // poison the slack so that any use of it shows up
    for (int x = 0; x < slots; x++) { sk[x] = 0xDEADBEEFDEADBEEFull ^ (uint64_t) x; sc[x] = 0xABCD0000u + x; }
    // layout: [halo of A][A slice][element of A after the tile][B slice][peek of B]
    for (int x = 0; x < tna + tnb + 3; x++) {
      uint64_t k = 0; uint32_t c = 0;
      if (x <= tna + 1) { if ((x > 0 || has_halo) && (x <= tna || has_peek_a)) { k = aw[a_lo + x - 1]; c = ac[a_lo + x - 1]; } }
      else { int j = x - 2 - tna; if (j < tnb || has_peek) { k = bw[b_lo + j]; c = bc[b_lo + j]; } }
      sk[x] = k; sc[x] = c;
    }
    const uint64_t *ka = sk.data () + 1; const uint32_t *ca = sc.data () + 1;
    const uint64_t *kb = ka + tna + 1; const uint32_t *cb = ca + tna + 1;
    const int n_tile = tna + tnb;
    const bool interior = has_halo && has_peek && has_peek_a && (uint64_t) n_tile == tile;   // what the stream kernel's producer flags
    // coarse co-ranks every 8 threads (what the splitter warp of setop2_stream_kernel computes)
    const int n_split = (nt + 7) / 8 + 1;
    std::vector<int> split (n_split);
    for (int g = 0; g < n_split; g++) {
      const int dg = (g * 8 * VT < n_tile) ? g * 8 * VT : n_tile;
      split[g] = merge_path<int> (ka, tna, kb, tnb, dg);
    }
    for (int tid = 0; tid < nt; tid++) {
      const int d0 = (tid * VT < n_tile) ? tid * VT : n_tile;
      const int i0 = merge_path_window<int> (ka, tna, kb, tnb, d0, split[tid / 8], split[tid / 8 + 1]);
      if (i0 != merge_path<int> (ka, tna, kb, tnb, d0)) __builtin_trap ();
      auto sink = [&] (int, uint64_t key, uint32_t c1, uint32_t c2, bool in_a, bool in_b, bool live) {
          for (int q = 0; q < 4; q++) {
            if (!((mask >> q) & 1u)) continue;
            uint32_t f = 0;
            if (live && eval_stream (p, q, c1, c2, in_a, in_b, f)) {
              if (ow[q]) { ow[q][n_out[q]] = key; oc[q][n_out[q]] = f; }
              n_out[q] += 1;
              sum_out[q] += f;
            }
          }
        };
      if (interior) merge_slots_interior<VT> (ka, ca, kb, cb, i0, d0, sink);
      else merge_slots<VT> (ka, ca, tna, has_halo, kb, cb, tnb, has_peek, i0, d0, sink);
    }
  }
}

extern "C" int emu_setop2 (const uint64_t *aw, const uint32_t *ac, uint64_t na, const uint64_t *bw, const uint32_t *bc, uint64_t nb,
                           int nt, int vt, uint32_t mask, int sem, int rule, uint32_t cutoff, uint32_t ov, int subtract,
                           uint64_t *const ow[4], uint32_t *const oc[4], uint64_t n_out[4], uint64_t sum_out[4])
{
  SetOpParams p;
  memset (&p, 0, sizeof (p));
  p.ops = mask; p.cutoff = cutoff; p.count_override = ov; p.subtract = subtract; p.sem = sem;
  for (int s = 0; s < 4; s++) p.rule[s] = (sem == SEM_PAIR) ? resolve_rule (rule, s) : rule;
  for (int q = 0; q < 4; q++) n_out[q] = sum_out[q] = 0;
  switch (vt) {
  case 1: run<1> (aw, ac, na, bw, bc, nb, nt, p, mask, ow, oc, n_out, sum_out); return 0;
  case 3: run<3> (aw, ac, na, bw, bc, nb, nt, p, mask, ow, oc, n_out, sum_out); return 0;
  case 7: run<7> (aw, ac, na, bw, bc, nb, nt, p, mask, ow, oc, n_out, sum_out); return 0;
  case 9: run<9> (aw, ac, na, bw, bc, nb, nt, p, mask, ow, oc, n_out, sum_out); return 0;
  default: return 1;
  }
}

// eval_fast<F> must agree with eval_stream wherever select_fast_path picks F.  Returns the number of
// disagreements over a grid of small and extreme values (0 = pass).
template <int F>
static uint64_t check_fast (const SetOpParams &p, int stream)
{
  static const uint32_t vals[] = {0u, 1u, 2u, 3u, 4u, 5u, 6u, 100u, 0x7fffffffu, 0x80000000u, 0x80000001u, 0xfffffffeu, 0xffffffffu};
  uint64_t bad = 0;
  for (uint32_t c1 : vals) for (uint32_t c2 : vals) for (int ia = 0; ia < 2; ia++) for (int ib = 0; ib < 2; ib++) {
    if (!ia && !ib) continue;
    uint32_t f0 = 0, f1 = 0;
    const bool k0 = eval_stream (p, stream, c1, c2, ia, ib, f0);
    const bool k1 = eval_fast<F> (p, stream, c1, c2, ia, ib, f1);
    if (k0 != k1 || (k0 && f0 != f1)) bad++;
  }
  return bad;
}

extern "C" uint64_t emu_check_fast_paths (void)
{
  static const uint32_t cutoffs[] = {0u, 1u, 2u, 5u, 100u, 0x80000000u, 0xffffffffu};
  uint64_t bad = 0, selected[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int sem = 0; sem < 5; sem++) for (int rule = 0; rule < 8; rule++) for (int sub = 0; sub < 2; sub++)
    for (uint32_t cutoff : cutoffs) for (int stream = 0; stream < 4; stream++) {
      SetOpParams p;
      memset (&p, 0, sizeof (p));
      p.ops = 1u << stream; p.cutoff = cutoff; p.count_override = 7; p.subtract = sub; p.sem = sem;
      for (int s = 0; s < 4; s++) p.rule[s] = (sem == SEM_PAIR) ? resolve_rule (rule, s) : (rule == 0 ? 1 : rule);
      const int f = select_fast_path (p, stream);
      selected[f]++;
      switch (f) {
      case FAST_U_ADD: bad += check_fast<FAST_U_ADD> (p, stream); break;
      case FAST_I_MIN: bad += check_fast<FAST_I_MIN> (p, stream); break;
      case FAST_D_SUB: bad += check_fast<FAST_D_SUB> (p, stream); break;
      case FAST_NU_ADD: bad += check_fast<FAST_NU_ADD> (p, stream); break;
      case FAST_NI_MIN: bad += check_fast<FAST_NI_MIN> (p, stream); break;
      case FAST_D2_SUB: bad += check_fast<FAST_D2_SUB> (p, stream); break;
      default: bad += check_fast<FAST_GENERIC> (p, stream); break;
      }
    }
  for (int f = 1; f < 7; f++) if (!selected[f]) bad += 1000000;   // every fast path must be reachable
  return bad;
}
