// gt4gpu_core.cuh -- arithmetic shared by every kernel of the engine: the count
// rules, the per-output predicates, the merge-path co-rank search and the
// per-thread serial merge.  Everything here is integer-only and
// __host__ __device__, so the unit tests can drive exactly the code the
// kernels run (tests/emulate_tile.cpp) without a GPU.
//
// Semantics follow /root/reference/src/glistcompare.c:433-489 (rules and
// predicates), :843-905 (two-list loop), :545-591 / :647-705 (N-list loops).
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define GT4_HD __host__ __device__ __forceinline__
#define GT4_UNROLL _Pragma ("unroll")
#else
#define GT4_HD inline
#define GT4_UNROLL
#endif

namespace gt4gpu {

enum Rule : int {
  RULE_DEFAULT = 0, RULE_ADD = 1, RULE_SUBTRACT = 2, RULE_MIN = 3, RULE_MAX = 4,
  RULE_FIRST = 5, RULE_SECOND = 6, RULE_NUMBER = 7
};

enum : uint32_t { OP_UNION = 1u, OP_INTRSEC = 2u, OP_DIFF = 4u, OP_DDIFF = 8u };

// How a merged key's (f1, f2) pair is turned into output records.
enum Semantics : int {
  SEM_PAIR = 0,            // compare_wordmaps: up to 4 streams, "!= 0" filter, cutoff on the inputs
  SEM_NUNION_PARTIAL = 1,  // inner node of an N-list union: combine, keep everything
  SEM_NUNION_FINAL = 2,    // last node of an N-list union: combine, keep iff combined >= cutoff
  SEM_NISECT_PARTIAL = 3,  // inner node of an N-list intersection chain
  SEM_NISECT_FINAL = 4
};

struct SetOpParams {
  uint32_t ops;             // requested streams (SEM_PAIR); N-list semantics use stream 0 only
  uint32_t cutoff;
  uint32_t count_override;  // RULE_NUMBER value (glistcompare.c:82)
  int rule[4];              // resolved rule per stream (RULE_DEFAULT already substituted)
  int subtract;             // -du
  int sem;
};

// calculate_freq, glistcompare.c:433-455
GT4_HD uint32_t calc_freq (uint32_t f1, uint32_t f2, int rule, uint32_t ov)
{
  switch (rule) {
  case RULE_ADD:      return f1 + f2;
  case RULE_SUBTRACT: return (f1 > f2) ? f1 - f2 : 0u;
  case RULE_MIN:      return (f1 < f2) ? f1 : f2;
  case RULE_MAX:      return (f1 > f2) ? f1 : f2;
  case RULE_FIRST:    return f1;
  case RULE_SECOND:   return f2;
  case RULE_NUMBER:   return ov;
  default:            return 0u;
  }
}

// Resolve RULE_DEFAULT per stream the way include_in_* do (glistcompare.c:462,471,485).
GT4_HD int resolve_rule (int rule, int stream)
{
  if (rule != RULE_DEFAULT) return rule;
  return stream == 0 ? RULE_ADD : stream == 1 ? RULE_MIN : RULE_SUBTRACT;
}

// One merged key -> does stream `s` emit it, and with which count?
// c1/c2 are the stored counts, in_a/in_b say which lists hold the key; an absent side
// contributes 0 exactly as the reference passes a literal 0 (glistcompare.c:875,880,891,896).
GT4_HD bool eval_stream (const SetOpParams &p, int s, uint32_t c1, uint32_t c2, bool in_a, bool in_b, uint32_t &f)
{
  const uint32_t f1 = in_a ? c1 : 0u;
  const uint32_t f2 = in_b ? c2 : 0u;
  const uint32_t c = p.cutoff;
  if (p.sem == SEM_PAIR) {
    switch (s) {
    case 0:   // include_in_union, :459-466
      if (f1 < c && f2 < c) return false;
      f = calc_freq (f1, f2, p.rule[0], p.count_override);
      return f != 0u;
    case 1:   // include_in_intersection, :468-475 (only reached for keys in both lists, :852)
      if (!(in_a && in_b)) return false;
      if (f1 < c || f2 < c) return false;
      f = calc_freq (f1, f2, p.rule[1], p.count_override);
      return f != 0u;
    case 2:   // include_in_complement (list 1 minus list 2), :477-489
      if (!in_a) return false;
      if (p.subtract) {
        if (f1 != f2 || f1 < c) return false;
        f = f1;
        return true;
      }
      if (f1 < c || f2 >= c) return false;
      f = calc_freq (f1, f2, p.rule[2], p.count_override);
      return f != 0u;
    default:  // diff2: arguments swapped, subtract forced to 0 (:862,:896)
      if (!in_b) return false;
      if (f2 < c || f1 >= c) return false;
      f = calc_freq (f2, f1, p.rule[3], p.count_override);
      return f != 0u;
    }
  }
  if (s != 0) return false;
  const int rule = p.rule[0];
  if (p.sem == SEM_NUNION_PARTIAL || p.sem == SEM_NUNION_FINAL) {
    // union_multi, :549-557: fold of the counts of the lists holding the key (0 is neutral for
    // add and max, so the absent side needs no special case); emit iff combined >= cutoff (:574)
    if (rule == RULE_ADD) f = f1 + f2;
    else if (rule == RULE_MAX) f = (f1 > f2) ? f1 : f2;
    else f = p.count_override;
    return p.sem == SEM_NUNION_PARTIAL ? true : f >= c;
  }
  // intersect_multi, :668-677: left fold over the lists in order; f1 is the fold so far
  if (!(in_a && in_b)) return false;
  if (rule == RULE_MIN) f = (f1 == 0u || f2 < f1) ? f2 : f1;     // the "!freq ||" guard of :669
  else if (rule == RULE_MAX) f = (f2 > f1) ? f2 : f1;
  else if (rule == RULE_ADD) f = f1 + f2;
  else f = p.count_override;
  return p.sem == SEM_NISECT_PARTIAL ? true : f >= c;
}

// Compile-time specialisations of eval_stream for the single-output configurations that matter
// for throughput (the defaults of glistcompare and the nodes of the N-list tree / chain).  Each
// must agree with eval_stream for the parameters it is selected for (select_fast_path);
// tests/test_core_emulation.py checks that exhaustively on random inputs.
enum FastPath : int {
  FAST_GENERIC = 0,   // runtime eval_stream
  FAST_U_ADD = 1,     // SEM_PAIR, union, rule add
  FAST_I_MIN = 2,     // SEM_PAIR, intersection, rule min
  FAST_D_SUB = 3,     // SEM_PAIR, diff1, rule subtract, no -du
  FAST_NU_ADD = 4,    // N-list union node, rule add
  FAST_NI_MIN = 5,    // N-list intersection link, rule min
  FAST_D2_SUB = 6     // SEM_PAIR, diff2 (list 2 minus list 1), rule subtract
};

GT4_HD int select_fast_path (const SetOpParams &p, int stream)
{
  if (p.sem == SEM_PAIR) {
    if (stream == 0 && p.rule[0] == RULE_ADD) return FAST_U_ADD;
    if (stream == 1 && p.rule[1] == RULE_MIN) return FAST_I_MIN;
    if (stream == 2 && p.rule[2] == RULE_SUBTRACT && !p.subtract) return FAST_D_SUB;
    if (stream == 3 && p.rule[3] == RULE_SUBTRACT) return FAST_D2_SUB;
    return FAST_GENERIC;
  }
  if (stream != 0) return FAST_GENERIC;
  if ((p.sem == SEM_NUNION_PARTIAL || p.sem == SEM_NUNION_FINAL) && p.rule[0] == RULE_ADD) return FAST_NU_ADD;
  if ((p.sem == SEM_NISECT_PARTIAL || p.sem == SEM_NISECT_FINAL) && p.rule[0] == RULE_MIN) return FAST_NI_MIN;
  return FAST_GENERIC;
}

template <int FAST>
GT4_HD bool eval_fast (const SetOpParams &p, int stream, uint32_t c1, uint32_t c2, bool in_a, bool in_b, uint32_t &f)
{
  const uint32_t c = p.cutoff;
  if (FAST == FAST_U_ADD) {
    const uint32_t f1 = in_a ? c1 : 0u, f2 = in_b ? c2 : 0u;
    f = f1 + f2;
    return (f1 >= c || f2 >= c) && f != 0u;
  } else if (FAST == FAST_I_MIN) {
    f = (c1 < c2) ? c1 : c2;
    return in_a && in_b && f >= c && f != 0u;        // min (c1, c2) >= c  <=>  both >= c
  } else if (FAST == FAST_D_SUB) {
    const uint32_t f2 = in_b ? c2 : 0u;
    f = (c1 > f2) ? c1 - f2 : 0u;
    return in_a && c1 >= c && f2 < c && f != 0u;
  } else if (FAST == FAST_D2_SUB) {
    const uint32_t f1 = in_a ? c1 : 0u;
    f = (c2 > f1) ? c2 - f1 : 0u;
    return in_b && c2 >= c && f1 < c && f != 0u;
  } else if (FAST == FAST_NU_ADD) {
    f = (in_a ? c1 : 0u) + (in_b ? c2 : 0u);
    return p.sem == SEM_NUNION_PARTIAL || f >= c;
  } else if (FAST == FAST_NI_MIN) {
    f = (c1 == 0u || c2 < c1) ? c2 : c1;
    return in_a && in_b && (p.sem == SEM_NISECT_PARTIAL || f >= c);
  } else {
    return eval_stream (p, stream, c1, c2, in_a, in_b, f);
  }
}

// Merge-path co-rank: how many elements of A are among the first `diag` elements of the merge
// of A and B when ties take A first.  Works on any random-access key arrays (global or shared).
template <typename Index>
GT4_HD Index merge_path (const uint64_t *a, Index na, const uint64_t *b, Index nb, Index diag)
{
  Index lo = diag > nb ? diag - nb : 0;
  Index hi = diag < na ? diag : na;
  while (lo < hi) {
    const Index mid = lo + ((hi - lo) >> 1);
    if (a[mid] <= b[diag - 1 - mid]) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// Same search restricted to a window known from coarser diagonals: the co-rank is monotone in the
// diagonal, so for D_lo <= diag <= D_hi it lies between the co-ranks of D_lo and D_hi.
template <typename Index>
GT4_HD Index merge_path_window (const uint64_t *a, Index na, const uint64_t *b, Index nb, Index diag, Index lo_hint, Index hi_hint)
{
  Index lo = diag > nb ? diag - nb : 0;
  Index hi = diag < na ? diag : na;
  if (lo < lo_hint) lo = lo_hint;
  if (hi > hi_hint) hi = hi_hint;
  while (lo < hi) {
    const Index mid = lo + ((hi - lo) >> 1);
    if (a[mid] <= b[diag - 1 - mid]) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// One thread's share of a tile: VT consecutive slots of the merged sequence starting at slot
// `d0`, A-cursor `i0` (= merge_path (..., d0)).  ka/ca and kb/cb are the tile's slices; ka[-1] is
// readable iff has_halo (the element of A just before the tile) and kb[nb] iff has_peek (the
// element of B just after it).  Every distinct key is reported exactly once:
//   - an A slot always reports, pairing with the next unconsumed B element when equal;
//   - the B slot that follows a paired A slot is the second half of that pair and is dead (it may
//     be the first slot of the next thread or tile: then the element of A just before the cursor
//     carries the same key).
// sink (slot, key, c1, c2, in_a, in_b, live) is called for all VT slots; c1 is meaningful iff in_a,
// c2 iff in_b; live == false marks slots past the end of the tile or the second half of a pair.
// Cursors are kept as pointers so that the per-slot work is a handful of compares, two or three
// shared-memory loads and predicated increments.
template <int VT, typename Sink>
GT4_HD void merge_slots (const uint64_t *ka, const uint32_t *ca, int na, bool has_halo,
                         const uint64_t *kb, const uint32_t *cb, int nb, bool has_peek,
                         int i0, int d0, Sink &&sink)
{
  const int n_live = na + nb - d0;                 // slots of this thread inside the tile (may exceed VT)
  const uint64_t *pa = ka + i0, *pb = kb + (d0 - i0);
  const uint32_t *pca = ca + i0, *pcb = cb + (d0 - i0);
  const uint64_t *const a_end = ka + na, *const b_end = kb + nb, *const b_ext = b_end + (has_peek ? 1 : 0);
  uint64_t key_a = *pa;       // may be past the slice: the buffers carry slack, value unused then
  uint64_t key_b = *pb;
  bool skip = ((i0 > 0) || has_halo) && (pb < b_ext) && (pa[-1] == key_b);
GT4_UNROLL
  for (int s = 0; s < VT; s++) {
    const bool a_avail = pa < a_end;
    const bool b_avail = pb < b_end;
    const bool take_a = a_avail && (!b_avail || key_a <= key_b);
    const bool pair = take_a && (pb < b_ext) && (key_b == key_a);
    const bool live = (s < n_live) && !(skip && !take_a);
    // one count load for the slot's own element, a second one only for the B half of a pair
    const uint32_t cnt_self = *(take_a ? pca : pcb);
    const uint32_t cnt_pair = pair ? *pcb : 0u;
    sink (s, take_a ? key_a : key_b, cnt_self, take_a ? cnt_pair : cnt_self, take_a, !take_a || pair, live);
    skip = pair;
    if (take_a) {
      pa += 1;
      pca += 1;
      key_a = *pa;
    } else {
      pb += 1;
      pcb += 1;
      key_b = *pb;
    }
  }
}

// The same for a thread whose VT slots all lie inside the tile, when the tile is staged with the element of A before
// it (ka[-1]), the element of A after it (ka[na]) and the element of B after it (kb[nb]).  Those neighbours are natural
// sentinels: tile boundaries are merge-path co-ranks with ties taking A first, so A[a_hi] > every B of the tile and
// B[b_hi] >= every A of the tile, and no cursor needs a bound.  A slot is reported one step late, when the following
// slot has been looked at: the B half of a pair is loaded anyway (it is the next slot) and brings the pair's second
// count, so only a pair in the thread's LAST slot costs a load of its own.
template <int VT, typename Sink>
GT4_HD void merge_slots_interior (const uint64_t *ka, const uint32_t *ca, const uint64_t *kb, const uint32_t *cb,
                                  int i0, int d0, Sink &&sink)
{
  const uint64_t *pa = ka + i0, *pb = kb + (d0 - i0);
  const uint32_t *pca = ca + i0, *pcb = cb + (d0 - i0);
  uint64_t key_a = *pa;
  uint64_t key_b = *pb;
  bool live = pa[-1] != key_b;      // else the first slot is the B half of a pair reported by the previous thread (or tile)
  uint64_t key_prev = 0;
  uint32_t cnt_prev = 0;
  bool a_prev = false, pair_prev = false, live_prev = false;
GT4_UNROLL
  for (int s = 0; s < VT; s++) {
    const bool take_a = key_a <= key_b;
    const bool pair = key_a == key_b;
    const uint32_t cnt = *(take_a ? pca : pcb);
    if (s > 0) sink (s - 1, key_prev, cnt_prev, a_prev ? (pair_prev ? cnt : 0u) : cnt_prev, a_prev, !a_prev || pair_prev, live_prev);
    key_prev = take_a ? key_a : key_b;
    cnt_prev = cnt;
    a_prev = take_a;
    pair_prev = pair;
    live_prev = live;
    live = !pair;                   // the slot after a pair is its B half
    if (take_a) {
      pa += 1;
      pca += 1;
      key_a = *pa;
    } else {
      pb += 1;
      pcb += 1;
      key_b = *pb;
    }
  }
  const uint32_t cnt_last = pair_prev ? *pcb : 0u;
  sink (VT - 1, key_prev, cnt_prev, a_prev ? cnt_last : cnt_prev, a_prev, !a_prev || pair_prev, live_prev);
}

}  // namespace gt4gpu
