"""CPU check of the kernel's arithmetic core: the tile / thread decomposition (merge-path
partition, halo + peek, per-thread serial merge, rule + cut-off predicates) is replayed on the
host with the same __host__ __device__ functions the CUDA kernel calls, and compared with the
oracle over the whole flag matrix and at many tile shapes (so that equal pairs straddle thread
and tile boundaries in every possible way)."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from tests import cases

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "tests" / "emulate_tile.cpp"
SO = ROOT / "tests" / "_build" / "libemulate_tile.so"

SEM_PAIR, SEM_NU_PART, SEM_NU_FINAL, SEM_NI_PART, SEM_NI_FINAL = range(5)


@pytest.fixture(scope="module")
def emu():
    SO.parent.mkdir(exist_ok=True)
    subprocess.run(["g++", "-O1", "-std=c++17", "-Wall", "-Wno-unknown-pragmas", "-fPIC", "-shared",
                    f"-I{ROOT / 'genometester4_b200' / 'csrc'}", "-o", str(SO), str(SRC)], check=True)
    return C.CDLL(str(SO))


def run_emu(emu, a, b, nt, vt, mask, sem, rule, cutoff, ov, subtract):
    aw, ac = np.ascontiguousarray(a[0], np.uint64), np.ascontiguousarray(a[1], np.uint32)
    bw, bc = np.ascontiguousarray(b[0], np.uint64), np.ascontiguousarray(b[1], np.uint32)
    cap = len(aw) + len(bw) + 1
    ow = [np.zeros(cap, np.uint64) for _ in range(4)]
    oc = [np.zeros(cap, np.uint32) for _ in range(4)]
    pw = (C.c_void_p * 4)(*[x.ctypes.data for x in ow])
    pc = (C.c_void_p * 4)(*[x.ctypes.data for x in oc])
    n_out = (C.c_uint64 * 4)()
    s_out = (C.c_uint64 * 4)()
    rc = emu.emu_setop2(C.c_void_p(aw.ctypes.data), C.c_void_p(ac.ctypes.data), C.c_uint64(len(aw)),
                        C.c_void_p(bw.ctypes.data), C.c_void_p(bc.ctypes.data), C.c_uint64(len(bw)),
                        nt, vt, C.c_uint32(mask), sem, rule, C.c_uint32(cutoff), C.c_uint32(ov), int(subtract),
                        pw, pc, n_out, s_out)
    assert rc == 0
    return [(ow[q][:n_out[q]], oc[q][:n_out[q]], n_out[q], s_out[q]) for q in range(4)]


SHAPES = [(1, 1), (2, 1), (3, 3), (4, 7), (2, 9), (32, 7), (5, 3)]


def test_pair_matrix_against_oracle(emu, oracle):
    inputs = cases.pair_inputs()
    n = 0
    for name, ops, rule, cutoff in cases.pair_cases(full=True):
        k, a, b = inputs[name]
        okw = cases.ops_to_kwargs(ops)
        rkw = cases.rule_to_kwargs(rule)
        want = oracle.compare2(oracle.SList(*a, k), oracle.SList(*b, k), cutoff=cutoff, **okw, **rkw)
        mask = (1 if okw["union"] else 0) | (2 if okw["intrsec"] else 0) | (4 if okw["diff"] else 0) | (8 if okw["ddiff"] else 0)
        shapes = SHAPES if n % 7 == 0 else [SHAPES[n % len(SHAPES)]]
        for nt, vt in shapes:
            got = run_emu(emu, a, b, nt, vt, mask, SEM_PAIR, oracle.RULES[rkw["rule"]], cutoff, rkw["count_override"], okw["subtract"])
            for q, key in enumerate(["union", "intrsec", "diff1", "diff2"]):
                if key not in want:
                    continue
                w = want[key]
                assert got[q][2] == w.n_words and got[q][3] == w.total_count, (name, ops, rule, cutoff, nt, vt, key)
                assert np.array_equal(got[q][0], w.words) and np.array_equal(got[q][1], w.counts), (name, ops, rule, cutoff, nt, vt, key)
        n += 1
    assert n > 1000


def test_every_tile_shape_on_boundary_heavy_input(emu, oracle):
    # identical lists: every key is a pair, so every thread/tile boundary splits a pair half the time
    rng = np.random.default_rng(7)
    w = np.unique(rng.integers(0, 1 << 40, size=2000, dtype=np.uint64))
    ca = rng.integers(1, 9, size=w.size, dtype=np.uint32)
    cb = rng.integers(1, 9, size=w.size, dtype=np.uint32)
    for drop in (0, 1, 2, 5):
        a = (w, ca)
        b = (w[drop:], cb[drop:])
        want = oracle.compare2(oracle.SList(*a, 20), oracle.SList(*b, 20), union=True, intrsec=True, diff=True, ddiff=True, cutoff=3)
        for nt in (1, 2, 3, 4, 5, 8, 32):
            for vt in (1, 3, 7, 9):
                got = run_emu(emu, a, b, nt, vt, 15, SEM_PAIR, 0, 3, 1, False)
                for q, key in enumerate(["union", "intrsec", "diff1", "diff2"]):
                    assert np.array_equal(got[q][0], want[key].words), (drop, nt, vt, key)
                    assert np.array_equal(got[q][1], want[key].counts), (drop, nt, vt, key)


def _tree_union(emu, lists, rule, cutoff, ov, nt, vt):
    level = [l for l in lists if len(l[0])]
    while len(level) < 2:
        level.append((np.zeros(0, np.uint64), np.zeros(0, np.uint32)))
    while len(level) > 2:
        nxt = []
        for i in range(0, len(level) - 1, 2):
            r = run_emu(emu, level[i], level[i + 1], nt, vt, 1, SEM_NU_PART, rule, cutoff, ov, False)[0]
            nxt.append((r[0], r[1]))
        if len(level) & 1:
            nxt.append(level[-1])
        level = nxt
    return run_emu(emu, level[0], level[1], nt, vt, 1, SEM_NU_FINAL, rule, cutoff, ov, False)[0]


def _chain_intersect(emu, lists, rule, cutoff, ov, nt, vt):
    acc = lists[0]
    for j in range(1, len(lists)):
        sem = SEM_NI_FINAL if j + 1 == len(lists) else SEM_NI_PART
        r = run_emu(emu, acc, lists[j], nt, vt, 1, sem, rule, cutoff, ov, False)[0]
        acc = (r[0], r[1])
    return r


def test_nlist_tree_and_chain_against_oracle(emu, oracle):
    """The N-list entry points are built from two-list merges (balanced tree for union, left chain
    for intersection); check that composition against union_multi / intersect_multi."""
    inputs = cases.multi_inputs()
    n = 0
    for name, ops, rule, cutoff in cases.multi_cases(full=True):
        k, lists = inputs[name]
        sl = [oracle.SList(w, c, k) for w, c in lists]
        rkw = cases.rule_to_kwargs(rule)
        r = oracle.RULES[rkw["rule"]]
        nt, vt = SHAPES[n % len(SHAPES)]
        if "-u" in ops:
            rc, want = oracle.union_multi(sl, cutoff=cutoff, **rkw)
            if rc == 0:
                got = _tree_union(emu, lists, 1 if r == 0 else r, cutoff, rkw["count_override"], nt, vt)
                assert np.array_equal(got[0], want.words) and np.array_equal(got[1], want.counts), (name, rule, cutoff)
                assert got[3] == want.total_count
        if "-i" in ops:
            rc, want = oracle.intersect_multi(sl, cutoff=cutoff, **rkw)
            if rc == 0 and all(len(w) for w, _ in lists):
                got = _chain_intersect(emu, lists, 3 if r == 0 else r, cutoff, rkw["count_override"], nt, vt)
                assert np.array_equal(got[0], want.words) and np.array_equal(got[1], want.counts), (name, rule, cutoff)
            elif rc == 0:
                assert want.n_words == 0
        n += 1
    assert n > 300


def test_fast_path_predicates_agree_with_generic(emu):
    emu.emu_check_fast_paths.restype = C.c_uint64
    assert emu.emu_check_fast_paths() == 0
