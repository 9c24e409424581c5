"""GPU parity tests of the single-pass N-list kernel (csrc/gt4gpu_kway_kernel.cu) behind gt4gpu_union_multi /
gt4gpu_intersect_multi / gt4gpu_write_union: against the oracle (union_multi, intersect_multi of
/root/reference/src/glistcompare.c:500-717) and against the tree / chain of two-list merges it replaces."""
from __future__ import annotations

import numpy as np
import pytest

from tests.util import make_counts, make_multi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    import genometester4_b200 as g
    g.init(0)
    g.set_option("use_kway", 2)        # unions AND intersections through the single-pass kernel (the default keeps the chain for intersections)
    yield g
    g.set_option("use_kway", 1)


def _check_all(g, oracle, lists, k, rules_u=("default", "max", "add"), rules_i=("default", "max", "add"), cutoffs=(1, 40), tag=""):
    gl = [g.WordList.from_arrays(w, c, k) for w, c in lists]
    ol = [oracle.SList(w, c, k) for w, c in lists]
    for cutoff in cutoffs:
        for rule in rules_u:
            kw = dict(rule="number", count_override=7) if rule == "number" else dict(rule=rule)
            rc, want = oracle.union_multi(ol, cutoff=cutoff, **kw)
            got = g.union_multi(gl, cutoff=cutoff, **kw)
            w, c = got.to_host()
            assert np.array_equal(w, want.words) and np.array_equal(c, want.counts), (tag, "union", rule, cutoff)
            assert (got.n_words, got.total_count) == (want.n_words, want.total_count), (tag, "union", rule, cutoff)
            co = g.union_multi(gl, cutoff=cutoff, countonly=1, **kw)
            assert (co.n_words, co.total_count) == (want.n_words, want.total_count), (tag, "union count-only", rule, cutoff)
        for rule in rules_i:
            kw = dict(rule="number", count_override=7) if rule == "number" else dict(rule=rule)
            rc, want = oracle.intersect_multi(ol, cutoff=cutoff, **kw)
            got = g.intersect_multi(gl, cutoff=cutoff, **kw)
            w, c = got.to_host()
            assert np.array_equal(w, want.words) and np.array_equal(c, want.counts), (tag, "isect", rule, cutoff)
            assert (got.n_words, got.total_count) == (want.n_words, want.total_count), (tag, "isect", rule, cutoff)
            co = g.intersect_multi(gl, cutoff=cutoff, countonly=1, **kw)
            assert (co.n_words, co.total_count) == (want.n_words, want.total_count), (tag, "isect count-only", rule, cutoff)


@pytest.mark.parametrize("n_lists", [3, 4, 5, 8, 9, 17, 64])
def test_kway_vs_oracle_every_list_count(g, oracle, n_lists):
    """3..8 lists: one pass; 9, 17, 64 lists: groups of up to 8 (union) / a chain with the running result first
    (intersection).  Hundreds of tiles for the small counts."""
    n_each = 40_000 if n_lists <= 9 else 6_000
    lists = make_multi(900 + n_lists, n_lists, n_each, int(n_each * 1.6), 32, "tail")
    rules = ("default", "max", "add", "number", "min")
    _check_all(g, oracle, lists, 32, rules_u=("default", "max", "number"), rules_i=rules, cutoffs=(0, 1, 40), tag=n_lists)


def test_kway_equals_the_tree_it_replaces(g):
    lists = make_multi(41, 6, 90_000, 200_000, 25, "tail")
    gl = [g.WordList.from_arrays(w, c, 25) for w, c in lists]
    try:
        g.set_option("use_kway", 0)
        tree_u = g.union_multi(gl, cutoff=3).to_host()
        tree_i = g.intersect_multi(gl, cutoff=1).to_host()
    finally:
        g.set_option("use_kway", 2)
    kway_u = g.union_multi(gl, cutoff=3).to_host()
    kway_i = g.intersect_multi(gl, cutoff=1).to_host()
    for a, b in ((tree_u, kway_u), (tree_i, kway_i)):
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_kway_skewed_and_degenerate_inputs(g, oracle):
    rng = np.random.default_rng(5)
    # (a) N copies of the same list: every word in every list, every boundary key a tie between all lists
    w = np.unique(rng.integers(0, 1 << 50, size=30_000, dtype=np.uint64))
    same = [(w.copy(), make_counts(rng, w.size, "tail")) for _ in range(5)]
    _check_all(g, oracle, same, 25, tag="identical")
    # (b) dense keys 0..n-1 (every bucket boundary falls between neighbours) against sparse ones
    dense = np.arange(50_000, dtype=np.uint64)
    lists = [(dense, make_counts(rng, dense.size, "small")),
             (dense[::3].copy(), make_counts(rng, dense[::3].size, "small")),
             (np.unique(rng.integers(0, 1 << 40, size=20_000, dtype=np.uint64)), None)]
    lists[2] = (lists[2][0], make_counts(rng, lists[2][0].size, "tail"))
    _check_all(g, oracle, lists, 20, tag="dense")
    # (c) one tight cluster plus far outliers: nearly every word of a tile lands in one bucket
    cluster = (np.uint64(1) << np.uint64(60)) + np.arange(0, 9_000, dtype=np.uint64) * np.uint64(3)
    outl = np.array([5, 1 << 20, (1 << 64) - 2, (1 << 64) - 1], dtype=np.uint64)
    la = np.unique(np.concatenate([cluster, outl]))
    lb = np.unique(np.concatenate([cluster[::2] + np.uint64(1), outl[1:]]))
    lc = np.unique(np.concatenate([cluster[::5], outl[:2]]))
    lists = [(x, make_counts(rng, x.size, "tail")) for x in (la, lb, lc)]
    _check_all(g, oracle, lists, 32, tag="cluster")
    # (d) very different lengths, lists shorter than the sampling distance, counts that wrap
    big = np.unique(rng.integers(0, 1 << 64, size=120_000, dtype=np.uint64, endpoint=False))
    lists = [(big, make_counts(rng, big.size, "huge")), (big[:50].copy(), make_counts(rng, 50, "huge")),
             (big[-3:].copy(), make_counts(rng, 3, "huge")), (big[1000:1001].copy(), make_counts(rng, 1, "huge"))]
    _check_all(g, oracle, lists, 32, rules_u=("default", "max"), rules_i=("default", "add", "min"), cutoffs=(0, 1, 5), tag="lengths")
    # (e) zero counts: rule min's "!freq ||" guard is order dependent (glistcompare.c:669)
    w = np.arange(1000, 9000, dtype=np.uint64)
    lists = []
    for j in range(4):
        c = rng.integers(0, 3, size=w.size).astype(np.uint32)
        lists.append((w.copy(), c))
    _check_all(g, oracle, lists, 16, rules_i=("default", "min", "max", "add"), cutoffs=(0, 1, 2), tag="zero counts")


def test_kway_caller_buffers_capacity_and_fallbacks(g, oracle):
    import torch
    lists = make_multi(77, 4, 30_000, 70_000, 22, "tail")
    gl = [g.WordList.from_arrays(w, c, 22) for w, c in lists]
    ol = [oracle.SList(w, c, 22) for w, c in lists]
    rc, want = oracle.union_multi(ol, cutoff=2)
    # misaligned device views cannot take the TMA slices of the k-way pass: the call falls back to the tree
    keep, views = [], []
    for w, c in lists:
        tw = torch.zeros(w.size + 1, dtype=torch.int64, device="cuda")
        tc = torch.zeros(c.size + 1, dtype=torch.int32, device="cuda")
        tw[1:] = torch.from_numpy(w.view(np.int64)).cuda()
        tc[1:] = torch.from_numpy(c.view(np.int32)).cuda()
        keep.append((tw, tc))
        views.append(g.WordList.from_device(tw.data_ptr() + 8, tc.data_ptr() + 4, w.size, 22, keepalive=(tw, tc)))
    torch.cuda.synchronize()
    got = g.union_multi(views, cutoff=2)
    w, c = got.to_host()
    assert np.array_equal(w, want.words) and np.array_equal(c, want.counts)
    # unsorted input: an error, never a crash or a hang
    bad_w = lists[0][0].copy()
    bad_w[1000:1010] = bad_w[1000:1010][::-1]
    bad = [g.WordList.from_arrays(bad_w, lists[0][1], 22)] + gl[1:]
    try:
        g.union_multi(bad, cutoff=1)
    except g.GT4GPUError as e:
        assert e.code != 0
    again = g.union_multi(gl, cutoff=2)             # the library still works afterwards
    w, c = again.to_host()
    assert np.array_equal(w, want.words) and np.array_equal(c, want.counts)


def test_kway_write_union_file(g, oracle, tmp_path):
    import os

    from tests import refrun
    lists = make_multi(78, 7, 20_000, 60_000, 20, "tail")
    gl = [g.WordList.from_arrays(w, c, 20) for w, c in lists]
    ol = [oracle.SList(w, c, 20) for w, c in lists]
    rc, want = oracle.write_union(ol, cutoff=2)
    p = tmp_path / "wu.list"
    fd = os.open(p, os.O_RDWR | os.O_CREAT, 0o644)
    h = g.gt4_write_union(gl, 2, fd)
    os.close(fd)
    assert (h.n_words, h.total_count, h.word_length) == (want.n_words, want.total_count, 20)
    assert p.read_bytes() == refrun.list_bytes(want, 20)


def test_kway_large_shared_universe_properties(g):
    """Config-5 shaped inputs (8 lists drawn from ONE universe) at a size the oracle would take minutes for:
    checked through size-independent properties and against the tree of two-list merges."""
    import torch

    from genometester4_b200 import synth
    n_each = 4_000_000
    m = n_each * 3
    lists, keep = [], []
    for j in range(8):
        w, c = synth.list_torch(5, 32, m, 0, m, j, 1 / 3)
        keep.append((w, c))
        lists.append(g.WordList.from_device(w.data_ptr(), c.data_ptr(), w.numel(), 32, keepalive=(w, c)))
    torch.cuda.synchronize()
    n_in = sum(len(x) for x in lists)
    sum_in = sum(int(c.to(torch.int64).sum().item()) for _, c in keep)
    u = g.union_multi(lists, cutoff=0)
    allk = torch.unique(torch.cat([w for w, _ in keep]))
    uw = u.as_torch()[0]
    assert u.n_words == allk.numel() and u.total_count == sum_in
    assert torch.equal(torch.sort(uw.view(torch.int64)).values, allk)          # (torch orders the bit patterns as signed)
    co = g.union_multi(lists, cutoff=0, countonly=1)
    assert (co.n_words, co.total_count) == (u.n_words, u.total_count)
    i = g.intersect_multi(lists, cutoff=0)
    assert 0 < i.n_words < min(len(x) for x in lists)
    try:
        g.set_option("use_kway", 0)
        ut = g.union_multi(lists, cutoff=5)
        it = g.intersect_multi(lists, cutoff=0)
    finally:
        g.set_option("use_kway", 2)
    uk = g.union_multi(lists, cutoff=5)
    for a, b in ((ut, uk), (it, i)):
        assert (a.n_words, a.total_count) == (b.n_words, b.total_count)
        aw, ac = a.as_torch()
        bw, bc = b.as_torch()
        assert torch.equal(aw, bw) and torch.equal(ac, bc)
    assert n_in > u.n_words
