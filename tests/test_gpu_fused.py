"""GPU parity tests of the multi-output merge kernel (csrc/gt4gpu_fused_kernel.cu): compare_wordmaps emits every requested
output from ONE loop over the two lists (/root/reference/src/glistcompare.c:843-905); so does setop2_fused_kernel.
Checked against the oracle and against one pass of the single-output kernel per output."""
from __future__ import annotations

import itertools

import numpy as np
import pytest

from tests.util import make_pair

pytestmark = pytest.mark.gpu
STREAMS = ("union", "intrsec", "diff1", "diff2")


@pytest.fixture(scope="module")
def g():
    import genometester4_b200 as g
    g.init(0)
    return g


# (the union with ONE more output runs as two single-output passes: measured faster; it stays in the list as a control)
COMBOS = [dict(find_union=1, find_intrsec=1), dict(find_union=1, find_diff=1), dict(find_union=1, find_intrsec=1, find_diff=1),
          dict(find_union=1, find_intrsec=1, find_diff=1, find_ddiff=1), dict(find_intrsec=1, find_diff=1), dict(find_diff=1, find_ddiff=1),
          dict(find_intrsec=1, find_ddiff=1), dict(find_union=1, find_ddiff=1)]


def _oracle_kw(kw):
    return dict(union=bool(kw.get("find_union")), intrsec=bool(kw.get("find_intrsec")), diff=bool(kw.get("find_diff") or kw.get("find_ddiff")),
                ddiff=bool(kw.get("find_ddiff")))


def _check(got, want, tag):
    assert sorted(got) == sorted(want), tag
    for s in want:
        w, c = got[s].to_host()
        assert np.array_equal(w, want[s].words) and np.array_equal(c, want[s].counts), (tag, s)
        assert (got[s].n_words, got[s].total_count) == (want[s].n_words, want[s].total_count), (tag, s)


def test_fused_every_combination_rule_and_cutoff(g, oracle):
    """Hundreds of tiles of both variants (union in place + auxiliary rest region; rest region in place), odd sizes,
    every rule (generic evaluation) and cut-offs 0 .. above most counts."""
    for seed, (na, nb, both), kind in ((21, (300_001, 250_003, 120_000), "tail"), (22, (65_537, 65_539, 65_537), "small"),
                                       (23, (3, 200_001, 2), "tail"), (24, (150_000, 7, 0), "huge"), (25, (5, 3, 2), "tail"),
                                       (26, (90_000, 90_000, 0), "tail")):
        a, b = make_pair(seed, na, nb, both, 25, kind)
        la, lb = g.WordList.from_arrays(*a, 25), g.WordList.from_arrays(*b, 25)
        sa, sb = oracle.SList(*a, 25), oracle.SList(*b, 25)
        rules = ("default", "add", "subtract", "min", "max", "first", "second", 3) if seed in (21, 22) else ("default", "max")
        for kw, rule, cutoff in itertools.product(COMBOS, rules, (0, 1, 3, 40)):
            rkw = dict(rule="number", count_override=rule) if isinstance(rule, int) else dict(rule=rule)
            want = oracle.compare2(sa, sb, cutoff=cutoff, **_oracle_kw(kw), **rkw)
            kw2 = dict(kw)
            if kw2.get("find_ddiff"):
                kw2["find_diff"] = 1                       # -dd implies -d (glistcompare.c:334); the Python mirror takes both flags
            got = g.compare_wordmaps(la, lb, cutoff=cutoff, **kw2, **rkw)
            _check(got, want, (seed, kw, rule, cutoff))


def test_fused_equals_one_pass_per_output_and_du_falls_back(g, oracle):
    a, b = make_pair(31, 400_000, 350_000, 200_000, 25, "small")
    la, lb = g.WordList.from_arrays(*a, 25), g.WordList.from_arrays(*b, 25)
    kw = dict(find_union=1, find_intrsec=1, find_diff=1, find_ddiff=1, cutoff=2)
    fused = g.compare_wordmaps(la, lb, **kw)
    p, m, launches_fused = g.last_timing()
    try:
        g.set_option("use_fused", 0)
        single = g.compare_wordmaps(la, lb, **kw)
        _, _, launches_single = g.last_timing()
    finally:
        g.set_option("use_fused", 1)
    assert launches_fused == 2 and launches_single == 5           # partition + one pass  vs  partition + four passes
    for s in STREAMS:
        fw, fc = fused[s].to_host()
        sw, sc = single[s].to_host()
        assert np.array_equal(fw, sw) and np.array_equal(fc, sc), s
    # -du: diff1 keeps words with EQUAL counts, which the intersection holds too -> one pass per output
    want = oracle.compare2(oracle.SList(*a, 25), oracle.SList(*b, 25), intrsec=True, diff=True, subtract=True, cutoff=1)
    got = g.compare_wordmaps(la, lb, find_intrsec=1, find_diff=1, subtract=1, cutoff=1)
    _check(got, want, "du")
    assert g.last_timing()[2] == 3


def test_fused_caller_buffers_and_capacity(g, oracle):
    import torch
    a, b = make_pair(32, 120_000, 100_000, 40_000, 20, "tail")
    la, lb = g.WordList.from_arrays(*a, 20), g.WordList.from_arrays(*b, 20)
    want = oracle.compare2(oracle.SList(*a, 20), oracle.SList(*b, 20), union=True, intrsec=True, diff=True, cutoff=1)
    bufs, keep = {}, []
    for s, cap in (("union", len(la) + len(lb)), ("intrsec", min(len(la), len(lb))), ("diff1", len(la))):
        w = torch.empty(cap, dtype=torch.int64, device="cuda")
        c = torch.empty(cap, dtype=torch.int32, device="cuda")
        keep.append((w, c))
        bufs[s] = (w.data_ptr(), c.data_ptr(), cap)
    got = g.compare_wordmaps(la, lb, find_union=1, find_intrsec=1, find_diff=1, cutoff=1, out_buffers=bufs)
    _check(got, want, "caller buffers")
    bufs["intrsec"] = (bufs["intrsec"][0], bufs["intrsec"][1], want["intrsec"].n_words - 1)
    with pytest.raises(g.GT4GPUError):
        g.compare_wordmaps(la, lb, find_union=1, find_intrsec=1, find_diff=1, cutoff=1, out_buffers=bufs)


def test_fused_large_properties(g):
    """2e7 + 2e7 device-resident 25-mers: |A u B| = |A| + |B| - |A n B|, diff1 = A minus B, sums conserved, and every
    output identical to the single-output kernel's."""
    import torch

    from genometester4_b200 import synth
    m = 30_000_000
    (wa, ca), (wb, cb) = synth.pair_torch(9, 25, m, 0, m, 1 / 3, 1 / 3)
    la = g.WordList.from_device(wa.data_ptr(), ca.data_ptr(), wa.numel(), 25, keepalive=(wa, ca))
    lb = g.WordList.from_device(wb.data_ptr(), cb.data_ptr(), wb.numel(), 25, keepalive=(wb, cb))
    kw = dict(find_union=1, find_intrsec=1, find_diff=1, find_ddiff=1, cutoff=1)
    f = g.compare_wordmaps(la, lb, **kw)
    assert f["union"].n_words == len(la) + len(lb) - f["intrsec"].n_words
    assert f["diff1"].n_words == len(la) - f["intrsec"].n_words and f["diff2"].n_words == len(lb) - f["intrsec"].n_words
    assert f["union"].total_count == int(ca.to(torch.int64).sum().item()) + int(cb.to(torch.int64).sum().item())
    try:
        g.set_option("use_fused", 0)
        s1 = g.compare_wordmaps(la, lb, **kw)
    finally:
        g.set_option("use_fused", 1)
    for s in STREAMS:
        fw, fc = f[s].as_torch()
        sw, sc = s1[s].as_torch()
        assert torch.equal(fw, sw) and torch.equal(fc, sc), s
