"""GPU parity tests proper: every call goes through the C ABI of libgt4gpu.so (ctypes) and is
compared bit-for-bit with the oracle and with the committed golden digests of the unmodified
reference binary."""
import ctypes as C
import json
import os
from pathlib import Path

import numpy as np
import pytest

from tests import cases, refrun
from tests.util import make_multi, make_pair

pytestmark = pytest.mark.gpu

GOLDEN = json.loads((Path(__file__).parent / "golden" / "golden.json").read_text())
STREAMS = ("union", "intrsec", "diff1", "diff2")
FILES = {"union": "union", "intrsec": "intrsec", "diff1": "0_diff1", "diff2": "0_diff2"}


@pytest.fixture(scope="module")
def g():
    import genometester4_b200 as g
    g.init(0)
    yield g
    g.set_tile(256, 9)


@pytest.fixture(scope="module")
def pair_lists(g):
    out = {}
    for name, (k, a, b) in cases.pair_inputs().items():
        out[name] = (k, g.WordList.from_arrays(a[0], a[1], k), g.WordList.from_arrays(b[0], b[1], k))
    return out


@pytest.fixture(scope="module")
def multi_lists(g):
    out = {}
    for name, (k, lists) in cases.multi_inputs().items():
        out[name] = (k, [g.WordList.from_arrays(w, c, k) for w, c in lists])
    return out


def gpu_pair(g, la, lb, ops, rule, cutoff, countonly=0):
    okw = cases.ops_to_kwargs(ops)
    rkw = cases.rule_to_kwargs(rule)
    return g.compare_wordmaps(la, lb, find_union=okw["union"], find_intrsec=okw["intrsec"], find_diff=okw["diff"],
                              find_ddiff=okw["ddiff"], subtract=okw["subtract"], countonly=countonly, cutoff=cutoff,
                              rule=rkw["rule"], count_override=rkw["count_override"])


def test_pair_golden_matrix(g, pair_lists):
    """552 flag combinations x 13 input sets: byte-identical list files vs the reference binary's digests."""
    for i, gold in enumerate(GOLDEN["pair"]):
        k, la, lb = pair_lists[gold["input"]]
        if i % 50 == 0:
            g.set_tile(*[(256, 9), (128, 15), (256, 7), (512, 7), (256, 15)][(i // 50) % 5])
        res = gpu_pair(g, la, lb, tuple(gold["ops"]), gold["rule"], gold["cutoff"])
        assert sorted(f"out_{k}_{FILES[s]}.list" for s in res) == sorted(gold["files"]), gold
        for s, r in res.items():
            ref = gold["files"][f"out_{k}_{FILES[s]}.list"]
            assert (r.n_words, r.total_count) == (ref["n_words"], ref["total_count"]), (gold, s)
            assert refrun.digest(r.list_bytes()) == ref["sha256"], (gold, s)
        co = gpu_pair(g, la, lb, tuple(gold["ops"]), gold["rule"], gold["cutoff"], countonly=1)
        stdout = "".join(f"NUnique\t{co[s].n_words}\nNTotal\t{co[s].total_count}\n" for s in STREAMS if s in co)
        assert stdout == gold["count_only_stdout"], gold
    g.set_tile(256, 9)


def test_pair_full_matrix_vs_oracle(g, pair_lists, oracle):
    inputs = cases.pair_inputs()
    n = 0
    for name, ops, rule, cutoff in cases.pair_cases(full=True):
        k, a, b = inputs[name]
        want = oracle.compare2(oracle.SList(*a, k), oracle.SList(*b, k), cutoff=cutoff,
                               **cases.ops_to_kwargs(ops), **cases.rule_to_kwargs(rule))
        got = gpu_pair(g, pair_lists[name][1], pair_lists[name][2], ops, rule, cutoff)
        assert set(got) == set(want)
        for s in want:
            w, c = got[s].to_host()
            assert np.array_equal(w, want[s].words) and np.array_equal(c, want[s].counts), (name, ops, rule, cutoff, s)
            assert got[s].total_count == want[s].total_count
        n += 1
    assert n > 1000


@pytest.mark.parametrize("shape", [(128, 15), (128, 17), (256, 7), (256, 9), (256, 11), (256, 15), (512, 7), (512, 9)])
def test_medium_lists_every_tile_shape(g, oracle, shape):
    """The one-CTA-per-tile kernel (fused multi-output; reached with use_stream_kernel=0): hundreds of tiles,
    partition, look-back offsets and compaction across CTAs."""
    g.set_tile(*shape)
    g.set_option("use_stream_kernel", 0)
    for seed, (na, nb, both), kind in ((1, (300_000, 250_000, 120_000), "tail"), (2, (200_000, 200_000, 200_000), "small"),
                                       (3, (1, 400_000, 1), "tail"), (4, (150_000, 10, 0), "huge")):
        a, b = make_pair(seed, na, nb, both, 25, kind)
        la, lb = g.WordList.from_arrays(*a, 25), g.WordList.from_arrays(*b, 25)
        want = oracle.compare2(oracle.SList(*a, 25), oracle.SList(*b, 25), union=True, intrsec=True, diff=True, ddiff=True, cutoff=3)
        got = g.compare_wordmaps(la, lb, 1, 1, 1, 1, cutoff=3)
        for s in STREAMS:
            w, c = got[s].to_host()
            assert np.array_equal(w, want[s].words) and np.array_equal(c, want[s].counts), (shape, seed, s)
            assert (got[s].n_words, got[s].total_count) == (want[s].n_words, want[s].total_count)
        for s, kw in (("union", dict(find_union=1)), ("intrsec", dict(find_intrsec=1)), ("diff1", dict(find_diff=1))):
            single = g.compare_wordmaps(la, lb, cutoff=3, **kw)[s]      # the 1-stream kernel variant
            w, c = single.to_host()
            assert np.array_equal(w, want[s].words) and np.array_equal(c, want[s].counts), (shape, seed, s, "single")
            co = g.compare_wordmaps(la, lb, cutoff=3, countonly=1, **kw)[s]
            assert (co.n_words, co.total_count) == (want[s].n_words, want[s].total_count)
    g.set_tile(256, 9)
    g.set_option("use_stream_kernel", 1)


@pytest.mark.parametrize("shape", [(256, 7), (256, 9), (256, 11), (384, 9), (384, 11), (512, 7), (512, 9), (512, 11)])
def test_stream_kernel_every_shape_rule_and_semantics(g, oracle, shape):
    """The persistent single-output kernel (TMA-staged, warp-specialised): every items-per-thread
    variant x every rule (fast paths and the generic path) x pair / N-list semantics, on inputs of
    hundreds of tiles with odd sizes so that slices start at every 16-byte misalignment."""
    items = shape
    g.set_option("stream_shape", shape[0] * 100 + shape[1])
    g.set_option("use_stream_kernel", 1)
    try:
        for seed, (na, nb, both), kind in ((11, (300_001, 250_003, 120_000), "tail"), (12, (65_537, 65_539, 65_537), "small"),
                                           (13, (3, 200_001, 2), "tail"), (14, (150_000, 7, 0), "huge"), (15, (5, 3, 2), "tail")):
            a, b = make_pair(seed, na, nb, both, 25, kind)
            # sub-views that start at odd element offsets: A[1:], B[3:] are 8- and 12-byte misaligned
            la, lb = g.WordList.from_arrays(*a, 25), g.WordList.from_arrays(*b, 25)
            sa, sb = oracle.SList(*a, 25), oracle.SList(*b, 25)
            for rule in ("default", "add", "subtract", "min", "max", "first", "second", 3):
                rkw = dict(rule="number", count_override=rule) if isinstance(rule, int) else dict(rule=rule)
                for cutoff in (1, 3):
                    want = oracle.compare2(sa, sb, union=True, intrsec=True, diff=True, ddiff=True, cutoff=cutoff, **rkw)
                    want_du = oracle.compare2(sa, sb, diff=True, subtract=True, cutoff=cutoff, **rkw)["diff1"]
                    for s, kw in (("union", dict(find_union=1)), ("intrsec", dict(find_intrsec=1)), ("diff1", dict(find_diff=1)),
                                  ("diff2", dict(find_ddiff=1))):
                        r = g.compare_wordmaps(la, lb, cutoff=cutoff, **kw, **rkw)[s]
                        w, c = r.to_host()
                        assert np.array_equal(w, want[s].words) and np.array_equal(c, want[s].counts), (items, seed, rule, cutoff, s)
                        assert (r.n_words, r.total_count) == (want[s].n_words, want[s].total_count)
                        co = g.compare_wordmaps(la, lb, cutoff=cutoff, countonly=1, **kw, **rkw)[s]
                        assert (co.n_words, co.total_count) == (want[s].n_words, want[s].total_count), (items, seed, rule, cutoff, s, "co")
                    r = g.compare_wordmaps(la, lb, cutoff=cutoff, find_diff=1, subtract=1, **rkw)["diff1"]
                    w, c = r.to_host()
                    assert np.array_equal(w, want_du.words) and np.array_equal(c, want_du.counts), (items, seed, rule, cutoff, "du")
    finally:
        g.set_option("stream_shape", 51209)


@pytest.mark.parametrize("variant", [2, 3])
def test_stream_kernel_side_buffer_variant_forced(g, oracle, variant):
    """The side-buffer variants of the stream kernel (2: three stages + a side buffer of a third of a tile per stage; 3: two
    stages + three output slots with side buffers of 61 % of a tile; sparse tiles give their stage back right after the
    merge, fuller tiles keep it) are chosen from a density sample on big calls only; here they are forced (stream_side = 2 / 3)
    on inputs of hundreds of tiles whose outputs are sparse, mixed and dense (identical lists: every tile overflows its side
    buffer), for the four outputs, cut-offs, and the N-list intersection chain."""
    g.set_option("stream_shape", 51209)
    g.set_option("stream_side", variant)
    try:
        for seed, (na, nb, both), kind in ((21, (300_001, 250_003, 120_000), "tail"), (22, (400_000, 400_000, 400_000), "small"),
                                           (23, (500_000, 480_000, 20_000), "tail"), (24, (3, 200_001, 2), "tail"), (25, (150_000, 7, 0), "huge"),
                                           (26, (5, 3, 2), "tail"), (27, (700_000, 100_000, 100_000), "small")):
            a, b = make_pair(seed, na, nb, both, 25, kind)
            la, lb = g.WordList.from_arrays(*a, 25), g.WordList.from_arrays(*b, 25)
            sa, sb = oracle.SList(*a, 25), oracle.SList(*b, 25)
            for cutoff in (1, 3):
                want = oracle.compare2(sa, sb, union=True, intrsec=True, diff=True, ddiff=True, cutoff=cutoff)
                for s, kw in (("union", dict(find_union=1)), ("intrsec", dict(find_intrsec=1)), ("diff1", dict(find_diff=1)), ("diff2", dict(find_ddiff=1))):
                    r = g.compare_wordmaps(la, lb, cutoff=cutoff, **kw)[s]
                    w, c = r.to_host()
                    assert np.array_equal(w, want[s].words) and np.array_equal(c, want[s].counts), (seed, cutoff, s)
                    assert (r.n_words, r.total_count) == (want[s].n_words, want[s].total_count)
        # N-list intersection: a chain of two-list links (rule min with the guard), every link through the variant
        g.set_option("use_kway", 0)
        rng = np.random.default_rng(5)
        base = np.unique(rng.integers(0, 1 << 50, size=600_000, dtype=np.uint64))
        lists, slists = [], []
        for j in range(4):
            keep = rng.random(base.size) < (0.9 if j < 2 else 0.5)
            w = base[keep]
            c = rng.integers(1, 7, size=w.size, dtype=np.uint32)
            lists.append(g.WordList.from_arrays(w, c, 25))
            slists.append(oracle.SList(w, c, 25))
        for cutoff in (1, 2):
            rc, want = oracle.intersect_multi(slists, cutoff=cutoff)
            got = g.intersect_multi(lists, cutoff=cutoff)
            w, c = got.to_host()
            assert rc == 0 and np.array_equal(w, want.words) and np.array_equal(c, want.counts), cutoff
    finally:
        g.set_option("stream_side", 1)
        g.set_option("use_kway", 1)


def test_stream_kernel_misaligned_device_views(g, oracle):
    """Caller-owned device arrays that start at every 8/4-byte phase of a 16-byte line (the TMA
    bulk copies over-fetch to 16-byte boundaries and the kernel must shift accordingly)."""
    import torch
    a, b = make_pair(31, 40_000, 50_000, 20_000, 25, "tail")
    for off_a in (0, 1, 2, 3):
        for off_b in (0, 1, 3):
            ta = torch.from_numpy(a[0][off_a:].view(np.int64).copy()).cuda()
            base_w = torch.empty(ta.numel() + 4, dtype=torch.int64, device="cuda")
            base_c = torch.empty(ta.numel() + 4, dtype=torch.int32, device="cuda")
            wa = base_w[off_a:off_a + ta.numel()]
            wa.copy_(ta)
            ca = base_c[off_a:off_a + ta.numel()]
            ca.copy_(torch.from_numpy(a[1][off_a:].view(np.int32).copy()).cuda())
            tb = torch.from_numpy(b[0][off_b:].view(np.int64).copy()).cuda()
            base_wb = torch.empty(tb.numel() + 4, dtype=torch.int64, device="cuda")
            base_cb = torch.empty(tb.numel() + 4, dtype=torch.int32, device="cuda")
            wb = base_wb[off_b:off_b + tb.numel()]
            wb.copy_(tb)
            cb = base_cb[off_b:off_b + tb.numel()]
            cb.copy_(torch.from_numpy(b[1][off_b:].view(np.int32).copy()).cuda())
            torch.cuda.synchronize()
            la = g.WordList.from_device(wa.data_ptr(), ca.data_ptr(), wa.numel(), 25, keepalive=(base_w, base_c))
            lb = g.WordList.from_device(wb.data_ptr(), cb.data_ptr(), wb.numel(), 25, keepalive=(base_wb, base_cb))
            want = oracle.compare2(oracle.SList(a[0][off_a:], a[1][off_a:], 25), oracle.SList(b[0][off_b:], b[1][off_b:], 25),
                                   union=True, intrsec=True)
            for s, kw in (("union", dict(find_union=1)), ("intrsec", dict(find_intrsec=1))):
                w, c = g.compare_wordmaps(la, lb, **kw)[s].to_host()
                assert np.array_equal(w, want[s].words) and np.array_equal(c, want[s].counts), (off_a, off_b, s)


def test_tile_kernel_still_serves_single_outputs(g, oracle):
    """use_stream_kernel=0 routes single-output merges through setop2_tile_kernel (the fused multi-output kernel)."""
    g.set_option("use_stream_kernel", 0)
    try:
        a, b = make_pair(41, 120_000, 90_000, 30_000, 25, "tail")
        la, lb = g.WordList.from_arrays(*a, 25), g.WordList.from_arrays(*b, 25)
        want = oracle.compare2(oracle.SList(*a, 25), oracle.SList(*b, 25), union=True, diff=True, cutoff=2)
        for s, kw in (("union", dict(find_union=1)), ("diff1", dict(find_diff=1))):
            w, c = g.compare_wordmaps(la, lb, cutoff=2, **kw)[s].to_host()
            assert np.array_equal(w, want[s].words) and np.array_equal(c, want[s].counts)
    finally:
        g.set_option("use_stream_kernel", 1)


def test_multi_golden_matrix(g, multi_lists):
    for gold in GOLDEN["multi"]:
        k, lists = multi_lists[gold["input"]]
        rkw = cases.rule_to_kwargs(gold["rule"])
        files, stdout, rc = {}, "", 0
        for flag, fn, fname in (("-u", g.union_multi, "union"), ("-i", g.intersect_multi, "intrsec")):
            if flag not in gold["ops"]:
                continue
            try:
                r = fn(lists, cutoff=gold["cutoff"], rule=rkw["rule"], count_override=rkw["count_override"])
                files[f"out_{k}_{fname}.list"] = r.list_bytes()
                co = fn(lists, cutoff=gold["cutoff"], rule=rkw["rule"], count_override=rkw["count_override"], countonly=1)
                assert (co.n_words, co.total_count) == (r.n_words, r.total_count)
                stdout += f"NUnique\t{r.n_words}\nNTotal\t{r.total_count}\n"
                rc = 0
            except g.GT4GPUError as e:
                assert e.code == 1          # the reference's "return 1" for a rejected rule
                stdout += "NUnique\t0\nNTotal\t0\n"
                rc = 1
        assert (rc != 0) == (gold["rc"] != 0), gold
        assert sorted(files) == sorted(gold["files"]), gold
        for name, b in files.items():
            assert refrun.digest(b) == gold["files"][name]["sha256"], (gold, name)
        assert stdout == gold["count_only_stdout"], gold


def test_multi_medium_vs_oracle(g, oracle):
    for n_lists in (1, 2, 3, 5, 8):
        lists = make_multi(300 + n_lists, n_lists, 60_000, 150_000, 32, "tail")
        gl = [g.WordList.from_arrays(w, c, 32) for w, c in lists]
        ol = [oracle.SList(w, c, 32) for w, c in lists]
        for rule in ("default", "max", "add"):
            for cutoff in (1, 40):
                rc, want = oracle.union_multi(ol, cutoff=cutoff, rule=rule)
                got = g.union_multi(gl, cutoff=cutoff, rule=rule)
                w, c = got.to_host()
                assert np.array_equal(w, want.words) and np.array_equal(c, want.counts), (n_lists, rule, cutoff)
                rc, want = oracle.intersect_multi(ol, cutoff=cutoff, rule=rule)
                got = g.intersect_multi(gl, cutoff=cutoff, rule=rule)
                w, c = got.to_host()
                assert np.array_equal(w, want.words) and np.array_equal(c, want.counts), (n_lists, rule, cutoff, "isect")


def test_write_union_and_matrix(g, oracle, tmp_path):
    lists = make_multi(77, 4, 5_000, 9_000, 20, "tail")
    gl = [g.WordList.from_arrays(w, c, 20) for w, c in lists]
    ol = [oracle.SList(w, c, 20) for w, c in lists]
    rc, want = oracle.write_union(ol, cutoff=2)
    p = tmp_path / "wu.list"
    fd = os.open(p, os.O_RDWR | os.O_CREAT, 0o644)
    h = g.gt4_write_union(gl, 2, fd)
    os.close(fd)
    assert (h.n_words, h.total_count, h.word_length) == (want.n_words, want.total_count, 20)
    assert p.read_bytes() == refrun.list_bytes(want, 20)
    h0 = g.gt4_write_union(gl, 2, 0)                      # ofile == 0: count only, header still filled
    assert (h0.n_words, h0.total_count) == (want.n_words, want.total_count)
    for is_union, fn in ((False, g.gt4_union), (True, g.gt4_is_union)):
        rc, ow, oc = oracle.union_matrix(ol, is_union=is_union)
        gw, gc = fn(gl)
        assert np.array_equal(gw, ow) and np.array_equal(gc, oc), is_union
    rows = []
    assert g.gt4_union(gl, lambda w, c, d: rows.append(w) or (1 if len(rows) == 3 else 0)) == 1 and len(rows) == 3


def test_list_files_and_header_rules(g, oracle, tmp_path):
    a, b = make_pair(9, 70_000, 50_000, 20_000, 18, "tail")
    want = oracle.compare2(oracle.SList(*a, 18), oracle.SList(*b, 18), union=True, diff=True, cutoff=2)
    for minor in (None, 0, 2):
        pa, pb = tmp_path / f"a{minor}.list", tmp_path / f"b{minor}.list"
        oracle.write_list(pa, a[0], a[1], 18, minor=minor)
        oracle.write_list(pb, b[0], b[1], 18, minor=minor)
        for stream in (False, True):
            if stream and minor in (0, 2):
                continue    # the stream reader mis-places list_start for 40-byte headers (word-list-stream.c:166-168)
            la, lb = g.WordList.open(pa, stream=stream), g.WordList.open(pb, stream=stream)
            assert (la.num_words, la.word_length, la.sum_counts) == (len(a[0]), 18, int(a[1].astype(np.uint64).sum()))
            got = g.compare_wordmaps(la, lb, find_union=1, find_diff=1, cutoff=2)
            for s in ("union", "diff1"):
                assert got[s].list_bytes() == refrun.list_bytes(want[s], 18), (minor, stream, s)
            out = tmp_path / "out.list"
            fd = os.open(out, os.O_RDWR | os.O_CREAT | os.O_TRUNC, 0o644)
            got["union"].write(fd)
            os.close(fd)
            assert out.read_bytes() == refrun.list_bytes(want["union"], 18)
    # ranges (shard loader) and rejects
    la = g.WordList.open(tmp_path / "aNone.list", first=1000, count=500)
    w, c = g.compare_wordmaps(la, g.WordList.from_arrays([], [], 18), find_union=1)["union"].to_host()
    assert np.array_equal(w, a[0][1000:1500]) and np.array_equal(c, a[1][1000:1500])
    bad = tmp_path / "bad.list"
    bad.write_bytes(b"GT4I" + bytes(60))
    with pytest.raises(g.GT4GPUError) as ei:
        g.WordList.open(bad)
    assert ei.value.code == 3
    trunc = tmp_path / "trunc.list"
    trunc.write_bytes((tmp_path / "aNone.list").read_bytes()[:-5])
    with pytest.raises(g.GT4GPUError):
        g.WordList.open(trunc)


def test_host_to_host_path_and_caller_buffers(g, oracle):
    import torch
    a, b = make_pair(21, 400_000, 300_000, 150_000, 25, "tail")
    sa, sb = oracle.SList(*a, 25), oracle.SList(*b, 25)
    want = oracle.compare2(sa, sb, union=True, intrsec=True, cutoff=1)
    from genometester4_b200 import api
    ou = np.zeros(len(a[0]) + len(b[0]), dtype=api.RECORD)
    oi = np.zeros(min(len(a[0]), len(b[0])), dtype=api.RECORD)
    n_out, t_out = api.compare2_host_records(sa.records(), sb.records(), 25, api.OP_UNION | api.OP_INTRSEC,
                                             out_records=[ou, oi, None, None])
    assert n_out[0] == want["union"].n_words and t_out[1] == want["intrsec"].total_count
    assert ou[:n_out[0]].tobytes() == want["union"].records().tobytes()
    assert oi[:n_out[1]].tobytes() == want["intrsec"].records().tobytes()
    # the pipelined variant (key-range parts, H2D || merge || D2H on three streams): force many small parts
    os.environ["GT4GPU_HOST_PART_RECORDS"] = "30000"
    try:
        want4 = oracle.compare2(sa, sb, union=True, intrsec=True, diff=True, ddiff=True, cutoff=2, rule="max")
        outs = [np.zeros(len(a[0]) + len(b[0]), dtype=api.RECORD) for _ in range(4)]
        n_out, t_out = api.compare2_host_records(sa.records(), sb.records(), 25, 15, rule="max", cutoff=2, out_records=outs)
        for q, s_ in enumerate(STREAMS):
            assert (n_out[q], t_out[q]) == (want4[s_].n_words, want4[s_].total_count), s_
            assert outs[q][:n_out[q]].tobytes() == want4[s_].records().tobytes(), s_
        n_out, t_out = api.compare2_host_records(sa.records(), sb.records(), 25, api.OP_DIFF, cutoff=2, countonly=1)
        w1 = oracle.compare2(sa, sb, diff=True, cutoff=2)["diff1"]
        assert (n_out[2], t_out[2]) == (w1.n_words, w1.total_count)
        with pytest.raises(g.GT4GPUError) as ei:
            api.compare2_host_records(sa.records(), sb.records(), 25, api.OP_UNION, out_records=[outs[0][:1000], None, None, None])
        assert ei.value.code == 5
    finally:
        del os.environ["GT4GPU_HOST_PART_RECORDS"]
    # caller-owned device buffers (torch tensors), then a too-small one
    la, lb = g.WordList.from_arrays(*a, 25), g.WordList.from_arrays(*b, 25)
    cap = len(a[0]) + len(b[0])
    tw = torch.empty(cap, dtype=torch.int64, device="cuda")
    tc = torch.empty(cap, dtype=torch.int32, device="cuda")
    r = g.compare_wordmaps(la, lb, find_union=1, out_buffers={"union": (tw.data_ptr(), tc.data_ptr(), cap)})["union"]
    torch.cuda.synchronize()
    assert np.array_equal(tw[:r.n_words].cpu().numpy().view(np.uint64), want["union"].words)
    assert np.array_equal(tc[:r.n_words].cpu().numpy().view(np.uint32), want["union"].counts)
    with pytest.raises(g.GT4GPUError) as ei:
        g.compare_wordmaps(la, lb, find_union=1, out_buffers={"union": (tw.data_ptr(), tc.data_ptr(), 1000)})
    assert ei.value.code == 5


def test_large_lists_properties(g):
    """Size-independent properties at a size no CPU check finishes quickly for (default 2e8 + 2e8 records;
    GT4GPU_TEST_UNIVERSE=1500000000 runs BASELINE config 2's full 1e9 + 1e9)."""
    import torch
    from genometester4_b200 import synth
    M = int(os.environ.get("GT4GPU_TEST_UNIVERSE", "300000000"))
    (wa, ca), (wb, cb) = synth.pair_torch(42, 25, M, 0, M, 1 / 3, 1 / 3, device="cuda")
    na, nb = wa.numel(), wb.numel()
    la = g.WordList.from_device(wa.data_ptr(), ca.data_ptr(), na, 25, keepalive=(wa, ca))
    lb = g.WordList.from_device(wb.data_ptr(), cb.data_ptr(), nb, 25, keepalive=(wb, cb))
    sum_a, sum_b = int(ca.sum(dtype=torch.int64)), int(cb.sum(dtype=torch.int64))
    res = g.compare_wordmaps(la, lb, find_union=1, find_intrsec=1, find_ddiff=1, rule="add", cutoff=0)
    u, i, d1, d2 = (res[s] for s in STREAMS)
    # cutoff 0 + rule add: union keeps every key, intersection every shared key, differences nothing
    # (f2 >= 0 always holds), so |U| = |A| + |B| - |I| and the count sums are conserved
    assert u.n_words == na + nb - i.n_words and (d1.n_words, d2.n_words) == (0, 0)
    assert u.total_count == sum_a + sum_b
    # sortedness / strict monotonicity of the union and agreement with torch's own set ops
    ut, _ = u.as_torch()
    assert bool((ut[1:] > ut[:-1]).all())            # keys < 2**50, signed compare is fine
    ref_union = torch.unique(torch.cat([wa, wb]))
    assert ref_union.numel() == u.n_words and bool((ref_union == ut).all())
    del ref_union, ut
    # difference with cut-off 5: a shared key survives iff B's count is below 5 (P-d of SURVEY 8(a))
    d = g.compare_wordmaps(la, lb, find_diff=1, cutoff=5, countonly=1)["diff1"]
    co = g.compare_wordmaps(la, lb, find_diff=1, cutoff=5)["diff1"]
    assert (d.n_words, d.total_count) == (co.n_words, co.total_count)
    # idempotence: A union A (rule max) == A ; A intersect A (default min) == A
    same = g.compare_wordmaps(la, la, find_union=1, find_intrsec=1, rule="max")
    assert same["union"].n_words == na and same["union"].total_count == sum_a
    assert same["intrsec"].n_words == na and same["intrsec"].total_count == sum_a


def test_sharded_driver_single_rank_gpu(g, oracle, tmp_path):
    """genometester4_b200/sharded.py with the real CUDA merge (world size 1: one shard = the whole key range);
    the multi-rank exchange / assembly logic is covered on CPU by tests/test_sharded_gloo.py."""
    from genometester4_b200 import sharded
    a, b = make_pair(71, 90_000, 60_000, 25_000, 22, "tail")
    oracle.write_list(tmp_path / "A.list", *a, 22)
    oracle.write_list(tmp_path / "B.list", *b, 22)
    tot = sharded.compare_files(tmp_path / "A.list", tmp_path / "B.list", str(tmp_path / "s"), find_union=1, find_diff=1, cutoff=2)
    want = oracle.compare2(oracle.SList(*a, 22), oracle.SList(*b, 22), union=True, diff=True, cutoff=2)
    assert (tmp_path / "s_22_union.list").read_bytes() == refrun.list_bytes(want["union"], 22)
    assert (tmp_path / "s_22_0_diff1.list").read_bytes() == refrun.list_bytes(want["diff1"], 22)
    assert tot["union"] == (want["union"].n_words, want["union"].total_count)
    lists = make_multi(72, 4, 20_000, 50_000, 22, "tail")
    for j, (w, c) in enumerate(lists):
        oracle.write_list(tmp_path / f"M{j}.list", w, c, 22)
    sharded.multi_files([tmp_path / f"M{j}.list" for j in range(4)], str(tmp_path / "m"), op="union", cutoff=3)
    rc, wu = oracle.union_multi([oracle.SList(w, c, 22) for w, c in lists], cutoff=3)
    assert (tmp_path / "m_22_union.list").read_bytes() == refrun.list_bytes(wu, 22)
    # virtual ranks on one GPU: merge every key range separately and concatenate (the shard loader + range merges)
    headers, bounds, _ = sharded.plan([tmp_path / "A.list", tmp_path / "B.list"], 5)
    parts = []
    for r in range(5):
        la = g.WordList.open(tmp_path / "A.list", first=int(bounds[0, r]), count=int(bounds[0, r + 1] - bounds[0, r]))
        lb = g.WordList.open(tmp_path / "B.list", first=int(bounds[1, r]), count=int(bounds[1, r + 1] - bounds[1, r]))
        parts.append(g.compare_wordmaps(la, lb, find_union=1, cutoff=2)["union"].records())
    assert np.concatenate(parts).tobytes() == want["union"].records().tobytes()


def test_unsorted_inputs_never_crash_the_device(g, oracle):
    """The reference produces garbage for lists that are not strictly ascending; so may we, but never an out-of-bounds
    access: either a (meaningless) result or error 1 comes back, and the context stays usable."""
    rng = np.random.default_rng(123)
    for trial in range(6):
        n = [50_000, 200_001, 7, 4609, 123_457, 99_999][trial]
        bad = rng.integers(0, 1 << 50, size=n, dtype=np.uint64)                      # unsorted, with duplicates
        if trial % 2:
            bad = np.sort(bad)[::-1].copy()                                             # strictly descending
        good = np.unique(rng.integers(0, 1 << 50, size=n, dtype=np.uint64))
        cnt = np.ones(n, np.uint32)
        lb_, lg_ = g.WordList.from_arrays(bad, cnt, 25), g.WordList.from_arrays(good, cnt[:good.size], 25)
        for a_, b_ in ((lb_, lg_), (lg_, lb_), (lb_, lb_)):
            for kw in (dict(find_union=1), dict(find_intrsec=1, countonly=1), dict(find_union=1, find_diff=1)):
                for use_stream in (1, 0):
                    g.set_option("use_stream_kernel", use_stream)
                    try:
                        r = g.compare_wordmaps(a_, b_, **kw)
                        for v in r.values():
                            assert v.n_words <= len(a_) + len(b_)
                    except g.GT4GPUError as e:
                        assert e.code in (1, 5), e      # "not ascending" or "more output than any valid input could give"
        g.set_option("use_stream_kernel", 1)
    # the device is still healthy
    a, b = make_pair(5, 30_000, 20_000, 10_000, 25, "tail")
    want = oracle.compare2(oracle.SList(*a, 25), oracle.SList(*b, 25), union=True)["union"]
    w, c = g.compare_wordmaps(g.WordList.from_arrays(*a, 25), g.WordList.from_arrays(*b, 25), find_union=1)["union"].to_host()
    assert np.array_equal(w, want.words) and np.array_equal(c, want.counts)
