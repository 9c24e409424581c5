"""GPU parity tests of the lookup row through the C ABI: gt4gpu_lookup against the
oracle and the committed glistquery output; the zipper as gt4gpu_compare2 (intersection, rule first)."""
import json
from pathlib import Path

import numpy as np
import pytest

from tests.test_query_host import GOLD, GOLD_DIR, lines, read_queries

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    import genometester4_b200 as g
    g.init(0)
    return g


@pytest.mark.parametrize("case", GOLD["cases"], ids=lambda c: f"k{c['k']}")
def test_lookup_matches_glistquery_golden(case, g, oracle):
    k = case["k"]
    main = g.WordList.open(GOLD_DIR / f"main_{k}.list")
    canon, counts = g.lookup(main, read_queries(k))
    assert lines(oracle, canon, counts, k) == (GOLD_DIR / f"lookup_{k}.out").read_bytes()


@pytest.mark.parametrize("case", GOLD["cases"], ids=lambda c: f"k{c['k']}")
def test_zipper_matches_glistquery_golden(case, g, oracle):
    k = case["k"]
    main = g.WordList.open(GOLD_DIR / f"main_{k}.list")
    sub = g.WordList.open(GOLD_DIR / f"sub_{k}.list")
    res = g.compare_wordmaps(sub, main, find_intrsec=1, rule=g.RULE_FIRST, cutoff=0)["intrsec"]
    w, c = res.to_host()
    assert lines(oracle, w, c, k) == (GOLD_DIR / f"zipper_{k}.out").read_bytes()


@pytest.mark.parametrize("k,n,nq", [(1, 3, 50), (9, 50_000, 200_000), (25, 1_000_000, 3_000_000), (32, 300_000, 500_000)])
def test_lookup_random_vs_oracle(k, n, nq, g, oracle):
    rng = np.random.default_rng(k)
    hi = min(4 ** k, 2 ** 63)
    words = np.unique(rng.integers(0, hi, size=n, dtype=np.uint64))
    counts = rng.integers(1, 2 ** 32, size=words.size, dtype=np.uint64).astype(np.uint32)
    lst = oracle.SList(words, counts, k)
    q = np.concatenate([rng.choice(words, size=nq // 2), rng.integers(0, hi, size=nq - nq // 2, dtype=np.uint64)])
    gl = g.WordList.from_arrays(words, counts, k)
    for canonize in (True, False):
        cw, cc = g.lookup(gl, q, canonize=canonize)
        ow, oc = oracle.lookup(lst, q, canonize=canonize)
        assert np.array_equal(cw, ow) and np.array_equal(cc, oc)
    assert (cc > 0).sum() >= nq // 2


def test_lookup_sorted_route_equals_direct_route(g, oracle, monkeypatch):
    """Large batches are sorted first (key-value radix sort) and scattered back; both routes must agree with the oracle."""
    k = 13
    rng = np.random.default_rng(44)
    words = np.unique(rng.integers(0, 4 ** k, size=400_000, dtype=np.uint64))
    counts = rng.integers(1, 1000, size=words.size).astype(np.uint32)
    gl = g.WordList.from_arrays(words, counts, k)
    lst = oracle.SList(words, counts, k)
    for nq in (1, 8191, 8193, 70_001):
        q = rng.integers(0, 4 ** k, size=nq, dtype=np.uint64)
        ow, oc = oracle.lookup(lst, q)
        for sort_min in ("1", "1000000000"):
            monkeypatch.setenv("GT4GPU_LOOKUP_SORT_MIN", sort_min)
            cw, cc = g.lookup(gl, q)
            assert np.array_equal(cw, ow) and np.array_equal(cc, oc), (nq, sort_min)
            # a batch whose canonical words are already ascending skips the sort
            cw2, cc2 = g.lookup(gl, np.sort(ow), canonize=False)
            i = np.argsort(ow, kind="stable")
            assert np.array_equal(cw2, ow[i]) and np.array_equal(cc2, oc[i])


def test_lookup_edges(g, oracle):
    empty = g.WordList.from_arrays(np.zeros(0, np.uint64), np.zeros(0, np.uint32), 16)
    cw, cc = g.lookup(empty, np.array([0, 5, 2 ** 32 - 1], dtype=np.uint64))
    assert not cc.any()
    one = g.WordList.from_arrays(np.array([2 ** 64 - 1], np.uint64), np.array([7], np.uint32), 32)
    cw, cc = g.lookup(one, np.array([2 ** 64 - 1, 0, 2 ** 64 - 2], dtype=np.uint64), canonize=False)
    assert cc.tolist() == [7, 0, 0]
    cw, cc = g.lookup(one, np.zeros(0, dtype=np.uint64))
    assert cc.size == 0


def test_lookup_device_arrays(g, oracle):
    import torch
    k = 20
    rng = np.random.default_rng(3)
    words = np.unique(rng.integers(0, 4 ** k, size=200_000, dtype=np.uint64))
    counts = rng.integers(1, 100, size=words.size).astype(np.uint32)
    gl = g.WordList.from_arrays(words, counts, k)
    q = torch.from_numpy(np.concatenate([rng.choice(words, 100_000), rng.integers(0, 4 ** k, 100_000, dtype=np.uint64)]).astype(np.int64)).cuda()
    out = torch.empty(q.numel(), dtype=torch.int32, device="cuda")
    canon = torch.empty(q.numel(), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    g.lookup_device(gl, q.data_ptr(), q.numel(), out.data_ptr(), canon.data_ptr())
    ow, oc = oracle.lookup(oracle.SList(words, counts, k), q.cpu().numpy().astype(np.uint64))
    assert np.array_equal(out.cpu().numpy().astype(np.uint32), oc)
    assert np.array_equal(canon.cpu().numpy().astype(np.uint64), ow)


def test_entry_points_are_reentrant(g, oracle):
    """glistmaker calls gt4_write_union from several worker threads at once.  Four host threads
    drive merges, list building and lookups concurrently (ctypes releases the GIL); every result must still be exact."""
    import threading
    k = 19
    errors = []

    def worker(seed):
        try:
            rng = np.random.default_rng(seed)
            for _ in range(6):
                wa = np.unique(rng.integers(0, 4 ** k, size=60_000, dtype=np.uint64))
                wb = np.unique(rng.integers(0, 4 ** k, size=50_000, dtype=np.uint64))
                wb[:20_000] = wa[:20_000]
                wb = np.unique(wb)
                ca = rng.integers(1, 100, size=wa.size).astype(np.uint32)
                cb = rng.integers(1, 100, size=wb.size).astype(np.uint32)
                la, lb = g.WordList.from_arrays(wa, ca, k), g.WordList.from_arrays(wb, cb, k)
                got = g.compare_wordmaps(la, lb, find_union=1, find_intrsec=1, cutoff=2)
                want = oracle.compare2(oracle.SList(wa, ca, k), oracle.SList(wb, cb, k), union=True, intrsec=True, cutoff=2)
                for s in ("union", "intrsec"):
                    w, c = got[s].to_host()
                    assert np.array_equal(w, want[s].words) and np.array_equal(c, want[s].counts)
                raw = rng.integers(0, 4 ** 8, size=90_000, dtype=np.uint64)
                res = g.count_words(raw, 8)
                exp = oracle.count_words(raw, 8)
                w, c = res.to_host()
                assert np.array_equal(w, exp.words) and np.array_equal(c, exp.counts)
                q = rng.integers(0, 4 ** k, size=30_000, dtype=np.uint64)
                cw, cc = g.lookup(la, q)
                ow, oc = oracle.lookup(oracle.SList(wa, ca, k), q)
                assert np.array_equal(cw, ow) and np.array_equal(cc, oc)
        except Exception as e:       # noqa: BLE001 -- reported below with the thread's seed
            errors.append((seed, repr(e)))

    threads = [threading.Thread(target=worker, args=(s,)) for s in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_query_cli_against_glistquery_golden():
    """gt4gpu-query prints byte for byte what the unmodified glistquery printed (committed outputs): dump, count
    matrices, exact lookups from a query file / a single word / a FastA file, frequency window, zipper, multi-list search."""
    import subprocess
    from genometester4_b200 import _lib
    cli = str(_lib.query_cli_path())

    def run(*args):
        r = subprocess.run([cli, *args], cwd=GOLD_DIR, capture_output=True)
        assert r.returncode == 0, (args, r.stderr)
        return r.stdout

    for k in (5, 16):
        assert run(f"main_{k}.list") == (GOLD_DIR / f"dump_{k}.out").read_bytes()
        assert run(f"main_{k}.list", f"sub_{k}.list") == (GOLD_DIR / f"matrix_{k}.out").read_bytes()
        assert run(f"main_{k}.list", f"sub_{k}.list", "--is_union", "--header") == (GOLD_DIR / f"matrix_isunion_{k}.out").read_bytes()
        assert run(f"main_{k}.list", f"sub_{k}.list", "-l", f"sub_{k}.list") == (GOLD_DIR / f"multi_{k}.out").read_bytes()
        assert run(f"main_{k}.list", "-f", f"queries_{k}.txt", "-min", "100", "-max", "500") == (GOLD_DIR / f"lookup_minmax_{k}.out").read_bytes()
    for case in GOLD["cases"]:
        k = case["k"]
        want = (GOLD_DIR / f"lookup_{k}.out").read_bytes()
        assert run(f"main_{k}.list", "-f", f"queries_{k}.txt") == want
        assert run(f"main_{k}.list", "-l", f"sub_{k}.list") == (GOLD_DIR / f"zipper_{k}.out").read_bytes()
        first = (GOLD_DIR / f"queries_{k}.txt").read_text().split()[0]
        assert run(f"main_{k}.list", "-q", first) == want.split(b"\n")[0] + b"\n"
    # -s: every k-mer of a FastA file, in file order (canonical forms, absent ones with count 0)
    import tempfile
    k = 5
    seq = "ACGTTGCAAGGCTTNACGTACGTTTGACCA"
    with tempfile.NamedTemporaryFile("w", suffix=".fa", delete=False) as f:
        f.write(">s\n" + seq + "\n")
    words = [seq[i:i + k] for i in range(len(seq) - k + 1) if "N" not in seq[i:i + k]]
    from oracle import oracle as O
    main = O.read_list(GOLD_DIR / f"main_{k}.list")
    q = np.array([sum("ACGT".index(c) << (2 * (k - 1 - j)) for j, c in enumerate(w)) for w in words], dtype=np.uint64)
    cw, cc = O.lookup(main, q)
    assert run(f"main_{k}.list", "-s", f.name) == lines(O, cw, cc, k)
