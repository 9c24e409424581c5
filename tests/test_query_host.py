"""CPU tests of the lookup row: the oracle's restatement of glistquery's exact
lookups and of the list-against-list zipper against the committed output of the unmodified glistquery binary."""
import json
from pathlib import Path

import numpy as np
import pytest

GOLD_DIR = Path(__file__).parent / "golden" / "query"
GOLD = json.loads((GOLD_DIR / "query_golden.json").read_text())


def string_to_word(s: str) -> int:
    w = 0
    for ch in s:
        w = (w << 2) | "ACGT".index(ch)
    return w


def read_queries(k):
    return np.array([string_to_word(l) for l in (GOLD_DIR / f"queries_{k}.txt").read_text().split()], dtype=np.uint64)


def lines(oracle, words, counts, k) -> bytes:
    return "".join(f"{oracle.word_to_string(w, k)}\t{int(c)}\n" for w, c in zip(words, counts)).encode()


@pytest.mark.parametrize("case", GOLD["cases"], ids=lambda c: f"k{c['k']}")
def test_oracle_lookup_matches_glistquery(case, oracle):
    k = case["k"]
    main = oracle.read_list(GOLD_DIR / f"main_{k}.list")
    canon, counts = oracle.lookup(main, read_queries(k))
    assert lines(oracle, canon, counts, k) == (GOLD_DIR / f"lookup_{k}.out").read_bytes()


@pytest.mark.parametrize("case", GOLD["cases"], ids=lambda c: f"k{c['k']}")
def test_oracle_zipper_is_intersection_rule_first(case, oracle):
    """search_list_zipper (src/glistquery.c:702-717) prints the query list's record for every shared word: the
    intersection of (query list, list) under rule `first`."""
    k = case["k"]
    main = oracle.read_list(GOLD_DIR / f"main_{k}.list")
    sub = oracle.read_list(GOLD_DIR / f"sub_{k}.list")
    res = oracle.compare2(sub, main, intrsec=True, rule="first", cutoff=0)["intrsec"]
    assert lines(oracle, res.words, res.counts, k) == (GOLD_DIR / f"zipper_{k}.out").read_bytes()


def test_oracle_lookup_vs_live_glistquery(oracle, tmp_path):
    if oracle.ref_binary("glistquery") is None:
        pytest.skip("oracle/_ref/glistquery not built here")
    rng = np.random.default_rng(5)
    k = 7
    words = np.unique(rng.integers(0, 4 ** k, size=6000, dtype=np.uint64))
    canon, _ = oracle.lookup(oracle.SList(words, np.ones(words.size, np.uint32), k), words)   # canonical forms
    words = np.unique(canon)
    counts = rng.integers(1, 2 ** 32, size=words.size, dtype=np.uint64).astype(np.uint32)
    oracle.write_list(tmp_path / "m.list", words, counts, k)
    q = rng.integers(0, 4 ** k, size=2000, dtype=np.uint64)
    (tmp_path / "q.txt").write_text("".join(oracle.word_to_string(w, k) + "\n" for w in q))
    r = oracle.run_ref("glistquery", ["m.list", "-f", "q.txt"], cwd=tmp_path, check=True, timeout=30)
    cw, cc = oracle.lookup(oracle.SList(words, counts, k), q)
    assert lines(oracle, cw, cc, k) == r.stdout
    assert (cc > 0).sum() > 20 and (cc == 0).sum() > 20
