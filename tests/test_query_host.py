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


# ---- gt4gpu-query: flag grammar, validation and -stat (no device needed on these paths)

QUERY_CLI_CASES = [
    [],
    ["-v"], ["--version"], ["-h"],
    ["-q", "ACGTA"],
    ["main_5.list", "--bogus"],
    ["nosuch.list"],
    ["queries_5.txt"],
    ["main_5.list", "main_16.list"],
    ["main_5.list", "-stat"], ["main_5.list", "sub_5.list", "--stats"],
    ["main_5.list", "-min", "x"], ["main_5.list", "-max", "3y"],
    ["main_5.list", "sub_5.list", "-q", "ACGTA"],
    ["main_5.list", "-l", "sub_16.list"],
    ["main_5.list", "-mm"], ["main_5.list", "-p", "40"],
]


@pytest.mark.parametrize("args", QUERY_CLI_CASES, ids=lambda a: " ".join(a)[:40] or "noargs")
def test_query_cli_validation_matches_reference(args, oracle):
    import subprocess
    from genometester4_b200 import _lib
    cli = _lib.query_cli_path()
    assert cli.exists(), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    if oracle.ref_binary("glistquery") is None:
        pytest.skip("oracle/_ref not built")
    mine = subprocess.run([str(cli), *args], cwd=GOLD_DIR, capture_output=True)
    ref = oracle.run_ref("glistquery", args, cwd=GOLD_DIR, timeout=20)
    assert mine.returncode == ref.returncode, (mine.stderr, ref.stderr)
    assert mine.stdout == ref.stdout
    if args == ["queries_5.txt"]:
        # the reference goes on after its diagnostic (an assertion with its own source path)
        assert mine.stderr.split(b"\n")[0] == ref.stderr.split(b"\n")[0]
    else:
        assert mine.stderr == ref.stderr


def test_query_cli_static_expectations():
    import subprocess
    from genometester4_b200 import _lib
    cli = _lib.query_cli_path()
    r = subprocess.run([str(cli), "-v"], capture_output=True)
    assert r.returncode == 0 and r.stdout == b"glistquery version 4.2.16 (stable)\n"
    r = subprocess.run([str(cli), "main_16.list", "-stat"], cwd=GOLD_DIR, capture_output=True)
    assert r.returncode == 0 and r.stdout == (GOLD_DIR / "stat_16.out").read_bytes()
    r = subprocess.run([str(cli), "main_16.list", "-q", "ACGT", "-mm", "1"], cwd=GOLD_DIR, capture_output=True)
    assert r.returncode == 1 and b"not supported" in r.stderr
