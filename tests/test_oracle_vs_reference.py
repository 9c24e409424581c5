"""Pins the oracle: (1) differentially against the unmodified reference binary
(oracle/_ref, when present) over the FULL flag matrix, (2) against the committed
golden digests (always)."""
import json
from pathlib import Path

import pytest

from tests import cases, refrun

GOLDEN = json.loads((Path(__file__).parent / "golden" / "golden.json").read_text())


def _have_ref(oracle):
    return oracle.ref_binary("glistcompare") is not None


def test_golden_inputs_are_reproducible(oracle, tmp_path):
    for which in ("pair", "multi"):
        paths = refrun.write_inputs(tmp_path / which, which)
        for name, ps in paths.items():
            assert [refrun.digest(p.read_bytes()) for p in ps] == GOLDEN["inputs"][name], name


def test_oracle_matches_golden_pair(oracle):
    assert len(GOLDEN["pair"]) > 300
    for g in GOLDEN["pair"]:
        files, stdout = refrun.oracle_pair(g["input"], tuple(g["ops"]), g["rule"], g["cutoff"])
        assert g["rc"] == 0
        assert sorted(files) == sorted(g["files"]), g
        for name, b in files.items():
            assert refrun.digest(b) == g["files"][name]["sha256"], (g, name)
        assert stdout == g["count_only_stdout"], g


def test_oracle_matches_golden_multi(oracle):
    assert len(GOLDEN["multi"]) > 100
    for g in GOLDEN["multi"]:
        files, stdout, rc = refrun.oracle_multi(g["input"], tuple(g["ops"]), g["rule"], g["cutoff"])
        assert rc == g["rc"], g
        assert sorted(files) == sorted(g["files"]), g
        for name, b in files.items():
            assert refrun.digest(b) == g["files"][name]["sha256"], (g, name)
        assert stdout == g["count_only_stdout"], g


def test_oracle_vs_reference_binary_pair_full(oracle, tmp_path):
    if not _have_ref(oracle):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    paths = refrun.write_inputs(tmp_path / "in", "pair")
    n = 0
    for name, ops, rule, cutoff in cases.pair_cases(full=True):
        rc, ref_files, _ = refrun.run_reference(tmp_path / "run", paths[name], ops, rule, cutoff)
        files, stdout = refrun.oracle_pair(name, ops, rule, cutoff)
        assert rc == 0
        assert files == ref_files, (name, ops, rule, cutoff)
        n += 1
    assert n > 1000


def test_oracle_vs_reference_binary_stream_and_countonly(oracle, tmp_path):
    if not _have_ref(oracle):
        pytest.skip("oracle/_ref not built")
    paths = refrun.write_inputs(tmp_path / "in", "pair")
    for name in ("p_tail", "p_huge", "p_extremes", "p_a_empty"):
        for ops in cases.PAIR_OPS:
            rc, ref_files, _ = refrun.run_reference(tmp_path / "run", paths[name], ops, "default", 2, extra=("--stream",))
            files, stdout = refrun.oracle_pair(name, ops, "default", 2)
            assert rc == 0 and files == ref_files, (name, ops)
            rc, ref_files, ref_stdout = refrun.run_reference(tmp_path / "run", paths[name], ops, "default", 2, count_only=True)
            assert rc == 0 and ref_files == {} and ref_stdout == stdout, (name, ops)


def test_oracle_vs_reference_binary_multi_full(oracle, tmp_path):
    if not _have_ref(oracle):
        pytest.skip("oracle/_ref not built")
    paths = refrun.write_inputs(tmp_path / "in", "multi")
    n = 0
    for name, ops, rule, cutoff in cases.multi_cases(full=True):
        for extra in ((), ("--stream",)):
            rc, ref_files, _ = refrun.run_reference(tmp_path / "run", paths[name], ops, rule, cutoff, extra=extra)
            files, stdout, orc = refrun.oracle_multi(name, ops, rule, cutoff)
            assert (rc != 0) == (orc != 0), (name, ops, rule, cutoff)
            assert files == ref_files, (name, ops, rule, cutoff, extra)
        rc, ref_files, ref_stdout = refrun.run_reference(tmp_path / "run", paths[name], ops, rule, cutoff, count_only=True)
        assert ref_stdout == stdout, (name, ops, rule, cutoff)
        n += 1
    assert n > 300


def test_config1_glistmaker_lists_intersection(oracle, tmp_path):
    """BASELINE config 1 end to end on the CPU: glistmaker-built k=16 lists, `glistcompare -i`."""
    paths = refrun.build_config1_lists(tmp_path / "c1")
    if paths is None:
        pytest.skip("oracle/_ref not built")
    a, b = oracle.read_list(paths[0]), oracle.read_list(paths[1])
    assert len(a) > 1_000_000 and len(b) > 1_000_000 and a.word_length == 16
    rc, ref_files, _ = refrun.run_reference(tmp_path / "run", paths, ("-i",), "default", 1)
    want = oracle.compare2(a, b, intrsec=True)["intrsec"]
    assert rc == 0 and ref_files == {"out_16_intrsec.list": refrun.list_bytes(want, 16)}
    rc, _, stdout = refrun.run_reference(tmp_path / "run", paths, ("-i",), "default", 1, count_only=True)
    assert stdout == f"NUnique\t{want.n_words}\nNTotal\t{want.total_count}\n"
