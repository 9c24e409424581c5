"""Host-side behaviour of the gt4gpu-compare CLI that needs no GPU: flag grammar, validation order,
messages and exit codes, compared with the unmodified reference binary when it is available."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

from genometester4_b200 import _lib
from tests import refrun

CLI = _lib.cli_path()


def run_cli(args, cwd):
    return subprocess.run([str(CLI), *map(str, args)], cwd=cwd, capture_output=True)


@pytest.fixture(scope="module")
def files(tmp_path_factory, oracle):
    d = tmp_path_factory.mktemp("cli")
    rng = np.random.default_rng(3)
    for name, k in (("a16", 16), ("b16", 16), ("c16", 16), ("d20", 20)):
        w = np.unique(rng.integers(0, 1 << 30, size=50, dtype=np.uint64))
        oracle.write_list(d / f"{name}.list", w, np.ones(w.size, np.uint32), k)
    (d / "junk.bin").write_bytes(b"not a list file at all")
    return d


CASES = [
    [],
    ["-v"], ["--version"], ["-h"], ["--help"], ["-?"],
    ["a16.list"],
    ["a16.list", "-u"],
    ["a16.list", "b16.list", "--bogus"],
    ["a16.list", "b16.list", "-c", "x1"],
    ["a16.list", "d20.list", "-u"],
    ["a16.list", "junk.bin", "-u"],
    ["a16.list", "b16.list", "c16.list", "-d"],
    ["a16.list", "b16.list", "c16.list"],
    ["a16.list", "b16.list", "-u", "-r", "min"],
    ["a16.list", "b16.list", "-u", "-r", "first"],
    ["a16.list", "b16.list", "-u", "-r", "subtract"],
    ["a16.list", "b16.list", "-u", "-o", "x" * 201],
    ["a16.list", "b16.list", "-r"],
]


@pytest.mark.parametrize("args", CASES, ids=lambda a: " ".join(a)[:40] or "noargs")
def test_cli_validation_matches_reference(args, files, oracle):
    assert CLI.exists(), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    mine = run_cli(args, files)
    ref = oracle.run_ref("glistcompare", args, cwd=files)
    if ref is None:
        pytest.skip("oracle/_ref not built")
    assert mine.returncode == ref.returncode, (mine.stderr, ref.stderr)
    assert mine.stdout == ref.stdout
    if "junk.bin" in args:
        # the reference goes on to dereference a container it never created and prints an assertion
        # with its own source path; only the diagnostic and the exit status are comparable
        assert mine.stderr.split(b"\n")[0] == ref.stderr.split(b"\n")[0] and mine.stderr.endswith(b"Stopping...\n")
    else:
        assert mine.stderr == ref.stderr


def test_cli_static_expectations(files):
    """The same facts, hard-coded, for boxes without the reference binary."""
    r = run_cli(["-v"], files)
    assert r.returncode == 0 and r.stdout == b"glistcompare version 4.2.16 (stable)\n"
    r = run_cli([], files)
    assert r.returncode == 1 and r.stdout.startswith(b"glistcompare version 4.2.16 (stable)\nUsage: glistcompare INPUTLIST1")
    r = run_cli(["a16.list", "d20.list", "-u"], files)
    assert r.returncode == 1 and b"has different word length (20 != 16)" in r.stderr and r.stderr.endswith(b"Stopping...\n")
    r = run_cli(["a16.list", "b16.list", "-u", "-mm", "2"], files)
    assert r.returncode == 1 and b"not supported" in r.stderr
