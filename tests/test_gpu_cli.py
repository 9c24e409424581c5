"""The drop-in CLI on a GPU: gt4gpu-compare must leave exactly the files, stdout and exit status that the
unmodified reference glistcompare leaves (committed golden digests; the binary itself when oracle/_ref is present)."""
import json
import subprocess
from pathlib import Path

import pytest

from genometester4_b200 import _lib
from tests import cases, refrun

pytestmark = pytest.mark.gpu
GOLDEN = json.loads((Path(__file__).parent / "golden" / "golden.json").read_text())


def run_cli(args, cwd):
    return subprocess.run([str(_lib.cli_path()), *map(str, args)], cwd=cwd, capture_output=True)


@pytest.fixture(scope="module")
def inputs(tmp_path_factory, oracle):
    d = tmp_path_factory.mktemp("cli_gpu")
    return d, refrun.write_inputs(d / "in", "pair"), refrun.write_inputs(d / "in", "multi")


def _collect(run_dir):
    return {f.name: f.read_bytes() for f in sorted(run_dir.glob("out_*"))}


def test_cli_pair_against_golden(inputs):
    d, pair_paths, _ = inputs
    run = d / "run_pair"
    run.mkdir()
    picked = [g for i, g in enumerate(GOLDEN["pair"]) if i % 24 == 0]
    assert len(picked) > 20
    for g in picked:
        for f in run.glob("out_*"):
            f.unlink()
        r = run_cli(refrun.cli_args(pair_paths[g["input"]], g["ops"], g["rule"], g["cutoff"]), run)
        assert r.returncode == g["rc"], (g, r.stderr)
        files = _collect(run)
        assert sorted(files) == sorted(g["files"]), g
        for name, b in files.items():
            assert refrun.digest(b) == g["files"][name]["sha256"], (g, name)
        for f in run.glob("out_*"):
            f.unlink()
        r = run_cli(refrun.cli_args(pair_paths[g["input"]], g["ops"], g["rule"], g["cutoff"], count_only=True), run)
        assert r.stdout.decode() == g["count_only_stdout"] and not list(run.glob("out_*")), g


def test_cli_multi_against_golden(inputs):
    d, _, multi_paths = inputs
    run = d / "run_multi"
    run.mkdir()
    picked = [g for i, g in enumerate(GOLDEN["multi"]) if i % 29 == 0]
    for g in picked:
        for f in run.glob("out_*"):
            f.unlink()
        r = run_cli(refrun.cli_args(multi_paths[g["input"]], g["ops"], g["rule"], g["cutoff"]), run)
        assert r.returncode == g["rc"], (g, r.stderr)
        files = _collect(run)
        assert sorted(files) == sorted(g["files"]), g
        for name, b in files.items():
            assert refrun.digest(b) == g["files"][name]["sha256"], (g, name)
        r = run_cli(refrun.cli_args(multi_paths[g["input"]], g["ops"], g["rule"], g["cutoff"], count_only=True), run)
        assert r.stdout.decode() == g["count_only_stdout"], g


def test_cli_vs_reference_binary_options(inputs, oracle):
    """-o, --stream, --print_operation, -D side by side with the real binary (skipped where oracle/_ref is absent)."""
    if oracle.ref_binary("glistcompare") is None:
        pytest.skip("oracle/_ref not built")
    d, pair_paths, multi_paths = inputs
    for k, args in enumerate([
        [*pair_paths["p_tail"], "-u", "-i", "-dd", "-o", "x/y", "-c", "2"],
        [*pair_paths["p_small"], "-du", "--stream", "--print_operation"],
        [*pair_paths["p_huge"], "-u", "-r", "add", "--disable_scouts", "--count_only", "--print_operation"],
        [*multi_paths["m8_tail"], "-u", "-i", "-r", "max", "-o", "mm"],
        [*multi_paths["m4_one_empty"], "-u", "-i", "--count_only", "--stream"],
    ]):
        outs = []
        for who in ("mine", "ref"):
            run = d / f"opt_{k}_{who}"
            (run / "x").mkdir(parents=True)
            r = run_cli(args, run) if who == "mine" else oracle.run_ref("glistcompare", args, cwd=run)
            outs.append((r.returncode, r.stdout, {str(f.relative_to(run)): f.read_bytes() for f in sorted(run.rglob("*.list"))}))
        assert outs[0] == outs[1], args
