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


def test_gt4i_index_inputs_against_golden(tmp_path):
    """GT4I index files as inputs (the reference reads their k-mer table through the same iterator interface,
    /root/reference/src/index-map.c:122-158): library loader vs a numpy parse, CLI vs the reference's outputs."""
    import numpy as np
    import genometester4_b200 as g
    idx_dir = Path(__file__).parent / "golden" / "index"
    golden = json.loads((Path(__file__).parent / "golden" / "index_golden.json").read_text())
    raw = (idx_dir / "x_16.index").read_bytes()
    hdr = np.frombuffer(raw, dtype=[("code", "<u4"), ("major", "<u4"), ("minor", "<u4"), ("k", "<u4"), ("n", "<u8"), ("nloc", "<u8"),
                                    ("fb", "<u4"), ("sb", "<u4"), ("pb", "<u4"), ("fill", "<u4"), ("files", "<u8"), ("kmers", "<u8"), ("locs", "<u8")], count=1)[0]
    rec = np.frombuffer(raw, dtype=[("word", "<u8"), ("loc", "<u8")], count=int(hdr["n"]), offset=int(hdr["kmers"]))
    counts = np.diff(np.concatenate([rec["loc"], [hdr["nloc"]]]).astype(np.uint64)).astype(np.uint32)
    g.init(0)
    l = g.WordList.open(idx_dir / "x_16.index")
    assert (l.num_words, l.word_length, l.sum_counts) == (int(hdr["n"]), 16, int(hdr["nloc"]))
    w, c = g.compare_wordmaps(l, g.WordList.from_arrays([], [], 16), find_union=1, cutoff=0)["union"].to_host()
    assert np.array_equal(w, rec["word"]) and np.array_equal(c, counts)
    part = g.WordList.open(idx_dir / "x_16.index", first=100, count=250)      # a shard that ends inside the table
    w, c = g.compare_wordmaps(part, g.WordList.from_arrays([], [], 16), find_union=1, cutoff=0)["union"].to_host()
    assert np.array_equal(w, rec["word"][100:350]) and np.array_equal(c, counts[100:350])
    for case in golden:
        for f in tmp_path.glob("out_*"):
            f.unlink()
        r = run_cli([idx_dir / f for f in case["files"]] + case["flags"], tmp_path)
        assert r.returncode == case["rc"], r.stderr
        got = {f.name: refrun.digest(f.read_bytes()) for f in sorted(tmp_path.glob("out_*"))}
        assert got == case["outputs"], case
        for f in tmp_path.glob("out_*"):
            f.unlink()
        r = run_cli([idx_dir / f for f in case["files"]] + case["flags"] + ["--count_only"], tmp_path)
        assert r.stdout.decode() == case["count_only_stdout"], case


def test_config1_fasta_built_lists(tmp_path, oracle):
    """BASELINE config 1 on the GPU: the two ~1.1 M-k-mer k=16 lists built by the reference glistmaker from synthetic
    FASTA, `-i` (and `-u -d`) through the CLI, byte-identical to the reference glistcompare."""
    paths = refrun.build_config1_lists(tmp_path / "c1")
    if paths is None:
        pytest.skip("oracle/_ref not built")
    for flags in (["-i"], ["-u", "-d", "-c", "2"]):
        outs = []
        for who in ("mine", "ref"):
            run = tmp_path / f"{who}_{len(flags)}"
            run.mkdir()
            r = run_cli([*paths, *flags], run) if who == "mine" else oracle.run_ref("glistcompare", [*paths, *flags], cwd=run)
            assert r.returncode == 0
            outs.append({f.name: f.read_bytes() for f in sorted(run.glob("out_*"))})
        assert outs[0] == outs[1] and len(outs[0]) == len([f for f in flags if f in ("-i", "-u", "-d")])


def test_file_pipeline_many_parts_against_oracle(tmp_path, oracle, monkeypatch):
    """gt4gpu_compare2_files (the path gt4gpu-compare takes for two list files): forced to cut the key space into many
    parts, every output file byte-identical to what the reference writes; count-only totals; headers of every minor
    version on the input side."""
    import numpy as np

    import genometester4_b200 as g
    from tests.util import make_pair
    g.init(0)
    monkeypatch.setenv("GT4GPU_HOST_PART_RECORDS", "50000")
    a, b = make_pair(91, 400_000, 350_000, 150_000, 25, "tail")
    pa, pb = tmp_path / "A.list", tmp_path / "B.list"
    oracle.write_list(pa, a[0], a[1], 25)
    oracle.write_list(pb, b[0], b[1], 25, minor=0)           # a 4.0 header (40 bytes) on one side
    sa, sb = oracle.SList(*a, 25), oracle.SList(*b, 25)
    tags = {"union": "union", "intrsec": "intrsec", "diff1": "0_diff1", "diff2": "0_diff2"}
    for kw, cutoff in ((dict(find_union=1), 1), (dict(find_intrsec=1, find_diff=1), 3), (dict(find_union=1, find_intrsec=1, find_diff=1, find_ddiff=1), 2),
                       (dict(find_diff=1, subtract=1), 1)):
        okw = dict(union=bool(kw.get("find_union")), intrsec=bool(kw.get("find_intrsec")), diff=bool(kw.get("find_diff")), ddiff=bool(kw.get("find_ddiff")),
                   subtract=bool(kw.get("subtract")))
        want = oracle.compare2(sa, sb, cutoff=cutoff, **okw)
        tot = g.api.compare_files(pa, pb, str(tmp_path / "o"), cutoff=cutoff, **kw)
        assert sorted(tot) == sorted(want)
        for s, r in want.items():
            assert tot[s] == (r.n_words, r.total_count), (kw, s)
            assert (tmp_path / f"o_25_{tags[s]}.list").read_bytes() == refrun.list_bytes(r, 25), (kw, s)
            (tmp_path / f"o_25_{tags[s]}.list").unlink()
        co = g.api.compare_files(pa, pb, str(tmp_path / "o"), cutoff=cutoff, countonly=1, **kw)
        assert co == tot and not list(tmp_path.glob("o_25_*"))
    # an empty list on either side, and a single part
    pe = tmp_path / "E.list"
    oracle.write_list(pe, np.zeros(0, np.uint64), np.zeros(0, np.uint32), 25)
    monkeypatch.setenv("GT4GPU_HOST_PART_RECORDS", "100000000")
    for x, y, sx, sy in ((pa, pe, sa, None), (pe, pb, None, sb), (pa, pb, sa, sb)):
        ex = sx or oracle.SList(np.zeros(0, np.uint64), np.zeros(0, np.uint32), 25)
        ey = sy or oracle.SList(np.zeros(0, np.uint64), np.zeros(0, np.uint32), 25)
        want = oracle.compare2(ex, ey, union=True, diff=True, cutoff=1)
        tot = g.api.compare_files(x, y, str(tmp_path / "e"), find_union=1, find_diff=1, cutoff=1)
        for s, r in want.items():
            assert tot[s] == (r.n_words, r.total_count)
            assert (tmp_path / f"e_25_{tags[s]}.list").read_bytes() == refrun.list_bytes(r, 25)


def test_cli_gpus_n_matches_one_process(inputs, tmp_path):
    """gt4gpu-compare --gpus N (one process per key-range shard, offsets exchanged over pipes, parallel pwrite): the files,
    stdout and exit status of the one-process run, for two-list and N-list modes.  On a one-GPU box the shards share it."""
    d, pair_paths, multi_paths = inputs
    jobs = [(pair_paths["p_tail"], ["-u", "-i", "-dd", "-c", "2"]), (pair_paths["p_small"], ["-du", "-i"]), (pair_paths["p_extremes"], ["-u", "-d"]),
            (pair_paths["p_a_empty"], ["-u", "-d"]), (multi_paths["m8_tail"], ["-u", "-i", "-c", "2"]), (multi_paths["m4_one_empty"], ["-u"]),
            (multi_paths["m3_small"], ["-u", "-i", "-r", "min"]), (multi_paths["m4_huge"], ["-u", "-r", "max"])]
    for n_gpus in (2, 3):
        for paths, flags in (jobs if n_gpus == 2 else jobs[:1] + jobs[4:5]):       # (every process start pays a CUDA context)
            outs = {}
            for tag, extra in (("one", []), ("many", ["--gpus", str(n_gpus)])):
                run = tmp_path / f"{tag}_{n_gpus}"
                run.mkdir(exist_ok=True)
                for f in run.glob("out_*"):
                    f.unlink()
                r = run_cli([*paths, *flags, *extra], run)
                rc = run_cli([*paths, *flags, "--count_only", *extra], run)
                outs[tag] = (r.returncode, _collect(run), rc.returncode, rc.stdout)
            assert outs["one"] == outs["many"], (n_gpus, flags)
