"""Parity at scale against the unmodified reference: two lists of ~5e7 k-mers each (1.2 GB of list files) and eight lists
of ~1e7, merged by oracle/_ref/glistcompare and by gt4gpu-compare file to file; every output file must be byte-identical
and the count-only lines equal.  This is the "1e8-scale slice" check of SURVEY section 7: the sizes are far beyond what the
Python oracle can merge, so the reference binary itself is the checker (skipped when oracle/_ref was not built).
GT4GPU_SCALE_TEST_N overrides the records per list."""
import hashlib
import os
import shutil
import subprocess
import tempfile
from pathlib import Path

import numpy as np
import pytest

from genometester4_b200 import _lib, synth

pytestmark = pytest.mark.gpu
N_PER_LIST = int(float(os.environ.get("GT4GPU_SCALE_TEST_N", "5e7")))


def _sha(path: Path) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for block in iter(lambda: f.read(1 << 24), b""):
            h.update(block)
    return h.hexdigest()


def _outputs(run_dir: Path):
    return {f.name: (f.stat().st_size, _sha(f)) for f in sorted(run_dir.glob("out_*"))}


def _both(oracle, args, work: Path, tag: str):
    """Runs the reference and the drop-in with the same arguments in two fresh directories; returns their outputs + stdout."""
    res = []
    for who in ("ref", "mine"):
        run = work / f"{tag}_{who}"
        run.mkdir()
        if who == "ref":
            r = oracle.run_ref("glistcompare", args, cwd=run, timeout=900, attempts=1)
        else:
            r = subprocess.run([str(_lib.cli_path()), *map(str, args)], cwd=run, capture_output=True, timeout=900)
        assert r.returncode == 0, (who, args, r.stderr[-500:])
        res.append((_outputs(run), r.stdout))
        if who == "ref":
            keep = res[0]
        shutil.rmtree(run)            # 1-2 GB per run: make room before the next one
    return keep, res[1]


@pytest.fixture(scope="module")
def work(oracle):
    if oracle.ref_binary("glistcompare") is None:
        pytest.skip("oracle/_ref not built")
    base = "/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > (24 << 30) else None
    d = Path(tempfile.mkdtemp(prefix="gt4gpu_scale_", dir=base))
    yield d
    shutil.rmtree(d, ignore_errors=True)


def _write(oracle, path, w, c, k):
    oracle.write_list(path, w.cpu().numpy().view(np.uint64), c.cpu().numpy().view(np.uint32), k)


def test_two_lists_5e7_byte_identical_to_reference(work, oracle):
    universe = int(1.5 * N_PER_LIST)
    (wa, ca), (wb, cb) = synth.pair_torch(7, 25, universe, 0, universe, 1 / 3, 1 / 3)
    a, b = work / "A_25.list", work / "B_25.list"
    _write(oracle, a, wa, ca, 25)
    _write(oracle, b, wb, cb, 25)
    n_in = wa.numel() + wb.numel()
    del wa, ca, wb, cb
    assert n_in > 1.9 * N_PER_LIST
    for tag, flags in (("ui", ["-u", "-i"]), ("dc5", ["-d", "-c", "5"]), ("dd_max", ["-dd", "-r", "max", "-c", "2"]), ("du", ["-du"])):
        (ref_files, _), (my_files, _) = _both(oracle, [a, b, *flags], work, tag)
        assert ref_files and my_files == ref_files, (flags, ref_files, my_files)
        (_, ref_out), (_, my_out) = _both(oracle, [a, b, *flags, "--count_only"], work, tag + "_co")
        assert my_out == ref_out and b"NUnique" in ref_out, (flags, ref_out, my_out)
        if tag == "ui":
            # the same files from three key-range shards (one process each, sharing the device on a one-GPU box): every shard
            # writes its slice of each output at an arbitrary record offset while the others write theirs
            run = work / "ui_shards"
            run.mkdir()
            r = subprocess.run([str(_lib.cli_path()), str(a), str(b), *flags, "--gpus", "3"], cwd=run, capture_output=True, timeout=900)
            assert r.returncode == 0, r.stderr[-500:]
            assert _outputs(run) == ref_files
            shutil.rmtree(run)
    a.unlink()
    b.unlink()


def test_eight_lists_byte_identical_to_reference(work, oracle):
    m = int(3 * N_PER_LIST / 5)               # universe; every list keeps a third of it (1e7 records at the default size)
    paths = []
    for j in range(8):
        w, c = synth.list_torch(5, 32, m, 0, m, j, 1 / 3)
        p = work / f"L{j}_32.list"
        _write(oracle, p, w, c, 32)
        paths.append(p)
    for tag, flags in (("u8", ["-u"]), ("u8c3", ["-u", "-c", "3", "-r", "max"]), ("i8", ["-i"])):
        (ref_files, ref_out), (my_files, my_out) = _both(oracle, [*paths, *flags], work, tag)
        assert ref_files and my_files == ref_files, (flags, ref_files, my_files)
        assert my_out == ref_out
    for p in paths:
        p.unlink()
