"""Host-only file output of the library (genometester4_b200/csrc/gt4gpu_fileio.h) on the CPU: large spans go into a shared
mapping of the output file from several threads (pwrite serialises on the inode lock), small or unmappable ones through
pwrite.  Checked here: arbitrary (unaligned) offsets, the header in front of a span survives, sequential mode moves the
descriptor, several PROCESSES writing disjoint ranges of one file at once (the sharded path) never shrink or clobber
each other's part, GT4GPU_NO_MAPPED_WRITES falls back, and non-regular files are declined."""
import ctypes as C
import multiprocessing as mp
import os
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "tests" / "emulate_fileio.cpp"
SO = ROOT / "tests" / "_build" / "libemu_fileio.so"
MB = 1 << 20


def _build():
    SO.parent.mkdir(exist_ok=True)
    hdr = ROOT / "genometester4_b200" / "csrc" / "gt4gpu_fileio.h"
    if not SO.exists() or SO.stat().st_mtime < max(SRC.stat().st_mtime, hdr.stat().st_mtime):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", f"-I{hdr.parent}", "-o", str(SO), str(SRC)], check=True)
    lib = C.CDLL(str(SO))
    lib.fileio_write.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_int]
    lib.fileio_check.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64, C.c_uint32]
    return lib


@pytest.fixture(scope="module")
def lib():
    return _build()


@pytest.fixture()
def big_dir(tmp_path):
    """a directory with room for a few hundred MB (tmpfs when there is one: that is where the bench and the CLI runs write)"""
    shm = Path("/dev/shm")
    if shm.is_dir() and os.statvfs(shm).f_bavail * os.statvfs(shm).f_frsize > (2 << 30):
        d = shm / f"gt4gpu_fileio_{os.getpid()}"
        d.mkdir(exist_ok=True)
        yield d
        for f in d.iterdir():
            f.unlink()
        d.rmdir()
    else:
        yield tmp_path


def test_spans_at_unaligned_offsets_keep_the_header(lib, big_dir):
    p = big_dir / "one.list"
    for at, n in ((48, 100 * MB + 7), (48 + 12 * 12345677, 64 * MB + 11), (4096, 33 * MB), (48, 5 * MB + 1), (1, 40 * MB)):
        p.write_bytes(b"H" * 48)
        assert lib.fileio_write(str(p).encode(), at, n, 3, 0, 0) == 0
        assert p.stat().st_size == at + n
        assert lib.fileio_check(str(p).encode(), at, n, 3) == 0
        if at >= 48:
            assert p.read_bytes()[:48] == b"H" * 48
        p.unlink()


def test_mapped_path_is_taken_for_regular_files_and_declined_otherwise(lib, big_dir, monkeypatch):
    p = big_dir / "two.list"
    assert lib.fileio_write(str(p).encode(), 12345, 48 * MB, 5, 1, 0) == 0            # write_mapped itself
    assert lib.fileio_check(str(p).encode(), 12345, 48 * MB, 5) == 0
    p.unlink()
    assert lib.fileio_write(b"/dev/null", 0, 40 * MB, 5, 1, 0) == 100                  # not a regular file
    assert lib.fileio_write(b"/dev/null", 0, 40 * MB, 5, 0, 0) == 0                    # ... pwrite takes it
    env = {**os.environ, "GT4GPU_NO_MAPPED_WRITES": "1"}
    code = (f"import ctypes as C; l = C.CDLL({str(SO)!r}); l.fileio_write.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_int]; "
            f"import sys; sys.exit(0 if l.fileio_write({str(p)!r}.encode(), 48, {40 * MB}, 9, 1, 0) == 100 and l.fileio_write({str(p)!r}.encode(), 48, {40 * MB}, 9, 0, 0) == 0 else 1)")
    assert subprocess.run(["python", "-c", code], env=env).returncode == 0            # switched off: declined, pwrite writes
    assert lib.fileio_check(str(p).encode(), 48, 40 * MB, 9) == 0
    p.unlink()


def test_sequential_mode_moves_the_descriptor(lib, big_dir):
    p = big_dir / "three.list"
    p.write_bytes(b"H" * 48)
    assert lib.fileio_write(str(p).encode(), 48, 36 * MB + 5, 7, 0, 1) == 0
    assert lib.fileio_check(str(p).encode(), 48, 36 * MB + 5, 7) == 0
    p.unlink()


def _shard(args):
    so, path, at, n, seed = args
    l = C.CDLL(so)
    l.fileio_write.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_int]
    return l.fileio_write(path.encode(), at, n, seed, 0, 0)


def test_processes_write_disjoint_ranges_of_one_file(lib, big_dir):
    """the sharded path: every process extends the file to its own end (never shrinking it) and fills its own range"""
    p = big_dir / "four.list"
    p.write_bytes(b"H" * 48)
    sizes = [40 * MB + 12 * 3, 35 * MB + 12 * 7, 12 * 1000, 50 * MB, 33 * MB + 12]
    offs, at = [], 48
    for n in sizes:
        offs.append(at)
        at += n
    jobs = [(str(SO), str(p), o, n, 11) for o, n in zip(offs, sizes)]
    with mp.get_context("spawn").Pool(len(jobs)) as pool:
        assert pool.map(_shard, list(reversed(jobs))) == [0] * len(jobs)      # (the last range first: the others must not cut it off)
    assert p.stat().st_size == at
    assert lib.fileio_check(str(p).encode(), 48, at - 48, 11) == 0
    assert p.read_bytes()[:48] == b"H" * 48
    p.unlink()
