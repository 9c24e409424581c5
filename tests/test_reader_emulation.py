"""CPU check of the device sequence readers' arithmetic core: the chunk / thread decomposition (line-state scan with
carries, per-thread walks, code compaction, windowed word assembly) is replayed on the host with the same
__host__ __device__ functions the CUDA kernels call (gt4gpu_fasta_core.cuh) and compared with the oracle -- over the
committed glistmaker inputs and over fuzzed images, at chunk shapes that put every boundary everywhere."""
import ctypes as C
import json
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "tests" / "emulate_reader.cpp"
SO = ROOT / "tests" / "_build" / "libemulate_reader.so"
GOLD_DIR = ROOT / "tests" / "golden" / "maker"
SHAPES = [(1, 1), (3, 1), (2, 5), (7, 3), (4, 16), (256, 16)]      # (threads per chunk, bytes per thread)


@pytest.fixture(scope="module")
def emu():
    SO.parent.mkdir(exist_ok=True)
    subprocess.run(["g++", "-O1", "-std=c++17", "-Wall", "-Wno-unknown-pragmas", "-fPIC", "-shared",
                    f"-I{ROOT / 'genometester4_b200' / 'csrc'}", "-o", str(SO), str(SRC)], check=True)
    return C.CDLL(str(SO))


def run_emu(emu, text: bytes, k: int, shape):
    out = np.zeros(max(1, len(text)), dtype=np.uint64)
    n = C.c_uint64()
    rc = emu.emu_sequence_words(text, C.c_uint64(len(text)), C.c_uint(k), shape[0], shape[1], C.c_void_p(out.ctypes.data), C.byref(n))
    return rc, out[:n.value].copy()


def expect(oracle, text, k):
    try:
        return oracle.sequence_words(text, k)
    except ValueError:
        return None


def test_golden_inputs(emu, oracle):
    gold = json.loads((GOLD_DIR / "maker_golden.json").read_text())
    for name in sorted({c["input"] for c in gold["cases"]}):
        text = (GOLD_DIR / name).read_bytes()
        for k in (1, 5, 16, 32):
            want = oracle.sequence_words(text, k)
            for shape in SHAPES[2:]:
                rc, got = run_emu(emu, text, k, shape)
                assert rc == 0 and np.array_equal(got, want), (name, k, shape)


def test_fuzz(emu, oracle):
    rng = np.random.default_rng(577)
    alphabet = np.frombuffer(b"ACGTacgtNn>@+\n\n\n\r \x00IJ-", dtype=np.uint8)
    n_ok = n_err = 0
    for case in range(2500):
        body = rng.choice(alphabet, size=int(rng.integers(0, 160))).tobytes()
        text = (b">" if case % 2 else b"@") + body
        k = int(rng.integers(1, 9))
        want = expect(oracle, text, k)
        for shape in (SHAPES[case % len(SHAPES)], SHAPES[(case // 7) % len(SHAPES)]):
            rc, got = run_emu(emu, text, k, shape)
            assert (rc != 0) == (want is None), (text, k, shape)
            if want is not None:
                assert np.array_equal(got, want), (text, k, shape)
        n_ok += want is not None
        n_err += want is None
    assert n_ok > 500 and n_err > 300
