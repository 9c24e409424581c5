"""CPU-side checks: the C-ABI library loads and exports every symbol include/gt4gpu.h declares (no compute calls),
fails loudly without a device, header layout, and the synthetic generators' numpy / torch twins agree."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from genometester4_b200 import _lib, api, synth

ROOT = Path(__file__).resolve().parent.parent


def declared_functions():
    text = (ROOT / "include" / "gt4gpu.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gt4gpu_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    names = declared_functions()
    assert len(names) >= 30
    lib = C.CDLL(str(_lib.lib_path()))
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/gt4gpu.h but not exported by libgt4gpu.so"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in genometester4_b200/_lib.py"
    assert sorted(_lib.SIGNATURES) == names


def test_header_layout_matches_reference_struct():
    h = _lib.Header()
    _lib.load().gt4gpu_header_init(C.byref(h), 25)
    raw = bytes(h)
    assert len(raw) == 48 and raw[:4] == b"C4TG"                       # 'GT4C' as a little-endian u32 (word-list.c:31)
    assert (h.version_major, h.version_minor, h.word_length, h.list_start, h.word_bytes, h.count_bytes) == (4, 2, 25, 48, 8, 4)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device behaviour")
def test_compute_fails_loudly_without_a_device():
    with pytest.raises(api.GT4GPUError) as ei:
        api.WordList.from_arrays(np.arange(4, dtype=np.uint64), np.ones(4, np.uint32), 16)
    assert ei.value.code == 4 and "no CPU fallback" in str(ei.value)
    with pytest.raises(api.GT4GPUError):
        api.compare2_host_records(np.zeros(2, api.RECORD), np.zeros(2, api.RECORD), 16, api.OP_UNION, countonly=1)


def test_bad_arguments_are_rejected_before_touching_the_device():
    lib = _lib.load()
    assert lib.gt4gpu_set_tile(100, 3) == 1 and b"unsupported" in lib.gt4gpu_last_error()
    assert lib.gt4gpu_set_option(b"no_such_option", 1) == 1
    h = _lib.Header()
    assert lib.gt4gpu_list_read_header(b"/nonexistent/file.list", 0, C.byref(h)) == 2


@pytest.mark.parametrize("k,m", [(16, 4000), (25, 3000), (32, 5000)])
def test_synthetic_generators_numpy_and_torch_agree(k, m):
    (wa, ca), (wb, cb) = synth.pair_numpy(42, k, m, 100, m - 50, 0.3, 0.25)
    (ta, tca), (tb, tcb) = synth.pair_torch(42, k, m, 100, m - 50, 0.3, 0.25, device="cpu", chunk=777)
    for x, y in ((wa, ta), (ca, tca), (wb, tb), (cb, tcb)):
        assert np.array_equal(x, y.numpy().view(x.dtype))
    assert np.all(wa[1:] > wa[:-1]) and np.all(wb[1:] > wb[:-1]) and int(wa.max()) < 4 ** k and ca.min() >= 1
    # shards of the universe concatenate to the whole (what makes per-rank generation a key-range sharding)
    (w1, _), _ = synth.pair_numpy(42, k, m, 0, m // 2, 0.3, 0.25)
    (w2, _), _ = synth.pair_numpy(42, k, m, m // 2, m, 0.3, 0.25)
    (w, _), _ = synth.pair_numpy(42, k, m, 0, m, 0.3, 0.25)
    assert np.array_equal(np.concatenate([w1, w2]), w)
    for j in range(3):
        w, c = synth.list_numpy(5, k, m, 0, m, j, 0.4)
        tw, tc = synth.list_torch(5, k, m, 0, m, j, 0.4, device="cpu", chunk=999)
        assert np.array_equal(w, tw.numpy().view(np.uint64)) and np.array_equal(c, tc.numpy().view(np.uint32))


def test_index_header_is_readable_without_a_device():
    h = _lib.Header()
    idx = ROOT / "tests" / "golden" / "index" / "x_16.index"
    assert _lib.load().gt4gpu_list_read_header(str(idx).encode(), 0, C.byref(h)) == 0
    assert bytes(h)[:4] == b"I4TG" and (h.word_length, h.word_bytes, h.count_bytes) == (16, 8, 8) and h.n_words > 2000


def test_header_compiles_as_c99_and_cxx(tmp_path):
    """include/gt4gpu.h is the drop-in boundary: it must be consumable by a plain C99 host (the reference is C) and by C++."""
    import subprocess
    root = Path(__file__).resolve().parent.parent
    src = tmp_path / "use_header.c"
    src.write_text('#include "gt4gpu.h"\nint main (void) { gt4gpu_header h; gt4gpu_result r; (void) h; (void) r; '
                   'return sizeof (gt4gpu_header) == 48 ? 0 : 1; }\n')
    for cc, std in (("gcc", "-std=c99"), ("g++", "-std=c++17")):
        lang = ["-x", "c++"] if cc == "g++" else []
        subprocess.run([cc, std, "-Wall", "-Wextra", "-pedantic", "-Werror", f"-I{root / 'include'}", "-fsyntax-only", *lang, str(src)], check=True)
