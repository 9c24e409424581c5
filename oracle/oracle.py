"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  It wraps
``oracle/_build/libgt4oracle.so`` (the C restatement in gt4_oracle.c, which
cites the reference lines it follows) and offers small numpy helpers to read
and write GT4 ``.list`` files (format: /root/reference/src/word-list.h:61-72)
plus a runner for the unmodified reference binaries in ``oracle/_ref``.

Parity status: pinned (tests/test_oracle_vs_reference.py, tests/golden/).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "_build" / "libgt4oracle.so"
REF_DIR = HERE / "_ref"

RULES = {"default": 0, "add": 1, "sum": 1, "subtract": 2, "min": 3, "max": 4,
         "first": 5, "second": 6, "number": 7}

RECORD = np.dtype([("word", "<u8"), ("count", "<u4")])  # packed: 12 bytes
assert RECORD.itemsize == 12

HEADER = np.dtype([("code", "<u4"), ("major", "<u4"), ("minor", "<u4"), ("word_length", "<u4"),
                   ("n_words", "<u8"), ("total_count", "<u8"), ("list_start", "<u8"),
                   ("word_bytes", "<u4"), ("count_bytes", "<u4")])
assert HEADER.itemsize == 48
LIST_CODE = (ord("G") << 24) | (ord("T") << 16) | (ord("4") << 8) | ord("C")


class _List(C.Structure):
    _fields_ = [("words", C.c_void_p), ("counts", C.c_void_p),
                ("n_words", C.c_uint64), ("word_length", C.c_uint32)]


class _Out(C.Structure):
    _fields_ = [("words", C.c_void_p), ("counts", C.c_void_p), ("capacity", C.c_uint64),
                ("n_words", C.c_uint64), ("total_count", C.c_uint64)]


class _Header(C.Structure):
    _fields_ = [("code", C.c_uint32), ("version_major", C.c_uint32), ("version_minor", C.c_uint32),
                ("word_length", C.c_uint32), ("n_words", C.c_uint64), ("total_count", C.c_uint64),
                ("list_start", C.c_uint64), ("word_bytes", C.c_uint32), ("count_bytes", C.c_uint32)]


_lib = None


def build(quiet: bool = True) -> None:
    """Compile the oracle library (and the reference binaries when /root/reference exists)."""
    subprocess.run(["make", "-C", str(HERE), "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None,
                   stderr=subprocess.DEVNULL if quiet else None)


def lib():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            build()
        _lib = C.CDLL(str(LIB_PATH))
        _lib.gt4o_calculate_freq.restype = C.c_uint32
        _lib.gt4o_calculate_freq.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_uint32]
    return _lib


class SList:
    """A sorted k-mer list held as two numpy arrays."""

    def __init__(self, words, counts, word_length: int):
        self.words = np.ascontiguousarray(words, dtype=np.uint64)
        self.counts = np.ascontiguousarray(counts, dtype=np.uint32)
        assert self.words.shape == self.counts.shape
        self.word_length = int(word_length)

    def __len__(self):
        return int(self.words.shape[0])

    def _c(self) -> _List:
        return _List(self.words.ctypes.data, self.counts.ctypes.data, len(self), self.word_length)

    def records(self) -> np.ndarray:
        r = np.empty(len(self), dtype=RECORD)
        r["word"] = self.words
        r["count"] = self.counts
        return r


class Result:
    def __init__(self, words, counts, n_words, total_count, word_length=0):
        self.words, self.counts = words, counts
        self.n_words, self.total_count, self.word_length = int(n_words), int(total_count), int(word_length)

    def records(self) -> np.ndarray:
        r = np.empty(self.n_words, dtype=RECORD)
        r["word"] = self.words
        r["count"] = self.counts
        return r


def _mk_out(capacity: int, count_only: bool):
    if count_only:
        return _Out(None, None, 0, 0, 0), None, None
    w = np.zeros(max(capacity, 1), dtype=np.uint64)
    c = np.zeros(max(capacity, 1), dtype=np.uint32)
    return _Out(w.ctypes.data, c.ctypes.data, capacity, 0, 0), w, c


def compare2(a: SList, b: SList, *, union=False, intrsec=False, diff=False, ddiff=False,
             subtract=False, cutoff=1, rule="default", count_override=1, count_only=False):
    """Oracle for compare_wordmaps.  Returns {"union"|"intrsec"|"diff1"|"diff2": Result}."""
    rule = RULES[rule] if isinstance(rule, str) else int(rule)
    diff = diff or ddiff                      # main(): -dd implies -d (src/glistcompare.c:334)
    caps = [len(a) + len(b), min(len(a), len(b)), len(a), len(b)]
    want = [union, intrsec, diff, ddiff]
    outs = (_Out * 4)()
    keep = []
    for k in range(4):
        o, w, c = _mk_out(caps[k] if want[k] else 0, count_only or not want[k])
        outs[k] = o
        keep.append((w, c))
    la, lb = a._c(), b._c()
    rc = lib().gt4o_compare2(C.byref(la), C.byref(lb), int(union), int(intrsec), int(diff), int(ddiff),
                             int(subtract), C.c_uint32(cutoff), rule, C.c_uint32(count_override), outs)
    assert rc == 0
    res = {}
    for k, name in enumerate(["union", "intrsec", "diff1", "diff2"]):
        if not want[k]:
            continue
        n = outs[k].n_words
        w, c = keep[k]
        res[name] = Result(None if w is None else w[:n], None if c is None else c[:n], n,
                           outs[k].total_count, a.word_length)
    return res


def _multi(fn_name: str, lists, cutoff, rule, count_override, count_only, capacity):
    rule = RULES[rule] if isinstance(rule, str) else int(rule)
    arr = (_List * len(lists))(*[l._c() for l in lists])
    o, w, c = _mk_out(capacity, count_only)
    wl = C.c_uint32(0)
    fn = getattr(lib(), fn_name)
    if fn_name == "gt4o_write_union":
        rc = fn(arr, len(lists), C.c_uint32(cutoff), C.byref(o), C.byref(wl))
    else:
        rc = fn(arr, len(lists), C.c_uint32(cutoff), rule, C.c_uint32(count_override), C.byref(o), C.byref(wl))
    if rc:
        return rc, None
    n = o.n_words
    return 0, Result(None if w is None else w[:n], None if c is None else c[:n], n, o.total_count, wl.value)


def union_multi(lists, *, cutoff=1, rule="default", count_override=1, count_only=False):
    return _multi("gt4o_union_multi", lists, cutoff, rule, count_override, count_only, sum(len(l) for l in lists))


def intersect_multi(lists, *, cutoff=1, rule="default", count_override=1, count_only=False):
    return _multi("gt4o_intersect_multi", lists, cutoff, rule, count_override, count_only, min(len(l) for l in lists))


def write_union(lists, *, cutoff=1, count_only=False):
    return _multi("gt4o_write_union", lists, cutoff, "add", 0, count_only, sum(len(l) for l in lists))


def union_matrix(lists, is_union=False):
    arr = (_List * len(lists))(*[l._c() for l in lists])
    cap = sum(len(l) for l in lists) + len(lists) + 1
    words = np.zeros(cap, dtype=np.uint64)
    counts = np.zeros((cap, len(lists)), dtype=np.uint32)
    n = C.c_uint64(0)
    fn = lib().gt4o_is_union_matrix if is_union else lib().gt4o_union_matrix
    rc = fn(arr, len(lists), words.ctypes.data_as(C.c_void_p), counts.ctypes.data_as(C.c_void_p),
            C.c_uint64(cap), C.byref(n))
    if rc:
        return rc, None, None
    return 0, words[:n.value], counts[:n.value]


def sequence_words(text: bytes, word_length: int) -> np.ndarray:
    """Canonical words of a FastA/FastQ image in file order (fasta_reader_read_nwords, src/fasta.c:88-290)."""
    n = C.c_uint64()
    buf = (C.c_ubyte * max(1, len(text))).from_buffer_copy(text or b"\0")
    out = np.empty(max(1, len(text)), dtype=np.uint64)
    rc = lib().gt4o_sequence_words(buf, C.c_uint64(len(text)), C.c_uint(word_length), C.c_void_p(out.ctypes.data),
                                   C.c_uint64(out.size), C.byref(n))
    if rc:
        raise ValueError(f"sequence reader error {rc}")
    return out[:n.value].copy()


def count_words(words, word_length: int) -> SList:
    """Sorted (word, count) list of one table of raw words (wordtable_sort + merge_tables_to_file,
    src/glistmaker.c:1080-1144)."""
    w = np.array(words, dtype=np.uint64)
    ow = np.empty(max(1, w.size), dtype=np.uint64)
    oc = np.empty(max(1, w.size), dtype=np.uint32)
    lib().gt4o_count_words.restype = C.c_uint64
    u = lib().gt4o_count_words(C.c_void_p(w.ctypes.data), C.c_uint64(w.size), C.c_void_p(ow.ctypes.data),
                               C.c_void_p(oc.ctypes.data))
    return SList(ow[:u].copy(), oc[:u].copy(), word_length)


def lookup(lst: "SList", queries, canonize: bool = True):
    """(canonical words, counts) of a batch of query words: search_one_word with no mismatches
    (src/glistquery.c:544-568) over word_map_lookup (src/word-map.c:134-163); count 0 = not in the list."""
    q = np.ascontiguousarray(queries, dtype=np.uint64)
    w = np.ascontiguousarray(lst.words, dtype=np.uint64)
    c = np.ascontiguousarray(lst.counts, dtype=np.uint32)
    ow = np.empty(max(1, q.size), dtype=np.uint64)
    oc = np.empty(max(1, q.size), dtype=np.uint32)
    lib().gt4o_lookup(C.c_void_p(w.ctypes.data), C.c_void_p(c.ctypes.data), C.c_uint64(w.size), C.c_uint(lst.word_length),
                      C.c_void_p(q.ctypes.data), C.c_uint64(q.size), C.c_int(int(canonize)), C.c_void_p(ow.ctypes.data),
                      C.c_void_p(oc.ctypes.data))
    return ow[:q.size].copy(), oc[:q.size].copy()


def word_to_string(word: int, word_length: int) -> str:
    """src/sequence.c:88-100"""
    return "".join("ACGT"[(int(word) >> (2 * (word_length - 1 - i))) & 3] for i in range(word_length))


def calculate_freq(f1, f2, rule, count_override=1):
    rule = RULES[rule] if isinstance(rule, str) else int(rule)
    return lib().gt4o_calculate_freq(f1, f2, rule, count_override)


# ---------------------------------------------------------------- list files

def header_bytes(word_length: int, n_words: int = 0, total_count: int = 0) -> bytes:
    h = _Header()
    lib().gt4o_header_init(C.byref(h), C.c_uint32(word_length))
    h.n_words, h.total_count = n_words, total_count
    return bytes(h)


def parse_header(buf: bytes, mode: int = 0):
    """Returns (rc, dict).  mode 0 = mmap rules, 1 = --stream rules."""
    h = _Header()
    rc = lib().gt4o_header_parse(buf, C.c_uint64(len(buf)), mode, C.byref(h))
    return rc, {f: getattr(h, f) for f, _ in _Header._fields_}


def write_list(path, words, counts, word_length: int, *, minor: int | None = None) -> None:
    """Write a .list file exactly as the reference does (48-byte 4.2 header + 12-byte records).

    ``minor=0`` writes the legacy 40-byte 4.0 header, ``minor=1|2`` the 40-byte
    4.2 layout (list_start field = 40) -- used to exercise the reader rules."""
    sl = SList(words, counts, word_length)
    total = int(sl.counts.astype(np.uint64).sum())
    hb = bytearray(header_bytes(word_length, len(sl), total))
    if minor is not None:
        hb[8:12] = np.uint32(minor).tobytes()
        if minor <= 2:
            hb = hb[:40]
            hb[32:40] = np.uint64(40).tobytes() if minor else np.uint64(0).tobytes()
    with open(path, "wb") as f:
        f.write(bytes(hb))
        f.write(sl.records().tobytes())


def read_list(path, mode: int = 0) -> SList:
    buf = Path(path).read_bytes()
    rc, h = parse_header(buf, mode)
    if rc:
        raise ValueError(f"{path}: header rejected (rc={rc})")
    rec = np.frombuffer(buf, dtype=RECORD, count=h["n_words"], offset=h["list_start"])
    return SList(rec["word"].copy(), rec["count"].copy(), h["word_length"])


# ------------------------------------------------------- reference binaries

def ref_binary(name: str) -> Path | None:
    p = REF_DIR / name
    if not p.exists() and Path("/root/reference/src").exists():
        build()
    return p if p.exists() else None


def run_ref(name: str, args, cwd=None, check=False, timeout=60, attempts=3):
    """Run an unmodified reference tool from oracle/_ref; returns CompletedProcess or None."""
    exe = ref_binary(name)
    if exe is None:
        return None
    # glistmaker's main thread leaves through pthread_exit (src/glistmaker.c:365) and now and then a worker stays
    # parked on the queue's condition variable for good: bound every run and try again
    for attempt in range(attempts):
        try:
            return subprocess.run([str(exe), *map(str, args)], cwd=cwd, capture_output=True, check=check,
                                  env={**os.environ, "LC_ALL": "C"}, timeout=timeout)
        except subprocess.TimeoutExpired:
            if attempt == attempts - 1:
                raise
