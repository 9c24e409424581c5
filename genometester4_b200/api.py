"""Host-side mirror of the reference's operator interface for the set-operation path.

Function names, argument meaning and error behaviour follow the reference so that the parity
tests read like calls into it:

* :func:`compare_wordmaps`  -- /root/reference/src/glistcompare.c:66,789 (two lists, up to 4 outputs)
* :func:`union_multi` / :func:`intersect_multi` -- :70-71,500,605 (N lists)
* :func:`gt4_write_union`, :func:`gt4_union`, :func:`gt4_is_union` -- /root/reference/src/set-operations.h:34-39
* :class:`WordList` -- the container (GT4WordMap / GT4WordListStream, src/word-map.c:165, src/word-list-stream.c:127)

Everything computes in ``libgt4gpu.so`` on the GPU; this file only marshals arguments.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import CResult, Header

RULE_DEFAULT, RULE_ADD, RULE_SUBTRACT, RULE_MIN, RULE_MAX, RULE_FIRST, RULE_SECOND, RULE_NUMBER = range(8)
RULES = {"default": 0, "add": 1, "sum": 1, "subtract": 2, "min": 3, "max": 4, "first": 5, "second": 6, "number": 7}
OP_UNION, OP_INTRSEC, OP_DIFF, OP_DDIFF = 1, 2, 4, 8
STREAM_NAMES = ("union", "intrsec", "diff1", "diff2")
FLAG_CALLER_BUFFERS, FLAG_COUNT_ONLY = 1, 2

RECORD = np.dtype([("word", "<u8"), ("count", "<u4")])


class GT4GPUError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"gt4gpu error {code}: {message}")
        self.code = code


def _check(rc: int):
    if rc:
        raise GT4GPUError(rc, _lib.load().gt4gpu_last_error().decode(errors="replace"))


def _rule(rule) -> int:
    return RULES[rule] if isinstance(rule, str) else int(rule)


def init(device: int = -1) -> None:
    _check(_lib.load().gt4gpu_init(device))


CUDA_STREAM_LEGACY = 1      # cudaStreamLegacy: the handle that names the default stream explicitly


def set_stream(cuda_stream: int | None) -> None:
    """Launch on an external cudaStream_t, e.g. ``torch.cuda.current_stream().cuda_stream``.

    torch reports its default stream as 0; that is translated to ``cudaStreamLegacy`` so that the library's
    kernels really are ordered with torch's (a NULL handle means "use the library's own non-blocking stream"
    in the C ABI, which is NOT ordered with the default stream).  ``None`` restores the library's own stream."""
    if cuda_stream is None:
        handle = 0
    else:
        handle = cuda_stream if cuda_stream else CUDA_STREAM_LEGACY
    _check(_lib.load().gt4gpu_set_stream(C.c_void_p(handle)))


def set_tile(threads: int, items: int) -> None:
    _check(_lib.load().gt4gpu_set_tile(threads, items))


def set_option(name: str, value: int) -> None:
    """Tuning knobs of the library: "stream_consumers" (256/512), "stream_items" (7/9/11), "use_stream_kernel" (0/1)."""
    _check(_lib.load().gt4gpu_set_option(name.encode(), int(value)))


def last_timing():
    """(partition_ms, merge_ms, launches) of the most recent merge call, from CUDA events."""
    a, b, n = C.c_float(), C.c_float(), C.c_uint32()
    _lib.load().gt4gpu_last_timing(C.byref(a), C.byref(b), C.byref(n))
    return a.value, b.value, n.value


class WordList:
    """A sorted k-mer list resident in HBM (SoA u64 words / u32 counts)."""

    def __init__(self, handle: int, keepalive=None):
        self._h = C.c_void_p(handle)
        self._keep = keepalive

    # ---- constructors
    @classmethod
    def open(cls, path, stream: bool = False, first: int = 0, count: int | None = None) -> "WordList":
        h = C.c_void_p()
        lib = _lib.load()
        if count is None and first == 0:
            _check(lib.gt4gpu_list_open(os.fsencode(path), int(stream), C.byref(h)))
        else:
            _check(lib.gt4gpu_list_open_range(os.fsencode(path), int(stream), first,
                                              (1 << 64) - 1 if count is None else count, C.byref(h)))
        return cls(h.value)

    @classmethod
    def from_arrays(cls, words, counts, word_length: int) -> "WordList":
        w = np.ascontiguousarray(words, dtype=np.uint64)
        c = np.ascontiguousarray(counts, dtype=np.uint32)
        assert w.shape == c.shape and w.ndim == 1
        h = C.c_void_p()
        _check(_lib.load().gt4gpu_list_from_host_soa(C.c_void_p(w.ctypes.data), C.c_void_p(c.ctypes.data),
                                                     w.size, word_length, C.byref(h)))
        return cls(h.value)

    @classmethod
    def from_records(cls, records: np.ndarray, word_length: int) -> "WordList":
        r = np.ascontiguousarray(records, dtype=RECORD)
        h = C.c_void_p()
        _check(_lib.load().gt4gpu_list_from_host_aos(C.c_void_p(r.ctypes.data), r.size, word_length, C.byref(h)))
        return cls(h.value)

    @classmethod
    def from_device(cls, words_ptr: int, counts_ptr: int, n_words: int, word_length: int, keepalive=None) -> "WordList":
        """Wrap device arrays owned by the caller (``tensor.data_ptr()``); pass the tensors as keepalive."""
        h = C.c_void_p()
        _check(_lib.load().gt4gpu_list_from_device(C.c_void_p(words_ptr), C.c_void_p(counts_ptr), n_words,
                                                   word_length, C.byref(h)))
        return cls(h.value, keepalive)

    # ---- GT4WordSListInstance fields
    @property
    def num_words(self) -> int:
        return _lib.load().gt4gpu_list_n_words(self._h)

    @property
    def word_length(self) -> int:
        return _lib.load().gt4gpu_list_word_length(self._h)

    @property
    def sum_counts(self) -> int:
        return _lib.load().gt4gpu_list_sum_counts(self._h)

    def __len__(self):
        return self.num_words

    def close(self):
        if self._h:
            _lib.load().gt4gpu_list_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


@dataclass
class Result:
    """One output stream: header totals plus (optionally) the records, still on the device."""
    n_words: int
    total_count: int
    word_length: int
    _c: CResult | None = None

    def to_host(self):
        """(words u64[n], counts u32[n]) on the host."""
        w = np.empty(self.n_words, dtype=np.uint64)
        c = np.empty(self.n_words, dtype=np.uint32)
        if self.n_words:
            _check(_lib.load().gt4gpu_result_to_host_soa(C.byref(self._c), C.c_void_p(w.ctypes.data), C.c_void_p(c.ctypes.data)))
        return w, c

    def records(self) -> np.ndarray:
        r = np.empty(self.n_words, dtype=RECORD)
        if self.n_words:
            _check(_lib.load().gt4gpu_result_to_host_aos(C.byref(self._c), C.c_void_p(r.ctypes.data)))
        return r

    def list_bytes(self) -> bytes:
        """The exact bytes of the .list file the reference would have written."""
        h = Header()
        _lib.load().gt4gpu_header_init(C.byref(h), self.word_length)
        h.n_words, h.total_count = self.n_words, self.total_count
        return bytes(h) + self.records().tobytes()

    def write(self, fd: int) -> None:
        _check(_lib.load().gt4gpu_write_list(C.byref(self._c), fd))

    def write_records_at(self, fd: int, first_record: int) -> None:
        _check(_lib.load().gt4gpu_write_records_at(C.byref(self._c), fd, first_record))

    @property
    def device_ptrs(self):
        return self._c.words, self._c.counts

    def as_torch(self):
        """Zero-copy torch views (int64 / int32 bit patterns) of the device-resident records."""
        import torch

        class _View:
            def __init__(self, ptr, n, typestr):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}

        if self.n_words == 0:
            return torch.empty(0, dtype=torch.int64, device="cuda"), torch.empty(0, dtype=torch.int32, device="cuda")
        return (torch.as_tensor(_View(self._c.words, self.n_words, "<i8"), device="cuda"),
                torch.as_tensor(_View(self._c.counts, self.n_words, "<i4"), device="cuda"))

    def free(self):
        if self._c is not None:
            _lib.load().gt4gpu_result_free(C.byref(self._c))
            self._c = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _wrap(c: CResult) -> Result:
    own = CResult()
    C.memmove(C.byref(own), C.byref(c), C.sizeof(CResult))
    return Result(own.n_words, own.total_count, own.word_length, own)


def compare_wordmaps(list1: WordList, list2: WordList, find_union=0, find_intrsec=0, find_diff=0, find_ddiff=0,
                     subtract=0, countonly=0, cutoff: int = 1, rule=RULE_DEFAULT, count_override: int = 1,
                     out_buffers: dict | None = None) -> dict:
    """Two-list merge (compare_wordmaps, src/glistcompare.c:789-955).

    Returns ``{"union"|"intrsec"|"diff1"|"diff2": Result}`` for the requested outputs.  As in
    ``main`` (:334) ``find_ddiff`` implies ``find_diff``.  ``out_buffers`` optionally maps a stream
    name to ``(words_ptr, counts_ptr, capacity)`` device buffers owned by the caller."""
    if find_ddiff:
        find_diff = 1
    ops = (OP_UNION if find_union else 0) | (OP_INTRSEC if find_intrsec else 0) | \
          (OP_DIFF if find_diff else 0) | (OP_DDIFF if find_ddiff else 0)
    out = (CResult * 4)()
    for name, (wp, cp, cap) in (out_buffers or {}).items():
        s = STREAM_NAMES.index(name)
        out[s].words, out[s].counts, out[s].capacity, out[s].flags = wp, cp, cap, FLAG_CALLER_BUFFERS
    _check(_lib.load().gt4gpu_compare2(list1._h, list2._h, ops, _rule(rule), cutoff, count_override,
                                       int(bool(subtract)), int(bool(countonly)), out))
    return {STREAM_NAMES[s]: _wrap(out[s]) for s in range(4) if (ops >> s) & 1}


def _handles(lists):
    return (C.c_void_p * len(lists))(*[l._h for l in lists])


def union_multi(lists, cutoff: int = 1, rule=RULE_DEFAULT, count_override: int = 1, countonly=0) -> Result:
    """N-list union (union_multi, src/glistcompare.c:500-603).  Raises GT4GPUError(code 1) for a
    rule outside {default, add, max, number}, where the reference returns 1."""
    out = CResult()
    _check(_lib.load().gt4gpu_union_multi(_handles(lists), len(lists), cutoff, _rule(rule), count_override,
                                          int(bool(countonly)), C.byref(out)))
    return _wrap(out)


def intersect_multi(lists, cutoff: int = 1, rule=RULE_DEFAULT, count_override: int = 1, countonly=0) -> Result:
    """N-list intersection (intersect_multi, src/glistcompare.c:605-717)."""
    out = CResult()
    _check(_lib.load().gt4gpu_intersect_multi(_handles(lists), len(lists), cutoff, _rule(rule), count_override,
                                              int(bool(countonly)), C.byref(out)))
    return _wrap(out)


def gt4_write_union(arrays, cutoff: int, ofile: int = 0) -> Header:
    """gt4_write_union (src/set-operations.c:41-129): writes to fd ``ofile`` unless it is 0; returns the header."""
    h = Header()
    _check(_lib.load().gt4gpu_write_union(_handles(arrays), len(arrays), cutoff, ofile, C.byref(h)))
    return h


def lookup(lst: WordList, queries, canonize: bool = True):
    """Batch form of glistquery's exact lookups (search_one_word, src/glistquery.c:544-568, over word_map_lookup,
    src/word-map.c:134-163).  Returns (canonical words, counts) as numpy arrays; count 0 = the list does not hold it."""
    q = np.ascontiguousarray(queries, dtype=np.uint64)
    counts = np.zeros(q.size, dtype=np.uint32)
    canon = np.zeros(q.size, dtype=np.uint64)
    _check(_lib.load().gt4gpu_lookup(lst._h, C.c_void_p(q.ctypes.data), q.size, 0, int(bool(canonize)),
                                     C.c_void_p(counts.ctypes.data), C.c_void_p(canon.ctypes.data)))
    return canon, counts


def lookup_device(lst: WordList, queries_ptr: int, n_queries: int, counts_ptr: int, canonical_ptr: int = 0, canonize: bool = True) -> None:
    """Same on device arrays (``tensor.data_ptr()``)."""
    _check(_lib.load().gt4gpu_lookup(lst._h, C.c_void_p(queries_ptr), n_queries, 1, int(bool(canonize)),
                                     C.c_void_p(counts_ptr), C.c_void_p(canonical_ptr) if canonical_ptr else None))


def sequence_words(text: bytes, word_length: int) -> np.ndarray:
    """Canonical words of a FastA/FastQ image in file order (fasta_reader_read_nwords, src/fasta.c:88-290).  Host only.
    Raises GT4GPUError(code 3) where the reference's reader reports a format error."""
    n = C.c_uint64()
    out = np.empty(max(1, len(text)), dtype=np.uint64)
    _check(_lib.load().gt4gpu_sequence_words(text, len(text), word_length, C.c_void_p(out.ctypes.data), out.size, C.byref(n)))
    return out[:n.value].copy()


class DeviceWords:
    """Canonical words of a FastA image, resident in HBM (gt4gpu_fasta_words_device)."""

    def __init__(self, ptr: int, n: int, word_length: int):
        self.ptr, self.n_words, self.word_length = ptr, n, word_length

    def to_host(self) -> np.ndarray:
        import torch

        class _View:
            def __init__(self, ptr, n):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 2}
        if not self.n_words:
            return np.zeros(0, dtype=np.uint64)
        return torch.as_tensor(_View(self.ptr, self.n_words), device="cuda").cpu().numpy().astype(np.uint64)

    def free(self):
        if self.ptr:
            _lib.load().gt4gpu_device_free(C.c_void_p(self.ptr))
            self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def fasta_words_device(text: bytes, word_length: int) -> DeviceWords:
    """FastA or FastQ image (host bytes) -> canonical words in HBM, parsed on the GPU (fasta_reader_read_nwords,
    src/fasta.c:88-290).  Feed ``.ptr`` / ``.n_words`` to :func:`count_words`.  A FastQ image the reference's reader
    would give up on raises GT4GPUError(code 3); :func:`sequence_words` reads such images up to that point."""
    ptr, n = C.c_void_p(), C.c_uint64()
    _check(_lib.load().gt4gpu_fasta_words_device(text, len(text), word_length, C.byref(ptr), C.byref(n)))
    return DeviceWords(ptr.value or 0, n.value, word_length)


def count_words(words, word_length: int, n_words: int | None = None) -> Result:
    """Back end of glistmaker for one table of raw words: sort (wordtable_sort, src/word-table.c) and count the
    occurrences of every distinct word (merge_tables_to_file, src/glistmaker.c:1080-1144).  ``words`` is a numpy
    uint64 array (host) or an int device pointer with ``n_words``; returns the sorted (word, count) list."""
    out = CResult()
    if isinstance(words, int):
        _check(_lib.load().gt4gpu_count_words(C.c_void_p(words), n_words, 1, word_length, C.byref(out)))
    else:
        w = np.ascontiguousarray(words, dtype=np.uint64)
        _check(_lib.load().gt4gpu_count_words(C.c_void_p(w.ctypes.data), w.size, 0, word_length, C.byref(out)))
    return _wrap(out)


def _matrix(lists, is_union: bool):
    lib = _lib.load()
    n = C.c_uint64()
    cap = sum(len(l) for l in lists) + len(lists) + 1
    words = np.zeros(cap, dtype=np.uint64)
    counts = np.zeros((cap, len(lists)), dtype=np.uint32)
    _check(lib.gt4gpu_union_matrix(_handles(lists), len(lists), int(is_union), C.c_void_p(words.ctypes.data),
                                   C.c_void_p(counts.ctypes.data), cap, C.byref(n)))
    return words[:n.value], counts[:n.value]


def gt4_union(objs, callback=None, data=None):
    """gt4_union (src/set-operations.c:132-183).  Without a callback returns (words, counts[n, n_objs]);
    with one, calls ``callback(word, counts_row, data)`` per row and stops on a non-zero return."""
    words, counts = _matrix(objs, False)
    if callback is None:
        return words, counts
    for w, row in zip(words, counts):
        r = callback(int(w), row, data)
        if r:
            return r
    return 0


def gt4_is_union(objs, callback=None, data=None):
    """gt4_is_union (src/set-operations.c:186-228): rows restricted to the words of list 0."""
    words, counts = _matrix(objs, True)
    if callback is None:
        return words, counts
    for w, row in zip(words, counts):
        r = callback(int(w), row, data)
        if r:
            return r
    return 0


def compare2_host_records(rec_a: np.ndarray, rec_b: np.ndarray, word_length: int, ops: int, rule=RULE_DEFAULT,
                          cutoff: int = 1, count_override: int = 1, subtract=0, countonly=0, out_records=None):
    """End-to-end host path (gt4gpu_compare2_host_aos): packed records in, packed records out.

    rec_a / rec_b / out_records[k] may be numpy arrays or (ptr, n) tuples naming pinned memory."""
    def ptr_n(x):
        if isinstance(x, tuple):
            return x
        return x.ctypes.data, x.size
    pa, na = ptr_n(rec_a)
    pb, nb = ptr_n(rec_b)
    outs = (C.c_void_p * 4)()
    caps = (C.c_uint64 * 4)()
    for s in range(4):
        if out_records and out_records[s] is not None:
            outs[s], caps[s] = ptr_n(out_records[s])
    n_out = (C.c_uint64 * 4)()
    t_out = (C.c_uint64 * 4)()
    _check(_lib.load().gt4gpu_compare2_host_aos(C.c_void_p(pa), na, C.c_void_p(pb), nb, word_length, ops, _rule(rule),
                                                cutoff, count_override, int(bool(subtract)), int(bool(countonly)),
                                                outs, caps, n_out, t_out))
    return list(n_out), list(t_out)


def compare_files(path_a, path_b, out_prefix: str = "out", find_union=0, find_intrsec=0, find_diff=0, find_ddiff=0, subtract=0, countonly=0,
                  cutoff: int = 1, rule=RULE_DEFAULT, count_override: int = 1, stream: bool = False, mode: int = 0o666) -> dict:
    """`glistcompare A B ...` file to file through the pipelined path (gt4gpu_compare2_files): same output names as the
    reference (`<out>_<k>_union.list`, ..., written through `.tmp` and renamed).  Returns {stream: (n_words, total_count)}."""
    import os
    if find_ddiff:
        find_diff = 1
    ops = (OP_UNION if find_union else 0) | (OP_INTRSEC if find_intrsec else 0) | (OP_DIFF if find_diff else 0) | (OP_DDIFF if find_ddiff else 0)
    h = _lib.Header()
    _check(_lib.load().gt4gpu_list_read_header(os.fsencode(str(path_a)), int(stream), C.byref(h)))
    k = h.word_length
    tags = {"union": "union", "intrsec": "intrsec", "diff1": "0_diff1", "diff2": "0_diff2"}
    fds = (C.c_int * 4)(-1, -1, -1, -1)
    names = {}
    try:
        for s in range(4):
            if (ops >> s) & 1 and not countonly:
                final = f"{out_prefix}_{k}_{tags[STREAM_NAMES[s]]}.list"
                names[s] = (final + ".tmp", final)
                fds[s] = os.open(names[s][0], os.O_WRONLY | os.O_CREAT | os.O_TRUNC, mode)
        n_out, tot = (C.c_uint64 * 4)(), (C.c_uint64 * 4)()
        wl = C.c_uint32(0)
        _check(_lib.load().gt4gpu_compare2_files(os.fsencode(str(path_a)), os.fsencode(str(path_b)), int(stream), ops, _rule(rule), cutoff,
                                                 count_override, int(bool(subtract)), int(bool(countonly)), fds, n_out, tot, C.byref(wl)))
    finally:
        for s in range(4):
            if fds[s] >= 0:
                os.close(fds[s])
    for tmp, final in names.values():
        os.rename(tmp, final)
    return {STREAM_NAMES[s]: (int(n_out[s]), int(tot[s])) for s in range(4) if (ops >> s) & 1}


def plan_splitters(key_arrays, n_parts: int):
    """Key-range sharding plan (gt4gpu_plan_splitters).  key_arrays: list of 1-D numpy arrays, either
    u64 keys (stride 8) or packed RECORD arrays (stride 12).  Returns (bounds[n_lists, n_parts+1], splitters)."""
    n = len(key_arrays)
    ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in key_arrays])
    strides = (C.c_size_t * n)(*[a.strides[0] if a.size else a.dtype.itemsize for a in key_arrays])
    sizes = (C.c_uint64 * n)(*[a.size for a in key_arrays])
    bounds = (C.c_uint64 * (n * (n_parts + 1)))()
    split = (C.c_uint64 * max(n_parts - 1, 1))()
    _check(_lib.load().gt4gpu_plan_splitters(ptrs, strides, sizes, n, n_parts, bounds, split))
    return np.array(bounds, dtype=np.uint64).reshape(n, n_parts + 1), np.array(split[:n_parts - 1], dtype=np.uint64)
