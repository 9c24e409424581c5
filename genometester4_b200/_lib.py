"""Loader of the in-tree ``libgt4gpu.so`` (ctypes).  Fails loudly when the library is missing:
there is deliberately no Python or CPU implementation to fall back to."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"


def lib_path() -> Path:
    import os
    return Path(os.environ["GT4GPU_LIB"]) if os.environ.get("GT4GPU_LIB") else PKG / "libgt4gpu.so"


def cli_path() -> Path:
    return PKG / "gt4gpu-compare"


def listmaker_cli_path() -> Path:
    return PKG / "gt4gpu-listmaker"


def query_cli_path() -> Path:
    return PKG / "gt4gpu-query"


def build(verbose: bool = False) -> Path:
    """Compile the CUDA library and the CLI for sm_100a, in-tree (nvcc cross-compiles without a GPU)."""
    subprocess.run(["make", "-C", str(CSRC), "all"], check=True,
                   stdout=None if verbose else subprocess.DEVNULL)
    return lib_path()


class Header(C.Structure):
    """gt4gpu_header == GT4ListHeader (/root/reference/src/word-list.h:61-72)."""
    _fields_ = [("code", C.c_uint32), ("version_major", C.c_uint32), ("version_minor", C.c_uint32),
                ("word_length", C.c_uint32), ("n_words", C.c_uint64), ("total_count", C.c_uint64),
                ("list_start", C.c_uint64), ("word_bytes", C.c_uint32), ("count_bytes", C.c_uint32)]


class CResult(C.Structure):
    _fields_ = [("n_words", C.c_uint64), ("total_count", C.c_uint64), ("words", C.c_void_p),
                ("counts", C.c_void_p), ("capacity", C.c_uint64), ("word_length", C.c_uint32),
                ("flags", C.c_uint32)]


# every symbol include/gt4gpu.h declares: name -> (restype, argtypes)
_P, _U64, _U32, _I = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
SIGNATURES = {
    "gt4gpu_header_init": (None, [C.POINTER(Header), _U32]),
    "gt4gpu_init": (_I, [_I]),
    "gt4gpu_shutdown": (None, []),
    "gt4gpu_set_stream": (_I, [_P]),
    "gt4gpu_last_error": (C.c_char_p, []),
    "gt4gpu_set_tile": (_I, [_I, _I]),
    "gt4gpu_set_option": (_I, [C.c_char_p, _I]),
    "gt4gpu_last_timing": (_I, [C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(_U32)]),
    "gt4gpu_list_open": (_I, [C.c_char_p, _I, C.POINTER(_P)]),
    "gt4gpu_list_open_range": (_I, [C.c_char_p, _I, _U64, _U64, C.POINTER(_P)]),
    "gt4gpu_list_read_header": (_I, [C.c_char_p, _I, C.POINTER(Header)]),
    "gt4gpu_list_from_host_aos": (_I, [_P, _U64, _U32, C.POINTER(_P)]),
    "gt4gpu_list_from_host_soa": (_I, [_P, _P, _U64, _U32, C.POINTER(_P)]),
    "gt4gpu_list_from_device": (_I, [_P, _P, _U64, _U32, C.POINTER(_P)]),
    "gt4gpu_list_close": (None, [_P]),
    "gt4gpu_list_n_words": (_U64, [_P]),
    "gt4gpu_list_word_length": (_U32, [_P]),
    "gt4gpu_list_sum_counts": (_U64, [_P]),
    "gt4gpu_list_device_words": (_P, [_P]),
    "gt4gpu_list_device_counts": (_P, [_P]),
    "gt4gpu_compare2": (_I, [_P, _P, _U32, _I, _U32, _U32, _I, _I, C.POINTER(CResult)]),
    "gt4gpu_union_multi": (_I, [C.POINTER(_P), C.c_uint, _U32, _I, _U32, _I, C.POINTER(CResult)]),
    "gt4gpu_intersect_multi": (_I, [C.POINTER(_P), C.c_uint, _U32, _I, _U32, _I, C.POINTER(CResult)]),
    "gt4gpu_write_union": (_I, [C.POINTER(_P), C.c_uint, _U32, _I, C.POINTER(Header)]),
    "gt4gpu_union_matrix": (_I, [C.POINTER(_P), C.c_uint, _I, _P, _P, _U64, C.POINTER(_U64)]),
    "gt4gpu_list_to_host_soa": (_I, [_P, _P, _P]),
    "gt4gpu_lookup": (_I, [_P, _P, _U64, _I, _I, _P, _P]),
    "gt4gpu_sequence_words": (_I, [_P, _U64, _U32, _P, _U64, C.POINTER(_U64)]),
    "gt4gpu_fasta_words_device": (_I, [_P, _U64, _U32, C.POINTER(_P), C.POINTER(_U64)]),
    "gt4gpu_device_free": (None, [_P]),
    "gt4gpu_count_words": (_I, [_P, _U64, _I, _U32, C.POINTER(CResult)]),
    "gt4gpu_result_to_host_soa": (_I, [C.POINTER(CResult), _P, _P]),
    "gt4gpu_result_to_host_aos": (_I, [C.POINTER(CResult), _P]),
    "gt4gpu_write_list": (_I, [C.POINTER(CResult), _I]),
    "gt4gpu_write_records_at": (_I, [C.POINTER(CResult), _I, _U64]),
    "gt4gpu_result_free": (None, [C.POINTER(CResult)]),
    "gt4gpu_compare2_host_aos": (_I, [_P, _U64, _P, _U64, _U32, _U32, _I, _U32, _U32, _I, _I,
                                      C.POINTER(_P), C.POINTER(_U64), C.POINTER(_U64), C.POINTER(_U64)]),
    "gt4gpu_device_count": (_I, []),
    "gt4gpu_compare2_files": (_I, [C.c_char_p, C.c_char_p, _I, _U32, _I, _U32, _U32, _I, _I, C.POINTER(_I),
                                   C.POINTER(_U64), C.POINTER(_U64), C.POINTER(_U32)]),
    "gt4gpu_plan_splitters": (_I, [C.POINTER(_P), C.POINTER(C.c_size_t), C.POINTER(_U64), C.c_uint, C.c_uint,
                                   C.POINTER(_U64), C.POINTER(_U64)]),
    "gt4gpu_deinterleave": (_I, [_P, _U64, _P, _P]),
    "gt4gpu_interleave": (_I, [_P, _P, _U64, _P]),
}

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        p = lib_path()
        if not p.exists():
            raise RuntimeError(
                f"{p} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C genometester4_b200/csrc`). genometester4_b200 has no CPU fallback.")
        lib = C.CDLL(str(p))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError here == ABI drift, on purpose
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
