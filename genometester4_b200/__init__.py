"""genometester4_b200 -- B200-native (sm_100a) engine for the sorted-merge set operations of
GenomeTester4 k-mer lists (the glistcompare hot path).

The product is ``libgt4gpu.so`` (C ABI in ``include/gt4gpu.h``, CUDA kernels in ``csrc/``) and the
``gt4gpu-compare`` CLI; this package is the thin host-side mirror used by the tests, the bench
and the multi-GPU driver.  There is no CPU fallback: importing works anywhere, computing needs
the built library and a CUDA device.
"""
from .api import (  # noqa: F401
    GT4GPUError, Result, WordList, compare_wordmaps, gt4_is_union, gt4_union, gt4_write_union,
    count_words, fasta_words_device, init, intersect_multi, last_timing, lookup, lookup_device, plan_splitters, sequence_words, set_option, set_stream, set_tile,
    union_multi,
)
from .api import RULE_ADD, RULE_DEFAULT, RULE_FIRST, RULE_MAX, RULE_MIN, RULE_NUMBER, RULE_SECOND, RULE_SUBTRACT  # noqa: F401
from ._lib import build, lib_path  # noqa: F401

__version__ = "0.1.0"
