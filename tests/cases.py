"""The flag / input matrix shared by the oracle-vs-reference differential test,
the golden fixtures and the GPU parity tests.

Input sets are tiny (hundreds of records) and deterministic; they cover the
edge vectors listed in SURVEY.md section 8(c): disjoint / identical / empty
lists, key 0 and key 2**64-1, counts that wrap a u32 under ADD, cut-offs 0, 1
and above every count, every rule x every operation, -du, -dd and fused
multi-output passes, and N-list runs with empty members.
"""
from __future__ import annotations

import itertools

import numpy as np

from tests.util import make_counts, make_multi, make_pair

U64MAX = np.uint64(0xFFFFFFFFFFFFFFFF)


def _arr(words, counts):
    return np.asarray(words, dtype=np.uint64), np.asarray(counts, dtype=np.uint32)


def pair_inputs():
    """name -> (k, (wordsA, countsA), (wordsB, countsB))"""
    sets = {}
    a, b = make_pair(101, 300, 250, 100, 16, "tail")
    sets["p_tail"] = (16, a, b)
    a, b = make_pair(102, 200, 220, 150, 25, "small")
    sets["p_small"] = (25, a, b)
    a, b = make_pair(103, 120, 130, 90, 20, "huge")
    sets["p_huge"] = (20, a, b)
    a, b = make_pair(104, 150, 170, 0, 16, "tail")
    sets["p_disjoint"] = (16, a, b)
    a, _ = make_pair(105, 180, 180, 180, 16, "small")
    sets["p_identical"] = (16, a, (a[0].copy(), a[1].copy()))
    a, b = make_pair(106, 90, 70, 30, 16, "tail")
    empty = _arr([], [])
    sets["p_a_empty"] = (16, empty, b)
    sets["p_b_empty"] = (16, a, empty)
    sets["p_both_empty"] = (16, empty, empty)
    # key 0 and key 2**64-1 in both lists (k = 32), plus neighbours
    wa = np.array([0, 1, 5, 1 << 40, (1 << 64) - 2, (1 << 64) - 1], dtype=np.uint64)
    wb = np.array([0, 2, 5, 1 << 41, (1 << 64) - 1], dtype=np.uint64)
    sets["p_extremes"] = (32, _arr(wa, [3, 1, 7, 2, 9, 4]), _arr(wb, [5, 1, 7, 8, 6]))
    # ADD wraps to exactly 0 for one shared key, to 1 for another
    wa = np.array([10, 20, 30, 40], dtype=np.uint64)
    wb = np.array([10, 20, 35, 40], dtype=np.uint64)
    sets["p_wrap0"] = (16, _arr(wa, [0xFFFFFFFF, 0x80000000, 5, 2]), _arr(wb, [1, 0x80000001, 6, 0xFFFFFFFF]))
    # one element each
    sets["p_single_eq"] = (16, _arr([7], [2]), _arr([7], [2]))
    sets["p_single_ne"] = (16, _arr([7], [2]), _arr([9], [3]))
    return sets


PAIR_OPS = [("-u",), ("-i",), ("-d",), ("-dd",), ("-du",), ("-u", "-i", "-d"), ("-u", "-i", "-dd"), ("-i", "-du")]
RULES = ["default", "add", "subtract", "min", "max", "first", "second", "2"]
CUTOFFS = [0, 1, 2, 5, 100000]


def pair_rule_ok(ops, rule) -> bool:
    """main()'s validation, /root/reference/src/glistcompare.c:344-352."""
    has_i = "-i" in ops
    has_d = any(o in ops for o in ("-d", "-dd", "-du"))
    if rule in ("min", "first", "second") and not has_i:
        return False
    if rule == "subtract" and not (has_i or has_d):
        return False
    return True


def pair_cases(full: bool):
    """Yield (input_name, ops, rule, cutoff).  full=False -> the golden subset."""
    names = list(pair_inputs().keys())
    for name in names:
        rich = name in ("p_tail", "p_small", "p_huge")
        for ops, rule, cutoff in itertools.product(PAIR_OPS, RULES, CUTOFFS):
            if not pair_rule_ok(ops, rule):
                continue
            if not full:
                if rich:
                    if cutoff not in (1, 2) and not (rule == "default" and cutoff in (0, 5)):
                        continue
                    if name != "p_tail" and ops in (("-u", "-i", "-dd"), ("-i", "-du")):
                        continue
                else:
                    if rule not in ("default", "add") or cutoff not in (0, 1, 5):
                        continue
                    if ops not in (("-u",), ("-i",), ("-dd",), ("-du",), ("-u", "-i", "-d")):
                        continue
            else:
                if not rich and (rule not in ("default", "add", "max", "2") or cutoff == 100000):
                    continue
            yield name, ops, rule, cutoff


def multi_inputs():
    """name -> (k, [(words, counts), ...])"""
    sets = {}
    sets["m4_tail"] = (16, make_multi(201, 4, 150, 260, 16, "tail"))
    sets["m3_small"] = (25, make_multi(202, 3, 200, 240, 25, "small"))
    sets["m8_tail"] = (32, make_multi(203, 8, 120, 300, 32, "tail"))
    sets["m4_huge"] = (20, make_multi(204, 4, 100, 130, 20, "huge"))
    lists = make_multi(205, 4, 90, 140, 16, "tail")
    lists[1] = _arr([], [])
    sets["m4_one_empty"] = (16, lists)
    lists = make_multi(206, 3, 60, 90, 16, "tail")
    lists[0] = _arr([], [])
    sets["m3_first_empty"] = (16, lists)
    sets["m3_all_empty"] = (16, [_arr([], []), _arr([], []), _arr([], [])])
    w0 = np.array([0, 3, 9, (1 << 64) - 1], dtype=np.uint64)
    w1 = np.array([0, 3, 10, (1 << 64) - 1], dtype=np.uint64)
    w2 = np.array([0, 4, 9, (1 << 64) - 1], dtype=np.uint64)
    sets["m3_extremes"] = (32, [_arr(w0, [1, 2, 3, 4]), _arr(w1, [5, 6, 7, 8]), _arr(w2, [9, 1, 2, 3])])
    # partial sums wrap to 0 (kept by N-list union iff cutoff == 0)
    w = np.array([11, 22, 33], dtype=np.uint64)
    sets["m3_wrap0"] = (16, [_arr(w, [0xFFFFFFFF, 0x80000000, 1]), _arr(w, [1, 0x80000000, 1]), _arr(w[:2], [0, 5])])
    return sets


MULTI_OPS = [("-u",), ("-i",), ("-u", "-i")]
MULTI_RULES = ["default", "add", "max", "min", "first", "3"]
MULTI_CUTOFFS = [0, 1, 2, 5]


def multi_rule_ok(ops, rule) -> bool:
    if rule in ("min", "first") and "-i" not in ops:      # rejected by main() before any merge (:344-347)
        return False
    return True


def multi_cases(full: bool):
    for name in multi_inputs().keys():
        for ops, rule, cutoff in itertools.product(MULTI_OPS, MULTI_RULES, MULTI_CUTOFFS):
            if not multi_rule_ok(ops, rule):
                continue
            if not full and cutoff == 5 and rule not in ("default",):
                continue
            yield name, ops, rule, cutoff


def ops_to_kwargs(ops):
    """CLI letters -> compare2 keyword arguments (mirrors main(), :152-162,:334)."""
    kw = dict(union="-u" in ops, intrsec="-i" in ops,
              diff=any(o in ops for o in ("-d", "-dd", "-du")), ddiff="-dd" in ops,
              subtract="-du" in ops)
    return kw


def rule_to_kwargs(rule: str):
    if rule[0] in "123456789":
        return dict(rule="number", count_override=int(rule))
    return dict(rule=rule, count_override=1)
