"""Shared helpers for the test-suite: deterministic small k-mer lists (numpy)."""
from __future__ import annotations

import numpy as np


def rand_list(rng: np.random.Generator, n: int, k: int, universe: np.ndarray | None = None,
              count_kind: str = "tail"):
    """n strictly ascending u64 keys < 4**k with u32 counts >= 1."""
    if universe is not None:
        words = np.sort(rng.choice(universe, size=n, replace=False)).astype(np.uint64)
    else:
        hi = (1 << (2 * k)) - 1 if k < 32 else (1 << 64) - 1
        words = np.unique(rng.integers(0, hi, size=int(n * 1.3) + 8, dtype=np.uint64, endpoint=True))
        while words.size < n:
            more = rng.integers(0, hi, size=n, dtype=np.uint64, endpoint=True)
            words = np.unique(np.concatenate([words, more]))
        words = np.sort(rng.choice(words, size=n, replace=False))
    counts = make_counts(rng, n, count_kind)
    return words, counts


def make_counts(rng, n, kind="tail"):
    if kind == "tail":          # 1..64 with a heavy tail so cutoffs bite
        c = 1 + rng.integers(0, 64, size=n, dtype=np.uint32)
        heavy = rng.random(n) < (1 / 64)
        c = np.where(heavy, c * 1000, c).astype(np.uint32)
    elif kind == "small":       # many equal counts (exercises -du)
        c = 1 + rng.integers(0, 4, size=n, dtype=np.uint32)
    elif kind == "huge":        # near 2**32 so ADD wraps
        c = (np.uint32(0xFFFFFFFF) - rng.integers(0, 3, size=n, dtype=np.uint32)).astype(np.uint32)
    else:
        raise ValueError(kind)
    return c


def make_pair(seed: int, n_a: int, n_b: int, n_both: int, k: int, count_kind="tail"):
    """Two lists sharing exactly n_both keys."""
    rng = np.random.default_rng(seed)
    total = n_a + n_b - n_both
    uni, _ = rand_list(rng, total, k)
    perm = rng.permutation(total)
    both = perm[:n_both]
    only_a = perm[n_both:n_a]
    only_b = perm[n_a:]
    ia = np.sort(np.concatenate([both, only_a]))
    ib = np.sort(np.concatenate([both, only_b]))
    wa, wb = uni[ia], uni[ib]
    ca, cb = make_counts(rng, wa.size, count_kind), make_counts(rng, wb.size, count_kind)
    return (wa, ca), (wb, cb)


def make_multi(seed: int, n_lists: int, n_each: int, universe_size: int, k: int, count_kind="tail"):
    rng = np.random.default_rng(seed)
    uni, _ = rand_list(rng, universe_size, k)
    out = []
    for _ in range(n_lists):
        n = min(n_each, universe_size)
        idx = np.sort(rng.choice(universe_size, size=n, replace=False))
        out.append((uni[idx], make_counts(rng, n, count_kind)))
    return out
