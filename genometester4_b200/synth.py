"""Deterministic synthetic k-mer lists, stateless per universe index so
that any key range -- one GPU's shard -- can be generated on its own, on the device, at 1e9 scale.

Universe element ``u`` (0 <= u < M) has key ``u * G + (h1(u) >> 1) % G`` with ``G = (4**k - 1) // M``
(strictly increasing in u, < 4**k), belongs to A only / B only / both according to a 24-bit hash
against the (p_a_only, p_b_only, p_both) thresholds, and carries count ``1 + (h & 63)``, times 1000
for 1/64 of the entries (a heavy tail so that cut-offs bite).  ``h*`` = splitmix64 of u xor a salt.

Two implementations with identical output: numpy (host; tests and the CPU reference arm) and torch
(any device; the bench).  torch has no unsigned 64-bit arithmetic, so the torch version works on
the int64 bit patterns and masks after right shifts.
"""
from __future__ import annotations

import numpy as np

GOLDEN = 0x9E3779B97F4A7C15
M1 = 0xBF58476D1CE4E5B9
M2 = 0x94D049BB133111EB
SALT_KEY, SALT_MEMBER, SALT_CA, SALT_CB = 0x0, 0xA5A5A5A5A5A5A5A5, 0x5DEECE66D1234567, 0x1F83D9ABFB41BD6B


def gap(k: int, universe: int) -> int:
    return ((1 << (2 * k)) - 1) // universe


# ------------------------------------------------------------------ numpy

def _mix_np(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = x + np.uint64(GOLDEN)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(M1)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(M2)
        return z ^ (z >> np.uint64(31))


def _counts_np(h: np.ndarray) -> np.ndarray:
    c = (np.uint64(1) + (h & np.uint64(63))).astype(np.uint32)
    heavy = ((h >> np.uint64(6)) & np.uint64(63)) == 0
    return np.where(heavy, c * np.uint32(1000), c).astype(np.uint32)


def pair_numpy(seed: int, k: int, universe: int, u0: int, u1: int, p_a_only: float, p_b_only: float):
    """Lists A and B restricted to universe indices [u0, u1).  Returns ((wa, ca), (wb, cb))."""
    g = gap(k, universe)
    u = np.arange(u0, u1, dtype=np.uint64)
    s = np.uint64(seed & 0xFFFFFFFFFFFFFFFF)
    with np.errstate(over="ignore"):
        key = u * np.uint64(g) + (_mix_np(u ^ s ^ np.uint64(SALT_KEY)) >> np.uint64(1)) % np.uint64(g)
    r = (_mix_np(u ^ s ^ np.uint64(SALT_MEMBER)) >> np.uint64(40)).astype(np.int64)
    ta = int(p_a_only * (1 << 24))
    tb = ta + int(p_b_only * (1 << 24))
    in_a = (r < ta) | (r >= tb)
    in_b = r >= ta
    ca = _counts_np(_mix_np(u ^ s ^ np.uint64(SALT_CA)))
    cb = _counts_np(_mix_np(u ^ s ^ np.uint64(SALT_CB)))
    return (key[in_a], ca[in_a]), (key[in_b], cb[in_b])


# ------------------------------------------------------------------ torch

def _s64(x: int) -> int:
    x &= 0xFFFFFFFFFFFFFFFF
    return x - (1 << 64) if x >= (1 << 63) else x


def _lsr(z, n: int):
    return (z >> n) & ((1 << (64 - n)) - 1)


def _mix_t(x):
    z = x + _s64(GOLDEN)
    z = (z ^ _lsr(z, 30)) * _s64(M1)
    z = (z ^ _lsr(z, 27)) * _s64(M2)
    return z ^ _lsr(z, 31)


def _counts_t(h):
    import torch
    c = (1 + (h & 63)).to(torch.int32)
    heavy = ((h >> 6) & 63) == 0
    return torch.where(heavy, c * 1000, c)


def pair_torch(seed: int, k: int, universe: int, u0: int, u1: int, p_a_only: float, p_b_only: float,
               device="cuda", chunk: int = 1 << 26):
    """Same lists as :func:`pair_numpy`, built on ``device``.  Keys are int64 tensors holding the u64 bit
    patterns, counts int32 tensors holding the u32 bit patterns."""
    import torch
    g = gap(k, universe)
    s = _s64(seed)
    ta = int(p_a_only * (1 << 24))
    tb = ta + int(p_b_only * (1 << 24))
    parts = {"wa": [], "ca": [], "wb": [], "cb": []}
    for c0 in range(u0, u1, chunk):
        c1 = min(c0 + chunk, u1)
        u = torch.arange(c0, c1, dtype=torch.int64, device=device)
        key = u * g + torch.remainder(_lsr(_mix_t(u ^ s ^ _s64(SALT_KEY)), 1), g)
        r = _lsr(_mix_t(u ^ s ^ _s64(SALT_MEMBER)), 40)
        in_a = (r < ta) | (r >= tb)
        in_b = r >= ta
        ca = _counts_t(_mix_t(u ^ s ^ _s64(SALT_CA)))
        cb = _counts_t(_mix_t(u ^ s ^ _s64(SALT_CB)))
        parts["wa"].append(key[in_a]); parts["ca"].append(ca[in_a])
        parts["wb"].append(key[in_b]); parts["cb"].append(cb[in_b])
        del u, key, r, in_a, in_b, ca, cb
    out = {name: (torch.cat(v) if len(v) != 1 else v[0]) for name, v in parts.items()}
    return (out["wa"], out["ca"]), (out["wb"], out["cb"])


def to_numpy_u(words_t, counts_t):
    """torch (int64, int32) bit patterns -> numpy (uint64, uint32)."""
    return words_t.cpu().numpy().view(np.uint64), counts_t.cpu().numpy().view(np.uint32)


# ------------------------------------------------------------------ N lists over one shared universe (config 5)

def list_numpy(seed: int, k: int, universe: int, u0: int, u1: int, list_id: int, p_member: float):
    """List `list_id` of a family drawn from ONE universe (same keys, independent membership and counts)."""
    g = gap(k, universe)
    u = np.arange(u0, u1, dtype=np.uint64)
    s = np.uint64(seed & 0xFFFFFFFFFFFFFFFF)
    salt = np.uint64((0xD1B54A32D192ED03 * (list_id + 1)) & 0xFFFFFFFFFFFFFFFF)
    with np.errstate(over="ignore"):
        key = u * np.uint64(g) + (_mix_np(u ^ s ^ np.uint64(SALT_KEY)) >> np.uint64(1)) % np.uint64(g)
    r = (_mix_np(u ^ s ^ salt ^ np.uint64(SALT_MEMBER)) >> np.uint64(40)).astype(np.int64)
    member = r < int(p_member * (1 << 24))
    c = _counts_np(_mix_np(u ^ s ^ salt ^ np.uint64(SALT_CA)))
    return key[member], c[member]


def list_torch(seed: int, k: int, universe: int, u0: int, u1: int, list_id: int, p_member: float, device="cuda", chunk: int = 1 << 26):
    import torch
    g = gap(k, universe)
    s = _s64(seed)
    salt = _s64(0xD1B54A32D192ED03 * (list_id + 1))
    thr = int(p_member * (1 << 24))
    ws, cs = [], []
    for c0 in range(u0, u1, chunk):
        c1 = min(c0 + chunk, u1)
        u = torch.arange(c0, c1, dtype=torch.int64, device=device)
        key = u * g + torch.remainder(_lsr(_mix_t(u ^ s ^ _s64(SALT_KEY)), 1), g)
        member = _lsr(_mix_t(u ^ s ^ salt ^ _s64(SALT_MEMBER)), 40) < thr
        c = _counts_t(_mix_t(u ^ s ^ salt ^ _s64(SALT_CA)))
        ws.append(key[member]); cs.append(c[member])
        del u, key, member, c
    return (torch.cat(ws) if len(ws) != 1 else ws[0]), (torch.cat(cs) if len(cs) != 1 else cs[0])
