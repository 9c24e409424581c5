#!/usr/bin/env python
"""Device-resident timings of the other BASELINE.json configurations on ONE GPU (per-GPU shard sizes for the
8-GPU configs).  Prints one JSON line per measurement; algorithmic bytes = 12 B x inputs + 12 B x outputs."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import genometester4_b200 as g
from genometester4_b200 import synth

PEAK = json.load(open(Path(__file__).resolve().parent.parent / 'MEASURED_PEAKS.json'))['hbm_gbs'] if (Path(__file__).resolve().parent.parent / 'MEASURED_PEAKS.json').exists() else 6541.5
g.init(0)
g.set_stream(torch.cuda.current_stream().cuda_stream)
which = sys.argv[1:] or ["3", "4", "5"]


def pair(n_a, n_b, n_both, k, seed=42):
    m = int(n_a + n_b - n_both)
    (wa, ca), (wb, cb) = synth.pair_torch(seed, k, m, 0, m, (n_a - n_both) / m, (n_b - n_both) / m)
    la = g.WordList.from_device(wa.data_ptr(), ca.data_ptr(), wa.numel(), k, keepalive=(wa, ca))
    lb = g.WordList.from_device(wb.data_ptr(), cb.data_ptr(), wb.numel(), k, keepalive=(wb, cb))
    return la, lb


def timed(fn, reps=5):
    ms = []
    for it in range(reps + 2):
        r = fn()
        if it >= 2:
            p, m, _ = g.last_timing()
            ms.append((p, m))
    return r, sum(x[0] for x in ms) / len(ms), sum(x[1] for x in ms) / len(ms)


def emit(**kw):
    print(json.dumps(kw), flush=True)


if "3" in which:   # config 3: -d -c 5, 32-mers, 3e9 / 1e9 over 8 GPUs -> one shard = 3.75e8 / 1.25e8
    la, lb = pair(3.75e8, 1.25e8, 1.0e8, 32)
    for co in (0, 1):
        r, p_ms, m_ms = timed(lambda: g.compare_wordmaps(la, lb, find_diff=1, cutoff=5, countonly=co)["diff1"])
        b = 12 * (len(la) + len(lb)) + (0 if co else 12 * r.n_words)
        emit(config=3, what="glistcompare -d -c 5, k=32, one of 8 key-range shards", countonly=co, n_a=len(la), n_b=len(lb),
             n_out=r.n_words, partition_ms=round(p_ms, 3), merge_ms=round(m_ms, 3), gbs=round(b / m_ms / 1e6, 1), frac=round(b / m_ms / 1e6 / PEAK, 3))
    del la, lb
    torch.cuda.empty_cache()

if "4" in which:   # config 4: counts-only intersect / union sweep
    for n in (1e6, 1e7, 1e8, 1e9):
        for ov in (0.01, 0.10, 0.50, 0.90, 0.99):
            la, lb = pair(n, n, ov * n, 25)
            for op, kw, name in (("intersect", dict(find_intrsec=1), "intrsec"), ("union", dict(find_union=1), "union")):
                r, p_ms, m_ms = timed(lambda: g.compare_wordmaps(la, lb, countonly=1, **kw)[name], reps=4)
                b = 12 * (len(la) + len(lb))
                emit(config=4, op=op, n_per_list=int(n), overlap=ov, n_out=r.n_words, partition_ms=round(p_ms, 4), merge_ms=round(m_ms, 4),
                     kmers_per_s=round((len(la) + len(lb)) / ((p_ms + m_ms) / 1e3)), gbs=round(b / m_ms / 1e6, 1), frac=round(b / m_ms / 1e6 / PEAK, 3))
            del la, lb
            torch.cuda.empty_cache()

if "5" in which:   # config 5: 8-list union (MakeUnion.pl equivalent), 8 x 5e8 over 8 GPUs -> one shard = 8 x 6.25e7
    for n_each in (6.25e7, 2.5e8):
        m = int(n_each * 3)
        lists, keep = [], []
        for j in range(8):
            wa, ca = synth.list_torch(5, 32, m, 0, m, j, 1 / 3)                    # every list: a third of ONE shared universe
            keep.append((wa, ca))
            lists.append(g.WordList.from_device(wa.data_ptr(), ca.data_ptr(), wa.numel(), 32))
        n_in = sum(len(l) for l in lists)
        for kway, what in ((1, "8-list union, single-pass k-way kernel"), (0, "8-list union, tree of two-list merges")):
            g.set_option("use_kway", 2 if kway else 0)
            for op, fn in (("union", g.union_multi), ("intersect", g.intersect_multi)):
                for co in (0, 1):
                    r, p_ms, m_ms = timed(lambda: fn(lists, cutoff=1, countonly=co), reps=3)
                    b = 12 * n_in + (0 if co else 12 * r.n_words)
                    emit(config=5, what=what, op=op, countonly=co, n_each=int(n_each), n_in=n_in, n_out=r.n_words,
                         prepass_ms=round(p_ms, 3), merge_ms=round(m_ms, 3), kernels_ms=round(p_ms + m_ms, 3),
                         gbs_algorithmic=round(b / (p_ms + m_ms) / 1e6, 1), frac=round(b / (p_ms + m_ms) / 1e6 / PEAK, 3),
                         frac_merge_kernel=round(b / max(m_ms, 1e-6) / 1e6 / PEAK, 3))
                    del r
        g.set_option("use_kway", 1)
        del lists, keep
        torch.cuda.empty_cache()
