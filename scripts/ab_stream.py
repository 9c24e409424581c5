#!/usr/bin/env python
"""Timing + checksums of the single-output merge kernel for A/B runs of library variants (GT4GPU_LIB=...): the checksums
of two variants must agree.  Usage: ab_stream.py [n_per_list] [ops...]      ops from union, intrsec, diff1 (default: union)"""
import json, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import genometester4_b200 as g
from genometester4_b200 import synth

n = float(sys.argv[1]) if len(sys.argv) > 1 else 1e9
ops = sys.argv[2:] or ["union"]
g.init(0); g.set_stream(torch.cuda.current_stream().cuda_stream)
if os.environ.get("STREAM_SIDE") is not None:
    g.set_option("stream_side", int(os.environ["STREAM_SIDE"]))
m = int(round(1.5 * n))
(wa, ca), (wb, cb) = synth.pair_torch(42, 25, m, 0, m, 1 / 3, 1 / 3)
na, nb = wa.numel(), wb.numel()
la = g.WordList.from_device(wa.data_ptr(), ca.data_ptr(), na, 25); lb = g.WordList.from_device(wb.data_ptr(), cb.data_ptr(), nb, 25)
cap = na + nb
FLAG = {"union": "union", "intrsec": "intrsec", "diff1": "diff"}
ow = torch.empty(cap, dtype=torch.int64, device="cuda"); oc = torch.empty(cap, dtype=torch.int32, device="cuda")


def run(op, debug, reps=6):
    os.environ["GT4GPU_DEBUG"] = str(debug)
    ms = []
    for it in range(reps):
        r = g.compare_wordmaps(la, lb, **{"find_" + FLAG[op]: 1}, out_buffers={op: (ow.data_ptr(), oc.data_ptr(), cap)})
        r = next(iter(r.values()))
        if it >= 2: ms.append(g.last_timing()[1])
    return r, sum(ms) / max(len(ms), 1)


for op in ops:
    ow.fill_(-1); oc.fill_(-1)
    r, t = run(op, 0)
    k = ow[:r.n_words]; c = oc[:r.n_words]
    idx = torch.arange(r.n_words, device="cuda", dtype=torch.int64)
    chk = [int(k.sum()), int((k * (idx | 1)).sum()), int(c.to(torch.int64).sum()), int((c.to(torch.int64) * (idx | 1)).sum())]
    del idx
    b = 12 * (na + nb) + 12 * r.n_words
    print(json.dumps(dict(op=op, n_in=na + nb, n_out=r.n_words, total=r.total_count, ms=round(t, 3), gbs=round(b / t / 1e6, 1), checksums=chk,
                          lib=os.environ.get("GT4GPU_LIB", "default"))), flush=True)
    run(op, 32, reps=2)
    run(op, 2, reps=2)
os.environ["GT4GPU_DEBUG"] = "0"
