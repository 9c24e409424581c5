#!/usr/bin/env python
"""Few launches of the multi-output merge: the target of ncu captures.  Usage: profile_fused.py n_per_list ops(e.g. uid, idd)"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import genometester4_b200 as g
from genometester4_b200 import synth
n = float(sys.argv[1]); ops = sys.argv[2] if len(sys.argv) > 2 else "uid"
g.init(0); g.set_stream(torch.cuda.current_stream().cuda_stream)
m = int(round(1.5 * n))
(wa, ca), (wb, cb) = synth.pair_torch(42, 25, m, 0, m, 1 / 3, 1 / 3)
la = g.WordList.from_device(wa.data_ptr(), ca.data_ptr(), wa.numel(), 25); lb = g.WordList.from_device(wb.data_ptr(), cb.data_ptr(), wb.numel(), 25)
kw = dict(find_union=int("u" in ops), find_intrsec=int("i" in ops), find_diff=int("d" in ops), find_ddiff=int("dd" in ops))
for it in range(3):
    r = g.compare_wordmaps(la, lb, cutoff=1, **kw)
    print({k: v.n_words for k, v in r.items()}, g.last_timing(), flush=True)
    del r
