#!/usr/bin/env python
"""Several outputs of one two-list merge: the fused kernel (one read of the lists) vs one pass of the single-output kernel
per output, device resident.  Usage: multi_output.py [n_per_list]"""
import json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import genometester4_b200 as g
from genometester4_b200 import synth
n = float(sys.argv[1]) if len(sys.argv) > 1 else 1e9
PEAK = json.load(open(Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json"))["hbm_gbs"]
g.init(0); g.set_stream(torch.cuda.current_stream().cuda_stream)
m = int(1.5 * n)
(wa, ca), (wb, cb) = synth.pair_torch(42, 25, m, 0, m, 1 / 3, 1 / 3)
torch.cuda.synchronize()
la = g.WordList.from_device(wa.data_ptr(), ca.data_ptr(), wa.numel(), 25); lb = g.WordList.from_device(wb.data_ptr(), cb.data_ptr(), wb.numel(), 25)
for label, kw in (("-u -i", dict(find_union=1, find_intrsec=1)), ("-u -i -d", dict(find_union=1, find_intrsec=1, find_diff=1)),
                  ("-u -i -dd", dict(find_union=1, find_intrsec=1, find_ddiff=1)), ("-i -dd", dict(find_intrsec=1, find_ddiff=1))):
    for fused in (1, 0):
        g.set_option("use_fused", fused)
        ms = []
        for it in range(4):
            r = g.compare_wordmaps(la, lb, cutoff=1, **kw)
            if it >= 1: ms.append(sum(g.last_timing()[:2]))
            n_out = sum(x.n_words for x in r.values())
            del r
        t = sum(ms) / len(ms)
        b = 12 * (len(la) + len(lb)) + 12 * n_out
        print(json.dumps(dict(ops=label, fused=bool(fused), ms=round(t, 3), n_in=len(la) + len(lb), n_out=n_out, algorithmic_gb=round(b / 1e9, 2),
                              gbs=round(b / t / 1e6, 1), frac_of_measured_peak=round(b / t / 1e6 / PEAK, 3))), flush=True)
g.set_option("use_fused", 1)
