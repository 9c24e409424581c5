#!/usr/bin/env python
"""Headline merge (union, 1e9 + 1e9 25-mers) over every supported shape of the single-output kernel.
Usage: sweep_stream.py [n_per_list] [shapes...]"""
import json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import genometester4_b200 as g
from genometester4_b200 import synth
n = float(sys.argv[1]) if len(sys.argv) > 1 else 1e9
shapes = sys.argv[2:] or ["256x7", "256x9", "256x11", "384x9", "384x11", "512x7", "512x9", "512x11"]
g.init(0); g.set_stream(torch.cuda.current_stream().cuda_stream)
m = int(round(1.5 * n))
(wa, ca), (wb, cb) = synth.pair_torch(42, 25, m, 0, m, 1 / 3, 1 / 3)
na, nb = wa.numel(), wb.numel()
la = g.WordList.from_device(wa.data_ptr(), ca.data_ptr(), na, 25); lb = g.WordList.from_device(wb.data_ptr(), cb.data_ptr(), nb, 25)
ow = torch.empty(na + nb, dtype=torch.int64, device="cuda"); oc = torch.empty(na + nb, dtype=torch.int32, device="cuda")
for shp in shapes:
    nc, vt = (int(x) for x in shp.split("x"))
    g.set_option("stream_shape", nc * 100 + vt)
    for co in (0, 1):
        ms = []
        for it in range(5):
            r = g.compare_wordmaps(la, lb, countonly=co, find_union=1, out_buffers=None if co else {"union": (ow.data_ptr(), oc.data_ptr(), na + nb)})["union"]
            if it >= 2: ms.append(g.last_timing()[1])
        t = sum(ms) / len(ms)
        b = 12 * (na + nb) + (0 if co else 12 * r.n_words)
        print(json.dumps(dict(shape=shp, countonly=co, merge_ms=round(t, 3), gbs=round(b / t / 1e6, 1))), flush=True)
