#!/usr/bin/env python
"""Tile-shape / operation sweep on one GPU: device time of the partition and tile kernels per configuration.
Usage: python scripts/sweep_tiles.py [n_per_list] [overlap]"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import genometester4_b200 as g
from genometester4_b200 import synth

n = float(sys.argv[1]) if len(sys.argv) > 1 else 1e9
overlap = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
shapes = [("s256", 7), ("s256", 9), ("s256", 11), ("s512", 7), ("s512", 9), ("s512", 11), (256, 7)]
g.init(0)
g.set_stream(torch.cuda.current_stream().cuda_stream)
m = int(round(2 * n - overlap * n))
p = (n - overlap * n) / m
(wa, ca), (wb, cb) = synth.pair_torch(42, 25, m, 0, m, p, p)
na, nb = wa.numel(), wb.numel()
la = g.WordList.from_device(wa.data_ptr(), ca.data_ptr(), na, 25)
lb = g.WordList.from_device(wb.data_ptr(), cb.data_ptr(), nb, 25)
ow = torch.empty(na + nb, dtype=torch.int64, device="cuda")
oc = torch.empty(na + nb, dtype=torch.int32, device="cuda")
rows = []
for op, kw, name in (("union", dict(find_union=1), "union"), ("intersect", dict(find_intrsec=1), "intrsec"), ("diff", dict(find_diff=1), "diff1")):
    for countonly in (0, 1):
        for nt, vt in shapes:
            if isinstance(nt, str):
                g.set_option("use_stream_kernel", 1)
                g.set_option("stream_shape", int(nt[1:]) * 100 + vt)
            else:
                g.set_option("use_stream_kernel", 0)
                g.set_tile(nt, vt)
            ms = []
            for it in range(6):
                r = g.compare_wordmaps(la, lb, countonly=countonly, out_buffers=None if countonly else {name: (ow.data_ptr(), oc.data_ptr(), na + nb)}, **kw)[name]
                if it >= 2:
                    ms.append(g.last_timing())
            part = sum(x[0] for x in ms) / len(ms)
            merge = sum(x[1] for x in ms) / len(ms)
            bytes_ = 12 * (na + nb) + (0 if countonly else 12 * r.n_words)
            row = dict(op=op, countonly=countonly, tile=f"{nt}x{vt}", part_ms=round(part, 4), merge_ms=round(merge, 4),
                       gbs=round(bytes_ / merge / 1e6, 1), n_out=r.n_words)
            rows.append(row)
            print(json.dumps(row), flush=True)
