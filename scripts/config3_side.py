#!/usr/bin/env python
"""Config 3 at full size on one GPU (3e9 / 1e9 32-mers, -d -c 5) with the side-buffer variants of the stream kernel off (0), chosen
by the density sample (1) and forced (2 = three stages, 3 = two stages + three slots).  Usage: config3_side.py [scale]"""
import json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import genometester4_b200 as g
from genometester4_b200 import synth
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
g.init(0); g.set_stream(torch.cuda.current_stream().cuda_stream)
n_a, n_b, n_both = 3e9 * scale, 1e9 * scale, 0.8e9 * scale
m = int(n_a + n_b - n_both)
(wa, ca), (wb, cb) = synth.pair_torch(42, 32, m, 0, m, (n_a - n_both) / m, (n_b - n_both) / m)
la = g.WordList.from_device(wa.data_ptr(), ca.data_ptr(), wa.numel(), 32); lb = g.WordList.from_device(wb.data_ptr(), cb.data_ptr(), wb.numel(), 32)
cap = wa.numel()
ow = torch.empty(cap, dtype=torch.int64, device="cuda"); oc = torch.empty(cap, dtype=torch.int32, device="cuda")
ref = None
for side in (0, 1, 3, 2):
    g.set_option("stream_side", side)
    ms = []
    for it in range(5):
        r = g.compare_wordmaps(la, lb, find_diff=1, cutoff=5, out_buffers={"diff1": (ow.data_ptr(), oc.data_ptr(), cap)})["diff1"]
        if it >= 2: ms.append(g.last_timing()[1])
    chk = (r.n_words, r.total_count, int(ow[:r.n_words].sum()), int(oc[:r.n_words].to(torch.int64).sum()))
    ref = ref or chk
    t = sum(ms) / len(ms)
    print(json.dumps(dict(stream_side=side, merge_ms=round(t, 3), n_in=wa.numel() + wb.numel(), n_out=r.n_words, density=round(r.n_words / (wa.numel() + wb.numel()), 3),
                          gbs=round(12 * (wa.numel() + wb.numel() + r.n_words) / t / 1e6, 1), same_as_plain=chk == ref)), flush=True)
g.set_option("stream_side", 1)
