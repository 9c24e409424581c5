#!/usr/bin/env python
"""Fused multi-output pass (setop2_tile_kernel) vs one stream-kernel pass per output, 1e9 + 1e9."""
import json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import genometester4_b200 as g
from genometester4_b200 import synth
n = float(sys.argv[1]) if len(sys.argv) > 1 else 1e9
g.init(0); g.set_stream(torch.cuda.current_stream().cuda_stream)
m = int(1.5 * n)
(wa, ca), (wb, cb) = synth.pair_torch(42, 25, m, 0, m, 1 / 3, 1 / 3)
torch.cuda.synchronize()
la = g.WordList.from_device(wa.data_ptr(), ca.data_ptr(), wa.numel(), 25); lb = g.WordList.from_device(wb.data_ptr(), cb.data_ptr(), wb.numel(), 25)
for label, kw in (("-u -i", dict(find_union=1, find_intrsec=1)), ("-u -i -d", dict(find_union=1, find_intrsec=1, find_diff=1)), ("-u -i -dd", dict(find_union=1, find_intrsec=1, find_ddiff=1))):
    for tile in ((256, 7), (256, 9)):
        g.set_tile(*tile)
        ms = []
        for it in range(4):
            r = g.compare_wordmaps(la, lb, cutoff=2, **kw)
            if it >= 1: ms.append(sum(g.last_timing()[:2]))
            del r
        fused = sum(ms) / len(ms)
        print(json.dumps(dict(ops=label, tile=f"{tile[0]}x{tile[1]}", fused_ms=round(fused, 2))), flush=True)
    tot = 0
    for name, k1 in (("union", dict(find_union=1)), ("intrsec", dict(find_intrsec=1)), ("diff1", dict(find_diff=1)), ("diff2", dict(find_ddiff=1))):
        key = {"union": "find_union", "intrsec": "find_intrsec", "diff1": "find_diff", "diff2": "find_ddiff"}[name]
        if not kw.get(key) and not (name == "diff1" and kw.get("find_ddiff")): continue
        ms = []
        for it in range(3):
            r = g.compare_wordmaps(la, lb, cutoff=2, **k1)
            if it >= 1: ms.append(sum(g.last_timing()[:2]))
            del r
        tot += sum(ms) / len(ms)
    print(json.dumps(dict(ops=label, separate_stream_passes_ms=round(tot, 2))), flush=True)
