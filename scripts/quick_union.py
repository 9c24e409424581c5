#!/usr/bin/env python
"""Quick timing of the single-output union kernel for a few shapes (experiments)."""
import json, sys, os
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import genometester4_b200 as g
from genometester4_b200 import synth
n = float(sys.argv[1]) if len(sys.argv) > 1 else 1e9
shapes = [tuple(int(x) for x in s.split("x")) for s in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["256x11", "512x11"])]
g.init(0); g.set_stream(torch.cuda.current_stream().cuda_stream)
m = int(round(1.5 * n)); p = 1 / 3
(wa, ca), (wb, cb) = synth.pair_torch(42, 25, m, 0, m, p, p)
na, nb = wa.numel(), wb.numel()
la = g.WordList.from_device(wa.data_ptr(), ca.data_ptr(), na, 25); lb = g.WordList.from_device(wb.data_ptr(), cb.data_ptr(), nb, 25)
ow = torch.empty(na + nb, dtype=torch.int64, device="cuda"); oc = torch.empty(na + nb, dtype=torch.int32, device="cuda")
for nc, vt in shapes:
    g.set_option("stream_shape", nc * 100 + vt)
    for co in (0, 1):
        ms = []
        for it in range(6):
            r = g.compare_wordmaps(la, lb, find_union=1, countonly=co, out_buffers=None if co else {"union": (ow.data_ptr(), oc.data_ptr(), na + nb)})["union"]
            if it >= 2: ms.append(g.last_timing()[1])
        print(json.dumps(dict(shape=f"{nc}x{vt}", countonly=co, merge_ms=round(sum(ms) / len(ms), 3), debug=os.environ.get("GT4GPU_DEBUG", "0"))), flush=True)
