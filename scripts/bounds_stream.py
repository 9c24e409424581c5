#!/usr/bin/env python
"""Upper bounds of the single-output merge (library built with -DGT4GPU_UNSAFE_EXPERIMENTS): time of the union / intersection
pass with the look-back skipped (GT4GPU_DEBUG bit 0: wrong offsets), with the output stores skipped (bit 3), with both."""
import json, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import genometester4_b200 as g
from genometester4_b200 import synth
n = float(sys.argv[1]) if len(sys.argv) > 1 else 1e9
g.init(0); g.set_stream(torch.cuda.current_stream().cuda_stream)
m = int(round(1.5 * n))
(wa, ca), (wb, cb) = synth.pair_torch(42, 25, m, 0, m, 1 / 3, 1 / 3)
na, nb = wa.numel(), wb.numel()
la = g.WordList.from_device(wa.data_ptr(), ca.data_ptr(), na, 25); lb = g.WordList.from_device(wb.data_ptr(), cb.data_ptr(), nb, 25)
cap = na + nb
ow = torch.empty(cap, dtype=torch.int64, device="cuda"); oc = torch.empty(cap, dtype=torch.int32, device="cuda")
for op, flag in (("union", "union"), ("intrsec", "intrsec")):
    for dbg in (0, 1, 8, 9, 32, 33, 40, 41):
        os.environ["GT4GPU_DEBUG"] = str(dbg)
        ms = []
        for it in range(5):
            g.compare_wordmaps(la, lb, **{"find_" + flag: 1}, out_buffers={op: (ow.data_ptr(), oc.data_ptr(), cap)})
            if it >= 2: ms.append(g.last_timing()[1])
        print(json.dumps(dict(op=op, debug=dbg, ms=round(sum(ms) / len(ms), 3))), flush=True)
