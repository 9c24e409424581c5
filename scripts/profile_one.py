#!/usr/bin/env python
"""One configuration, few launches: the target of ncu captures.
Usage: profile_one.py n_per_list shape(e.g. 256x11) countonly(0/1) [op]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import genometester4_b200 as g
from genometester4_b200 import synth
n = float(sys.argv[1]); nc, vt = (int(x) for x in sys.argv[2].split("x")); co = int(sys.argv[3])
op = sys.argv[4] if len(sys.argv) > 4 else "union"
g.init(0); g.set_stream(torch.cuda.current_stream().cuda_stream)
g.set_option("stream_shape", nc * 100 + vt)
m = int(round(1.5 * n))
(wa, ca), (wb, cb) = synth.pair_torch(42, 25, m, 0, m, 1 / 3, 1 / 3)
na, nb = wa.numel(), wb.numel()
la = g.WordList.from_device(wa.data_ptr(), ca.data_ptr(), na, 25); lb = g.WordList.from_device(wb.data_ptr(), cb.data_ptr(), nb, 25)
ow = torch.empty(na + nb, dtype=torch.int64, device="cuda"); oc = torch.empty(na + nb, dtype=torch.int32, device="cuda")
name = {"union": "union", "intersect": "intrsec", "diff": "diff1"}[op]
kw = {"union": dict(find_union=1), "intersect": dict(find_intrsec=1), "diff": dict(find_diff=1)}[op]
for it in range(3):
    r = g.compare_wordmaps(la, lb, countonly=co, out_buffers=None if co else {name: (ow.data_ptr(), oc.data_ptr(), na + nb)}, **kw)[name]
print(r.n_words, g.last_timing())
