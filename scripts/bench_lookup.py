#!/usr/bin/env python
"""Device-resident timing of gt4gpu_lookup: random exact lookups (half present) in a sorted list."""
import json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import genometester4_b200 as g
from genometester4_b200 import synth

g.init(0); g.set_stream(torch.cuda.current_stream().cuda_stream)
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
nq = int(float(sys.argv[2])) if len(sys.argv) > 2 else 100_000_000
k = 25
(w, c) = synth.list_torch(42, k, int(n / 0.75), 0, int(n / 0.75), 0, 0.75)
lst = g.WordList.from_device(w.data_ptr(), c.data_ptr(), w.numel(), k)
gen = torch.Generator(device="cuda").manual_seed(3)
pick = torch.randint(0, w.numel(), (nq,), generator=gen, device="cuda")
q = w[pick]
q[::2] += 1                     # every other query misses (neighbouring words differ by more than 1 almost always)
out = torch.empty(nq, dtype=torch.int32, device="cuda")
torch.cuda.synchronize()
for sort_queries in (False, True):
    qq = torch.sort(q).values if sort_queries else q
    torch.cuda.synchronize()
    ms = []
    for it in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.lookup_device(lst, qq.data_ptr(), nq, out.data_ptr(), 0, canonize=False)
        e1.record(); torch.cuda.synchronize()
        if it: ms.append(e0.elapsed_time(e1))
    t = sum(ms) / len(ms)
    print(json.dumps({"list_words": int(w.numel()), "queries": nq, "sorted_queries": sort_queries, "ms": round(t, 3),
                      "lookups_per_s": round(nq / t * 1e3), "hits": int((out > 0).sum())}), flush=True)
