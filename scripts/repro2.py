import sys, os
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import genometester4_b200 as g
from genometester4_b200 import synth
g.init(0); g.set_stream(torch.cuda.current_stream().cuda_stream)
cases = [(float(a), float(b)) for a, b in (x.split(":") for x in sys.argv[1].split(","))]
for n, ov in cases:
    m = int(2 * n - ov * n)
    (wa, ca), (wb, cb) = synth.pair_torch(42, 25, m, 0, m, (n - ov * n) / m, (n - ov * n) / m)
    print("case", n, ov, wa.numel(), wb.numel(), hex(wa.data_ptr()), hex(wb.data_ptr()), hex(ca.data_ptr()), hex(cb.data_ptr()), flush=True)
    la = g.WordList.from_device(wa.data_ptr(), ca.data_ptr(), wa.numel(), 25); lb = g.WordList.from_device(wb.data_ptr(), cb.data_ptr(), wb.numel(), 25)
    for op, kw, name in (("intersect", dict(find_intrsec=1), "intrsec"), ("union", dict(find_union=1), "union")):
        for it in range(3):
            r = g.compare_wordmaps(la, lb, countonly=1, **kw)[name]
        print(op, r.n_words, r.total_count, flush=True)
    del la, lb, wa, ca, wb, cb
    torch.cuda.empty_cache()
