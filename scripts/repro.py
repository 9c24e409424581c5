import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import genometester4_b200 as g
from genometester4_b200 import synth
g.init(0); g.set_stream(torch.cuda.current_stream().cuda_stream)
n, ov = float(sys.argv[1]), float(sys.argv[2])
m = int(2 * n - ov * n)
(wa, ca), (wb, cb) = synth.pair_torch(42, 25, m, 0, m, (n - ov * n) / m, (n - ov * n) / m)
print("sizes", wa.numel(), wb.numel(), hex(wa.data_ptr()), hex(wb.data_ptr()), hex(ca.data_ptr()), hex(cb.data_ptr()), flush=True)
la = g.WordList.from_device(wa.data_ptr(), ca.data_ptr(), wa.numel(), 25); lb = g.WordList.from_device(wb.data_ptr(), cb.data_ptr(), wb.numel(), 25)
for op, kw, name in (("intersect", dict(find_intrsec=1), "intrsec"), ("union", dict(find_union=1), "union")):
    for co in (1, 0):
        r = g.compare_wordmaps(la, lb, countonly=co, **kw)[name]
        print(op, co, r.n_words, r.total_count, flush=True)
