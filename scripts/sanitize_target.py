"""Small workload touching every kernel: target of compute-sanitizer runs."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import genometester4_b200 as g
from genometester4_b200 import synth, api
g.init(0)
(wa, ca), (wb, cb) = synth.pair_numpy(42, 25, 60_000, 0, 60_000, 1 / 3, 1 / 3)
la, lb = g.WordList.from_arrays(wa, ca, 25), g.WordList.from_arrays(wb, cb, 25)
for shape in (51209, 25607):
    g.set_option("stream_shape", shape)
    for kw in (dict(find_union=1), dict(find_intrsec=1, rule="max"), dict(find_ddiff=1, cutoff=3), dict(find_union=1, find_intrsec=1, find_diff=1)):
        r = g.compare_wordmaps(la, lb, **kw)
        c = g.compare_wordmaps(la, lb, countonly=1, **kw)
        for s in r:
            assert (r[s].n_words, r[s].total_count) == (c[s].n_words, c[s].total_count)
            r[s].records()
g.set_option("use_stream_kernel", 0)
g.compare_wordmaps(la, lb, find_union=1, find_intrsec=1, find_diff=1, find_ddiff=1)
g.set_option("use_stream_kernel", 1)
lists = [g.WordList.from_arrays(*synth.list_numpy(5, 25, 30_000, 0, 30_000, j, 0.4), 25) for j in range(5)]
g.union_multi(lists, cutoff=2).records(); g.intersect_multi(lists).records(); g.gt4_union(lists); g.gt4_is_union(lists)
print("sanitize target ok")
