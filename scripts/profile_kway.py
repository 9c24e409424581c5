#!/usr/bin/env python
"""Config-5 shaped N-list union / intersection, few launches: the target of ncu captures and quick timings.
Usage: profile_kway.py n_each n_lists [op] [countonly] [reps]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import genometester4_b200 as g
from genometester4_b200 import synth
n_each = int(float(sys.argv[1])); n_lists = int(sys.argv[2])
op = sys.argv[3] if len(sys.argv) > 3 else "union"
co = int(sys.argv[4]) if len(sys.argv) > 4 else 0
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
g.init(0); g.set_stream(torch.cuda.current_stream().cuda_stream)
g.set_option("use_kway", 2)
m = n_each * 3
lists, keep = [], []
for j in range(n_lists):
    w, c = synth.list_torch(5, 32, m, 0, m, j, 1 / 3)
    keep.append((w, c))
    lists.append(g.WordList.from_device(w.data_ptr(), c.data_ptr(), w.numel(), 32))
n_in = sum(len(l) for l in lists)
fn = g.union_multi if op == "union" else g.intersect_multi
for it in range(reps):
    r = fn(lists, cutoff=1, countonly=co)
    p_ms, m_ms, nl = g.last_timing()
    b = 12 * n_in + (0 if co else 12 * r.n_words)
    print(f"{op} n_lists={n_lists} n_in={n_in} n_out={r.n_words} prepass {p_ms:.3f} ms merge {m_ms:.3f} ms -> {b / m_ms / 1e6:.0f} GB/s", flush=True)
