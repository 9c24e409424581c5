#!/usr/bin/env python
"""Target for ncu: a few gt4gpu_count_words calls on 1e8 random 25-mers."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import genometester4_b200 as g
g.init(0); g.set_stream(torch.cuda.current_stream().cuda_stream)
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
gen = torch.Generator(device="cuda").manual_seed(1)
w = torch.randint(0, 4 ** 25, (n,), generator=gen, device="cuda", dtype=torch.int64)
torch.cuda.synchronize()
for _ in range(2):
    r = g.count_words(w.data_ptr(), 25, n_words=n); r.free()
