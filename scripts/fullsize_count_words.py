#!/usr/bin/env python
"""gt4gpu_count_words at a size whose element indices pass 2^31: 3e9 device-resident words with a known multiset
(word i = (i * A) mod M, M prime), checked through size-independent properties: strictly ascending distinct words, every
residue present, counts = floor/ceil of n / M exactly as the construction dictates, sum of counts = n."""
import json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import genometester4_b200 as g

g.init(0); g.set_stream(torch.cuda.current_stream().cuda_stream)
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 3_000_000_000
M = 1_000_000_007
A = 2_654_435_761
k = 16                      # words < 4^16 = 2^32 > M
w = torch.empty(n, dtype=torch.int64, device="cuda")
step = 1 << 28
for lo in range(0, n, step):
    hi = min(n, lo + step)
    i = torch.arange(lo, hi, dtype=torch.int64, device="cuda")
    w[lo:hi] = (i % M) * A % M          # (i*A) mod M without overflowing 63 bits: i % M < 2^30, A < 2^32
    del i
torch.cuda.synchronize()
t0 = time.time()
res = g.count_words(w.data_ptr(), k, n_words=n)
dt = time.time() - t0
tw, tc = res.as_torch()
ok_sorted = bool((tw[1:] > tw[:-1]).all())
# i -> (i mod M) * A mod M is a bijection on residues; residue r of i occurs ceil/floor(n / M) times
q, rem = divmod(n, M)
ok_unique = res.n_words == min(n, M)
ok_counts = bool(((tc == q) | (tc == q + 1)).all()) and int((tc == q + 1).sum()) == rem
ok_sum = int(tc.to(torch.int64).sum()) == n and res.total_count == n
ok_range = int(tw[0]) == 0 and int(tw[-1]) == M - 1
print(json.dumps({"n_words": n, "k": k, "n_unique": res.n_words, "wall_s": round(dt, 3), "timing_ms": g.last_timing()[:2],
                  "strictly_ascending": ok_sorted, "all_residues_present": ok_unique and ok_range,
                  "counts_exact": ok_counts, "sum_counts_equals_n": ok_sum}))
assert ok_sorted and ok_unique and ok_counts and ok_sum and ok_range
