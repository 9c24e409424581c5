#!/usr/bin/env python
"""Device-resident timing of the list-building back end (gt4gpu_count_words): radix sort + run-length counts of raw
k-mer words, like one glistmaker table.  Prints one JSON line per configuration."""
import json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import genometester4_b200 as g

PEAK = 6549.1
try:
    PEAK = json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]
except Exception:
    pass

def run(n, k, coverage, reps=4):
    gen = torch.Generator(device="cuda").manual_seed(1)
    distinct = max(1, int(n / coverage))
    # `coverage` occurrences per distinct word on average: draw indices into a universe of spread-out words
    idx = torch.randint(0, distinct, (n,), generator=gen, device="cuda", dtype=torch.int64)
    stride = max(1, (4 ** k) // distinct) if k < 32 else (2 ** 63) // distinct
    w = idx * stride + (idx * 2654435761 % max(1, min(stride, 2 ** 31)))
    del idx
    torch.cuda.synchronize()
    ms_sort, ms_rle = [], []
    for it in range(reps):
        res = g.count_words(w.data_ptr(), k, n_words=n)
        t = g.last_timing()
        if it >= 1:
            ms_sort.append(t[0]); ms_rle.append(t[1])
        nu = res.n_words
        res.free()
    n_pass = (2 * k + 7) // 8
    s, r = sum(ms_sort) / len(ms_sort), sum(ms_rle) / len(ms_rle)
    sort_bytes = 8 * n * (1 + 2 * n_pass)
    rle_bytes = 8 * n + 16 * nu + 16 * nu + 12 * nu
    print(json.dumps({"n_words": n, "k": k, "passes": n_pass, "n_unique": nu, "sort_ms": round(s, 3), "rle_ms": round(r, 3),
                      "words_per_s": round(n / (s + r) * 1e3), "sort_gbs": round(sort_bytes / s / 1e6, 1),
                      "sort_frac_of_hbm_peak": round(sort_bytes / s / 1e6 / PEAK, 3),
                      "rle_gbs": round(rle_bytes / r / 1e6, 1)}), flush=True)

if __name__ == "__main__":
    g.init(0)
    g.set_stream(torch.cuda.current_stream().cuda_stream)
    sizes = [float(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["1e8", "1e9"])]
    for n in sizes:
        for k, cov in ((16, 30.0), (25, 1.2), (32, 30.0)):
            run(int(n), k, cov)
