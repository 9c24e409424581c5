#!/usr/bin/env python
"""BASELINE config 3 at FULL size on one GPU: 3e9 + 1e9 32-mers (4e9 records: every 64-bit index path is exercised),
`-d -c 5`, `-u` and `-i`.  No CPU check finishes at this size, so the one-shot result is compared with the same
operation done as 8 independent key-range shards (each < 2^32 records), through totals and order-sensitive checksums."""
import json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import genometester4_b200 as g
from genometester4_b200 import synth

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
n_a, n_b, n_both, k, parts = 3e9 * scale, 1e9 * scale, 0.8e9 * scale, 32, 8
g.init(0)
g.set_stream(torch.cuda.current_stream().cuda_stream)
m = int(n_a + n_b - n_both)
t0 = time.time()
(wa, ca), (wb, cb) = synth.pair_torch(42, k, m, 0, m, (n_a - n_both) / m, (n_b - n_both) / m)
torch.cuda.synchronize()
print(f"generated |A|={wa.numel()} |B|={wb.numel()} in {time.time() - t0:.1f}s, HBM in use {torch.cuda.memory_allocated() / 1e9:.1f} GB", flush=True)
la = g.WordList.from_device(wa.data_ptr(), ca.data_ptr(), wa.numel(), k)
lb = g.WordList.from_device(wb.data_ptr(), cb.data_ptr(), wb.numel(), k)


def checksum(res):
    w, c = res.as_torch()
    idx = torch.arange(1, w.numel() + 1, dtype=torch.int64, device="cuda")
    return int((w * idx).sum()), int((c.to(torch.int64) * idx).sum()), int(w[1:].gt(w[:-1]).all()) if k < 32 else None


def shard_bounds(keys, n_parts):
    # equal-count splitters on A's keys; B cut by key value (lower_bound) so equal keys share a shard
    cuts = [0] + [int(keys.numel() * p / n_parts) for p in range(1, n_parts)] + [keys.numel()]
    return cuts


for name, kw in (("diff1", dict(find_diff=1, cutoff=5)), ("union", dict(find_union=1)), ("intrsec", dict(find_intrsec=1))):
    res = g.compare_wordmaps(la, lb, **kw)[name]
    p_ms, m_ms, _ = g.last_timing()
    full = (res.n_words, res.total_count)
    w_full, c_full = res.as_torch()
    # position-weighted checksums (wrap mod 2^64: exact in int64 arithmetic), shard by shard
    cuts_a = shard_bounds(wa, parts)
    off, n_sum, t_sum, ok = 0, 0, 0, True
    for p in range(parts):
        a0, a1 = cuts_a[p], cuts_a[p + 1]
        lo_key = wa[a0] if p > 0 else None
        hi_key = wa[a1] if a1 < wa.numel() else None
        # unsigned compare of 64-bit keys held in int64: flip the sign bit
        flip = -(1 << 63)
        def ub(x): return int(torch.searchsorted(wb ^ flip, (x ^ flip).reshape(1)).item())
        b0 = ub(lo_key) if lo_key is not None else 0
        b1 = ub(hi_key) if hi_key is not None else wb.numel()
        sa = g.WordList.from_device(wa[a0:a1].data_ptr(), ca[a0:a1].data_ptr(), a1 - a0, k)
        sb = g.WordList.from_device(wb[b0:b1].data_ptr(), cb[b0:b1].data_ptr(), b1 - b0, k)
        r = g.compare_wordmaps(sa, sb, **kw)[name]
        w, c = r.as_torch()
        ok &= bool(torch.equal(w, w_full[off:off + r.n_words])) and bool(torch.equal(c, c_full[off:off + r.n_words]))
        off += r.n_words; n_sum += r.n_words; t_sum += r.total_count
        del r, w, c
    bytes_ = 12 * (wa.numel() + wb.numel()) + 12 * full[0]
    print(json.dumps(dict(op=name, n_a=wa.numel(), n_b=wb.numel(), n_out=full[0], total_count=full[1], shards_n=n_sum, shards_total=t_sum,
                          records_identical_to_8_shards=bool(ok and (n_sum, t_sum) == full), partition_ms=round(p_ms, 3), merge_ms=round(m_ms, 3),
                          gbs=round(bytes_ / m_ms / 1e6, 1))), flush=True)
    del res, w_full, c_full
    torch.cuda.empty_cache()
