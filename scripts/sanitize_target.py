"""Small workload touching every kernel: target of compute-sanitizer runs."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import genometester4_b200 as g
from genometester4_b200 import synth, api
g.init(0)
def rand_bytes(rng, alphabet: bytes, n) -> bytes:
    """n random characters of `alphabet` (one byte each: bytes() of an int64 array would be its raw 8-byte image)."""
    return rng.choice(np.frombuffer(alphabet, dtype=np.uint8), size=int(n)).tobytes()
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
# list building, device FastA reader, lookups
rng = np.random.default_rng(2)
for k, n in ((5, 1), (16, 8191), (25, 8193), (32, 40_000)):
    hi = min(4 ** k, 2 ** 63)
    raw = rng.integers(0, hi, size=n, dtype=np.uint64)
    res = g.count_words(raw, k)
    w, c = res.to_host()
    assert int(c.sum()) == n
    lst = g.WordList.from_arrays(w, c, k)
    g.lookup(lst, rng.integers(0, hi, size=3001, dtype=np.uint64))
for text in (b">a\nACGT", b">x\n" + rand_bytes(rng, b"ACGTN\n", 20_001) + b"\n>y z\n" + rand_bytes(rng, b"ACGT", 4095),
             b">" + b"n" * 5000 + b"\n" + b"ACGT" * 2500):
    for k in (1, 13, 32):
        d = g.fasta_words_device(text, k)
        if d.n_words:
            g.count_words(d.ptr, k, n_words=d.n_words).to_host()
        d.free()
print("sanitize target ok")
