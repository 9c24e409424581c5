#!/usr/bin/env python
"""Generates tests/golden/query/: list files, query word files and the text the UNMODIFIED reference glistquery
(oracle/_ref/glistquery, built by oracle/Makefile) prints for `glistquery LIST -f QUERIES` (exact lookups) and
`glistquery LIST -l QUERYLIST` (the list-against-list zipper).  Run in the build container."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import oracle as O  # noqa: E402

OUT = Path(__file__).parent / "golden" / "query"


def canonical(words, k):
    out = []
    for w in words:
        r = 0
        x = ~int(w)
        for _ in range(k):
            r = (r << 2) | (x & 3)
            x >>= 2
        out.append(min(int(w), r))
    return np.unique(np.array(out, dtype=np.uint64))


def main():
    if O.ref_binary("glistquery") is None:
        raise SystemExit("oracle/_ref/glistquery missing: run make -C oracle")
    OUT.mkdir(parents=True, exist_ok=True)
    rng = np.random.default_rng(77)
    cases = []
    for k in (5, 16, 25, 32):
        hi = 4 ** k if k < 32 else 2 ** 64
        raw = rng.integers(0, min(hi, 2 ** 63), size=400, dtype=np.uint64)
        if k == 32:
            raw = raw * np.uint64(2) + rng.integers(0, 2, size=400, dtype=np.uint64)
        words = canonical(raw, k)
        counts = rng.integers(1, 1000, size=words.size).astype(np.uint32)
        O.write_list(OUT / f"main_{k}.list", words, counts, k)
        # queries: present words, their reverse complements (same canonical word), absent words
        present = rng.choice(words, size=60)
        absent = rng.integers(0, min(hi, 2 ** 63), size=60, dtype=np.uint64)
        def revcomp(w):
            r, x = 0, ~int(w)
            for _ in range(k):
                r = (r << 2) | (x & 3)
                x >>= 2
            return r
        q = np.concatenate([present[:30], np.array([revcomp(w) for w in present[30:]], dtype=np.uint64), absent,
                            np.array([0, hi - 1], dtype=np.uint64)])
        rng.shuffle(q)
        (OUT / f"queries_{k}.txt").write_text("".join(O.word_to_string(w, k) + "\n" for w in q))
        r = O.run_ref("glistquery", [f"main_{k}.list", "-f", f"queries_{k}.txt"], cwd=OUT, check=True, timeout=20)
        (OUT / f"lookup_{k}.out").write_bytes(r.stdout)
        # zipper: a second list sharing about half of the words
        sub = np.unique(np.concatenate([rng.choice(words, size=150), canonical(absent, k)]))
        sub_counts = rng.integers(1, 50, size=sub.size).astype(np.uint32)
        O.write_list(OUT / f"sub_{k}.list", sub, sub_counts, k)
        r = O.run_ref("glistquery", [f"main_{k}.list", "-l", f"sub_{k}.list"], cwd=OUT, check=True, timeout=20)
        (OUT / f"zipper_{k}.out").write_bytes(r.stdout)
        # whole-tool outputs for the gt4gpu-query CLI: dump, count matrices, multi-list search, frequency window, -stat
        for name, args in ((f"dump_{k}.out", [f"main_{k}.list"]),
                           (f"matrix_{k}.out", [f"main_{k}.list", f"sub_{k}.list"]),
                           (f"matrix_isunion_{k}.out", [f"main_{k}.list", f"sub_{k}.list", "--is_union", "--header"]),
                           (f"multi_{k}.out", [f"main_{k}.list", f"sub_{k}.list", "-l", f"sub_{k}.list"]),
                           (f"lookup_minmax_{k}.out", [f"main_{k}.list", "-f", f"queries_{k}.txt", "-min", "100", "-max", "500"]),
                           (f"stat_{k}.out", [f"main_{k}.list", "-stat"])):
            if k in (5, 16):
                r = O.run_ref("glistquery", args, cwd=OUT, check=True, timeout=20)
                (OUT / name).write_bytes(r.stdout)
        cases.append({"k": k, "n_main": int(words.size), "n_queries": int(q.size), "n_sub": int(sub.size)})
    (OUT / "query_golden.json").write_text(json.dumps({"tool": "glistquery 4.2.16 (oracle/_ref)", "cases": cases}, indent=1) + "\n")
    print("written", OUT)


if __name__ == "__main__":
    main()
