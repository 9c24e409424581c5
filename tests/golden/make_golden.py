#!/usr/bin/env python
"""Regenerate tests/golden/golden.json by running the UNMODIFIED reference
glistcompare (oracle/_ref/glistcompare, built from /root/reference/src by
oracle/Makefile) over the golden subset of tests/cases.py.

Stored per case: the sha256, n_words and total_count of every output file the
reference produced, and the stdout of the matching --count_only run.  The input
lists are regenerated deterministically from tests/cases.py (numpy PCG64, fixed
seeds) and their sha256 is stored too, so a drifting generator is detected.

Usage (in the build container, where /root/reference exists):
    python tests/golden/make_golden.py
"""
import json
import struct
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import oracle as O  # noqa: E402
from tests import cases, refrun  # noqa: E402


def describe(files):
    out = {}
    for name, b in files.items():
        n_words, total = struct.unpack_from("<QQ", b, 16)
        out[name] = {"sha256": refrun.digest(b), "n_words": n_words, "total_count": total, "size": len(b)}
    return out


def main():
    O.build()
    assert O.ref_binary("glistcompare") is not None, "reference binary missing"
    golden = {"reference": "GenomeTester4 4.2.16 glistcompare", "inputs": {}, "pair": [], "multi": []}
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        for which, gen in (("pair", cases.pair_cases), ("multi", cases.multi_cases)):
            paths = refrun.write_inputs(td / "in", which)
            for name, ps in paths.items():
                golden["inputs"][name] = [refrun.digest(p.read_bytes()) for p in ps]
            for name, ops, rule, cutoff in gen(full=False):
                rc, files, _ = refrun.run_reference(td / "run", paths[name], ops, rule, cutoff)
                rc2, _, stdout = refrun.run_reference(td / "run", paths[name], ops, rule, cutoff, count_only=True)
                golden[which].append({"input": name, "ops": list(ops), "rule": rule, "cutoff": cutoff,
                                      "rc": rc, "files": describe(files), "count_only_stdout": stdout})
    out = Path(__file__).with_name("golden.json")
    out.write_text(json.dumps(golden, indent=0, sort_keys=True) + "\n")
    print(f"wrote {out}: {len(golden['pair'])} pair cases, {len(golden['multi'])} multi cases")


if __name__ == "__main__":
    main()
