#!/usr/bin/env python
"""File -> file comparison of the drop-in CLI with the unmodified reference binary on the same .list files
(page-cache-hot, /dev/shm): wall-clock, output equality.  Usage: cli_e2e.py [n_per_list]"""
import json, subprocess, sys, tempfile, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
from genometester4_b200 import _lib
n = float(sys.argv[1]) if len(sys.argv) > 1 else 1e8
ref = bench.reference_binary()
with tempfile.TemporaryDirectory(dir="/dev/shm") as td:
    td = Path(td)
    paths, n_in = bench.build_sample_lists(td, n, 0.5, 25)
    for flags in (["-u"], ["-i"], ["-d", "-c", "5"], ["-u", "-i", "-d"], ["-u", "--count_only"]):
        row = {"flags": " ".join(flags), "input_kmers": n_in}
        for who, exe in (("gt4gpu", _lib.cli_path()), ("reference", ref)):
            if exe is None:
                continue
            best = 1e9
            for rep in range(2):
                t0 = time.perf_counter()
                r = subprocess.run([str(exe), str(paths[0]), str(paths[1]), *flags, "-o", str(td / who)], capture_output=True)
                best = min(best, time.perf_counter() - t0)
                assert r.returncode == 0, r.stderr
            row[who + "_s"] = round(best, 3)
            row[who + "_kmers_per_s"] = round(n_in / best)
            row[who + "_stdout"] = r.stdout.decode()
        same = True
        for f in sorted(td.glob("gt4gpu_*.list")):
            g = td / f.name.replace("gt4gpu_", "reference_", 1)
            if ref is not None:
                same &= subprocess.run(["cmp", "-s", str(f), str(g)]).returncode == 0
        row["outputs_identical"] = bool(same) and row.get("gt4gpu_stdout") == row.get("reference_stdout", row.get("gt4gpu_stdout"))
        row.pop("gt4gpu_stdout", None); row.pop("reference_stdout", None)
        for f in td.glob("*_25_*.list"):
            f.unlink()
        print(json.dumps(row), flush=True)
