#!/usr/bin/env python
"""gt4gpu-compare on one device vs --gpus N (one process per key-range shard), file -> file on /dev/shm.
Usage: cli_gpus.py n_per_list n_gpus"""
import json, os, subprocess, sys, tempfile, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from genometester4_b200 import _lib
sys.argv = [sys.argv[0], sys.argv[1], "--no-ref", *sys.argv[2:]]
n = float(sys.argv[1]); n_gpus = int(sys.argv[3])
import importlib.util
spec = importlib.util.spec_from_file_location("cli_e2e_mod", Path(__file__).resolve().parent / "cli_e2e.py")
src = Path(spec.origin).read_text().split("with tempfile.TemporaryDirectory")[0]
ns = {"__name__": "cli_e2e_mod", "__file__": spec.origin}
exec(compile(src, spec.origin, "exec"), ns)
with tempfile.TemporaryDirectory(dir="/dev/shm") as td:
    td = Path(td)
    paths, n_in = ns["write_lists"](td, n)
    for flags in (["-u"], ["-d", "-c", "5"], ["-u", "-i", "-d"]):
        row = {"flags": " ".join(flags), "input_kmers": n_in}
        for who, extra in (("one_gpu", []), (f"{n_gpus}_gpus", ["--gpus", str(n_gpus)])):
            best = 1e9
            for rep in range(2):
                t0 = time.perf_counter()
                r = subprocess.run([str(_lib.cli_path()), str(paths[0]), str(paths[1]), *flags, *extra, "-o", str(td / who)], capture_output=True)
                best = min(best, time.perf_counter() - t0)
                assert r.returncode == 0, r.stderr
            row[who + "_s"] = round(best, 3)
        same = True
        for f in sorted(td.glob("one_gpu_25_*.list")):
            same &= subprocess.run(["cmp", "-s", str(f), str(td / f.name.replace("one_gpu", f"{n_gpus}_gpus", 1))]).returncode == 0
        row["outputs_identical"] = bool(same)
        for f in td.glob("*_25_*.list"):
            f.unlink()
        print(json.dumps(row), flush=True)
