#!/usr/bin/env python
"""File -> file comparison of the drop-in CLI with the unmodified reference binary on the same .list files
(page-cache-hot, /dev/shm): wall-clock, output equality.  Usage: cli_e2e.py [n_per_list] [--no-ref]
At sizes where the reference would run for minutes pass --no-ref: only the GPU CLI is timed (pipelined and, for
comparison, the load-all / merge / write-all path it replaced, GT4GPU_NO_FILE_PIPELINE=1)."""
import json, os, subprocess, sys, tempfile, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import bench
import genometester4_b200 as g
from genometester4_b200 import _lib, api, synth
n = float(sys.argv[1]) if len(sys.argv) > 1 else 1e8
use_ref = "--no-ref" not in sys.argv
ref = bench.reference_binary() if use_ref else None


def write_lists(td, n, k=25):
    """the bench's synthetic pair, generated on the GPU and written as list files"""
    g.init(0)
    m, pa, pb = bench.universe_for(n, 0.5)
    (wa, ca), (wb, cb) = synth.pair_torch(42, k, m, 0, m, pa, pb)
    paths = []
    for name, w, c in (("A", wa, ca), ("B", wb, cb)):
        cnt = w.numel()
        p = td / f"sample_{name}.list"
        hdr = np.zeros(1, dtype=[("code", "<u4"), ("major", "<u4"), ("minor", "<u4"), ("k", "<u4"), ("n", "<u8"), ("total", "<u8"), ("start", "<u8"), ("wb", "<u4"), ("cb", "<u4")])
        hdr["code"], hdr["major"], hdr["minor"], hdr["k"] = 0x47543443, 4, 2, k
        hdr["n"], hdr["total"], hdr["start"], hdr["wb"], hdr["cb"] = cnt, int(c.to(torch.int64).sum().item()), 48, 8, 4
        with open(p, "wb") as f:
            f.write(hdr.tobytes())
            step = 1 << 27
            for lo in range(0, cnt, step):
                hi = min(cnt, lo + step)
                dev = torch.empty((hi - lo) * 12, dtype=torch.uint8, device="cuda")
                assert api._lib.load().gt4gpu_interleave(w[lo:hi].data_ptr(), c[lo:hi].data_ptr(), hi - lo, dev.data_ptr()) == 0
                dev.cpu().numpy().tofile(f)
                del dev
        paths.append(p)
    n_in = wa.numel() + wb.numel()
    del wa, ca, wb, cb
    torch.cuda.empty_cache()
    g.shutdown() if hasattr(g, "shutdown") else None
    return paths, n_in


with tempfile.TemporaryDirectory(dir="/dev/shm") as td:
    td = Path(td)
    paths, n_in = write_lists(td, n)
    for flags in (["-u"], ["-i"], ["-d", "-c", "5"], ["-u", "-i", "-d"], ["-u", "--count_only"]):
        row = {"flags": " ".join(flags), "input_kmers": n_in}
        for who, exe, env in (("gt4gpu", _lib.cli_path(), {}), ("gt4gpu_unpipelined", _lib.cli_path(), {"GT4GPU_NO_FILE_PIPELINE": "1"}), ("reference", ref, {})):
            if exe is None:
                continue
            best = 1e9
            for rep in range(2):
                t0 = time.perf_counter()
                r = subprocess.run([str(exe), str(paths[0]), str(paths[1]), *flags, "-o", str(td / who)], capture_output=True, env={**os.environ, **env})
                best = min(best, time.perf_counter() - t0)
                assert r.returncode == 0, r.stderr
            row[who + "_s"] = round(best, 3)
            row[who + "_kmers_per_s"] = round(n_in / best)
            row[who + "_stdout"] = r.stdout.decode()
        same = True
        for f in sorted(td.glob("gt4gpu_25_*.list")):
            for other in ("reference_", "gt4gpu_unpipelined_"):
                o = td / f.name.replace("gt4gpu_", other, 1)
                if o.exists():
                    same &= subprocess.run(["cmp", "-s", str(f), str(o)]).returncode == 0
        row["outputs_identical"] = bool(same) and row.get("gt4gpu_stdout") == row.get("reference_stdout", row.get("gt4gpu_stdout"))
        for k2 in [k2 for k2 in row if k2.endswith("_stdout")]:
            row.pop(k2)
        for f in td.glob("*_25_*.list"):
            f.unlink()
        print(json.dumps(row), flush=True)
