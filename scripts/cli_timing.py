#!/usr/bin/env python
"""Phases of the drop-in CLI's file -> file pipeline at full size (GT4GPU_FILE_TIMING=1): two list files of n k-mers each in
/dev/shm, gt4gpu-compare -u / -i / -d -c 5.  Usage: cli_timing.py [n_per_list] [extra env assignments NAME=VALUE ...]"""
import json, os, subprocess, sys, tempfile, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import bench
import genometester4_b200 as g
from genometester4_b200 import _lib, api, synth
n = float(sys.argv[1]) if len(sys.argv) > 1 else 1e9
extra_env = dict(a.split("=", 1) for a in sys.argv[2:])


def write_lists(td, n, k=25):
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
    return paths, n_in


with tempfile.TemporaryDirectory(dir="/dev/shm") as td:
    td = Path(td)
    paths, n_in = write_lists(td, n)
    print(json.dumps({"input_kmers": n_in, "cpus": os.cpu_count(), "env": extra_env}), flush=True)
    for flags in (["-u"], ["-u"], ["-i"], ["-d", "-c", "5"]):
        t0 = time.perf_counter()
        r = subprocess.run([str(_lib.cli_path()), str(paths[0]), str(paths[1]), *flags, "-o", str(td / "out")], capture_output=True,
                           env={**os.environ, "GT4GPU_FILE_TIMING": "1", **extra_env})
        dt = time.perf_counter() - t0
        assert r.returncode == 0, r.stderr
        sizes = {f.name: f.stat().st_size for f in td.glob("out_25_*.list")}
        print(json.dumps({"flags": " ".join(flags), "wall_s": round(dt, 3), "kmers_per_s": round(n_in / dt), "outputs": sizes}), flush=True)
        print(r.stderr.decode().strip(), flush=True)
        for f in td.glob("out_25_*.list"):
            f.unlink()
