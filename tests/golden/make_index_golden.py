#!/usr/bin/env python
"""GT4I index inputs: build two small indices and one list with the UNMODIFIED reference glistmaker
(oracle/_ref/glistmaker --index) from deterministic random FASTA, run the reference glistcompare on them and store
the fixtures (tests/golden/index/*.index, *.list) plus the sha256 of every output (index_golden.json)."""
import hashlib, json, random, shutil, subprocess, sys, tempfile
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import oracle as O
O.build()
out = Path(__file__).parent / "index"
out.mkdir(exist_ok=True)
rnd = random.Random(11)
base = "".join(rnd.choice("ACGT") for _ in range(2500))
def mutate(s, rate, seed):
    r = random.Random(seed)
    return "".join(r.choice("ACGT") if r.random() < rate else c for c in s)
seqs = {"x": base + base[100:400], "y": mutate(base, 0.02, 1), "z": mutate(base, 0.05, 2) + base[:300]}
cases = [(["x_16.index", "y_16.list"], ["-u", "-i", "-dd"], "default", 1), (["x_16.index", "y_16.index"], ["-u"], "add", 1),
         (["y_16.list", "x_16.index"], ["-d", "-c", "2"], "default", 2), (["x_16.index", "y_16.index", "z_16.list"], ["-u", "-i"], "default", 1),
         (["x_16.index", "y_16.index"], ["-i", "-r", "max", "-c", "2"], "max", 2)]
golden = []
with tempfile.TemporaryDirectory() as td:
    td = Path(td)
    for name, s in seqs.items():
        (td / f"{name}.fa").write_text(f">{name}\n" + "\n".join(s[i:i + 70] for i in range(0, len(s), 70)) + "\n")
        O.run_ref("glistmaker", [f"{name}.fa", "-w", "16", "--index", "-o", name], cwd=td, check=True)
        O.run_ref("glistmaker", [f"{name}.fa", "-w", "16", "-o", name], cwd=td, check=True)
    for f in ("x_16.index", "y_16.index", "y_16.list", "z_16.list"):
        shutil.copy(td / f, out / f)
    for files, flags, _, _ in cases:
        for f in td.glob("out_*"):
            f.unlink()
        cp = O.run_ref("glistcompare", files + flags, cwd=td)
        res = {f.name: hashlib.sha256(f.read_bytes()).hexdigest() for f in sorted(td.glob("out_*"))}
        co = O.run_ref("glistcompare", files + flags + ["--count_only"], cwd=td)
        golden.append({"files": files, "flags": flags, "rc": cp.returncode, "outputs": res, "count_only_stdout": co.stdout.decode()})
(Path(__file__).parent / "index_golden.json").write_text(json.dumps(golden, indent=1) + "\n")
print("wrote", len(golden), "cases;", [(f.name, f.stat().st_size) for f in sorted(out.iterdir())])
