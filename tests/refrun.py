"""Run a flag-matrix case through (a) the unmodified reference binary, (b) the oracle,
and build the exact bytes each output file must have."""
from __future__ import annotations

import hashlib
import subprocess
from pathlib import Path

from oracle import oracle as O
from tests import cases

PAIR_FILES = {"union": "union", "intrsec": "intrsec", "diff1": "0_diff1", "diff2": "0_diff2"}


def write_inputs(dirpath: Path, which: str):
    """Materialise the input sets as .list files; returns name -> [paths]."""
    dirpath.mkdir(parents=True, exist_ok=True)
    paths = {}
    if which == "pair":
        for name, (k, a, b) in cases.pair_inputs().items():
            pa, pb = dirpath / f"{name}_A.list", dirpath / f"{name}_B.list"
            O.write_list(pa, a[0], a[1], k)
            O.write_list(pb, b[0], b[1], k)
            paths[name] = [pa, pb]
    else:
        for name, (k, lists) in cases.multi_inputs().items():
            ps = []
            for j, (w, c) in enumerate(lists):
                p = dirpath / f"{name}_{j}.list"
                O.write_list(p, w, c, k)
                ps.append(p)
            paths[name] = ps
    return paths


def cli_args(paths, ops, rule, cutoff, count_only=False, extra=()):
    args = [str(p) for p in paths] + list(ops) + ["-c", str(cutoff)]
    if rule != "default":
        args += ["-r", rule]
    if count_only:
        args.append("--count_only")
    return args + list(extra)


def list_bytes(res: "O.Result", word_length: int) -> bytes:
    return O.header_bytes(word_length, res.n_words, res.total_count) + res.records().tobytes()


def oracle_pair(name, ops, rule, cutoff):
    k, a, b = cases.pair_inputs()[name]
    la, lb = O.SList(a[0], a[1], k), O.SList(b[0], b[1], k)
    res = O.compare2(la, lb, cutoff=cutoff, **cases.ops_to_kwargs(ops), **cases.rule_to_kwargs(rule))
    files = {f"out_{k}_{PAIR_FILES[key]}.list": list_bytes(r, k) for key, r in res.items()}
    stdout = "".join(f"NUnique\t{res[key].n_words}\nNTotal\t{res[key].total_count}\n"
                     for key in ("union", "intrsec", "diff1", "diff2") if key in res)
    return files, stdout


def oracle_multi(name, ops, rule, cutoff):
    """Mirrors main()'s N-list dispatch (/root/reference/src/glistcompare.c:366-422): a
    rejected rule unlinks that output but the run continues."""
    k, lists = cases.multi_inputs()[name]
    sl = [O.SList(w, c, k) for w, c in lists]
    files, stdout, rc_last = {}, "", 0
    kw = cases.rule_to_kwargs(rule)
    if "-u" in ops:
        rc_last, r = O.union_multi(sl, cutoff=cutoff, **kw)
        if rc_last == 0:
            files[f"out_{k}_union.list"] = list_bytes(r, r.word_length)
            stdout += f"NUnique\t{r.n_words}\nNTotal\t{r.total_count}\n"
        else:
            # the reference prints its never-initialised stack header here (:394); zeros in practice
            stdout += "NUnique\t0\nNTotal\t0\n"
    if "-i" in ops:
        rc_last, r = O.intersect_multi(sl, cutoff=cutoff, **kw)
        if rc_last == 0:
            files[f"out_{k}_intrsec.list"] = list_bytes(r, r.word_length)
            stdout += f"NUnique\t{r.n_words}\nNTotal\t{r.total_count}\n"
        else:
            stdout += "NUnique\t0\nNTotal\t0\n"
    return files, stdout, rc_last


def run_reference(tmp: Path, paths, ops, rule, cutoff, count_only=False, extra=()):
    """Returns (returncode, {filename: bytes}, stdout) or None when oracle/_ref is missing."""
    tmp.mkdir(parents=True, exist_ok=True)
    for f in tmp.glob("out_*"):
        f.unlink()
    cp = O.run_ref("glistcompare", cli_args(paths, ops, rule, cutoff, count_only, extra), cwd=tmp)
    if cp is None:
        return None
    files = {f.name: f.read_bytes() for f in sorted(tmp.glob("out_*"))}
    return cp.returncode, files, cp.stdout.decode()


def digest(b: bytes) -> str:
    return hashlib.sha256(b).hexdigest()


def build_config1_lists(tmp: Path):
    """BASELINE config 1: a 1.1 Mbp uniform random genome (python random, seed 11) and a copy
    with 1 % point substitutions, turned into k = 16 lists by the UNMODIFIED reference glistmaker.  Returns the two
    list paths, or None when oracle/_ref is not available."""
    import random
    if O.ref_binary("glistmaker") is None:
        return None
    tmp.mkdir(parents=True, exist_ok=True)
    rnd = random.Random(11)
    g1 = "".join(rnd.choice("ACGT") for _ in range(1_100_000))
    r2 = random.Random(12)
    g2 = "".join(r2.choice("ACGT") if r2.random() < 0.01 else c for c in g1)
    out = []
    for name, s in (("g1", g1), ("g2", g2)):
        (tmp / f"{name}.fa").write_text(f">{name}\n" + "\n".join(s[i:i + 80] for i in range(0, len(s), 80)) + "\n")
        try:
            O.run_ref("glistmaker", [f"{name}.fa", "-w", "16", "-o", name], cwd=tmp, check=True)
        except subprocess.TimeoutExpired:      # the reference's own shutdown race (see oracle.run_ref): no lists, callers skip
            return None
        out.append(tmp / f"{name}_16.list")
    return out
