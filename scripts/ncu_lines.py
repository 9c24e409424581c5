#!/usr/bin/env python
"""Summarise an ncu report per CUDA source line: share of executed instructions, stall samples and the
dominant stall reasons.  Usage: python scripts/ncu_lines.py report.ncu-rep [min_pct]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.7
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur_file, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        func = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        ix = {}
        for i, h in enumerate(hdr):
            ix.setdefault(h, i)
        continue
    if hdr is None or len(r) < len(hdr) or r[0] == "":
        continue
    try:
        inst = int(r[ix["Instructions Executed"]])
        samp = int(r[ix["# Samples"]])
    except ValueError:
        continue
    stalls = {h[6:]: int(r[i]) for h, i in ix.items() if h.startswith("stall_") and "Not Issued" not in h and r[i].isdigit()}
    shw = r[ix["L1 Wavefronts Shared"]], r[ix["L1 Wavefronts Shared Ideal"]]
    out.append((cur_file, r[0], r[1].strip()[:90], inst, samp, stalls, shw))
ti = sum(o[3] for o in out) or 1
ts = sum(o[4] for o in out) or 1
print(f"total warp-instructions {ti}, samples {ts}")
for f, ln, src, inst, samp, stalls, shw in out:
    if inst / ti * 100 >= min_pct or samp / ts * 100 >= min_pct:
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:3]
        tops = " ".join(f"{k}={v / max(samp, 1) * 100:.0f}%" for k, v in top if v)
        print(f"{f[:18]:18s}:{ln:>4} inst {inst / ti * 100:5.1f}% samp {samp / ts * 100:5.1f}% smem {shw[0]:>10}/{shw[1]:>10} [{tops}] {src}")
