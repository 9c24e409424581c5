#!/usr/bin/env python
"""Generates tests/golden/maker/: small synthetic FastA/FastQ inputs and, for each (input, k), the digest of the
.list file the UNMODIFIED reference glistmaker (oracle/_ref/glistmaker, built by oracle/Makefile from
/root/reference/src) writes for it.  Run in the build container; the fixtures travel to the GPU box."""
import hashlib
import json
import random
import struct
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import oracle as O  # noqa: E402

OUT = Path(__file__).parent / "golden" / "maker"


def inputs():
    rnd = random.Random(2024)

    def seq(n, alphabet="ACGT"):
        return "".join(rnd.choice(alphabet) for _ in range(n))

    def wrap(s, width):
        return "\n".join(s[i:i + width] for i in range(0, len(s), width))

    genome = seq(6000)
    reads = [genome[p:p + 100] for p in (rnd.randrange(0, 5900) for _ in range(150))]
    files = {
        "plain.fa": ">chr1 synthetic\n" + wrap(genome, 60) + "\n",
        "multi.fa": "".join(f">rec{i} len\n{wrap(seq(rnd.randrange(1, 400)), 70)}\n" for i in range(25)),
        "messy.fa": ">s1\n" + wrap(seq(500, "ACGTNacgtn"), 50) + "\r\n>s2 with > inside\n" + seq(300, "ACGTUuRYKM-*") +
                    "\n\n\n>empty\n>s3\nACGT>inline\n" + "G" * 90 + "\n" + seq(64) + ">tail",
        "repeats.fa": ">r\n" + wrap(("ACGTTGCA" * 40 + "AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA") * 6, 80) + "\n",
        "reads.fq": "".join(f"@read{i}/1\n{r if i % 7 else r[:40] + 'N' + r[41:]}\n+\n"
                            f"{''.join(rnd.choice('IJK>@+#') for _ in range(len(r)))}\n" for i, r in enumerate(reads)),
        "short.fa": ">x\nACG\n",
    }
    return files


def main():
    if O.ref_binary("glistmaker") is None:
        raise SystemExit("oracle/_ref/glistmaker missing: run make -C oracle")
    OUT.mkdir(parents=True, exist_ok=True)
    cases = []
    for name, text in inputs().items():
        (OUT / name).write_text(text, newline="")
        for k in (1, 2, 5, 11, 16, 21, 25, 31, 32):
            with tempfile.TemporaryDirectory() as tmp:
                O.run_ref("glistmaker", [str(OUT / name), "-w", str(k), "-o", "g"], cwd=Path(tmp), check=True)
                data = (Path(tmp) / f"g_{k}.list").read_bytes()
            n_words, total = struct.unpack_from("<QQ", data, 16)
            cases.append({"input": name, "k": k, "n_words": n_words, "total_count": total, "bytes": len(data),
                          "sha256": hashlib.sha256(data).hexdigest()})
    (OUT / "maker_golden.json").write_text(json.dumps({"tool": "glistmaker 4.2.16 (oracle/_ref)", "cases": cases}, indent=1) + "\n")
    print(f"{len(cases)} cases written to {OUT}")


if __name__ == "__main__":
    main()
