"""GPU parity tests of the list-building back end: gt4gpu_count_words (radix sort +
run-length counts) through the C ABI against the oracle's restatement of wordtable_sort + merge_tables_to_file."""
import os

import numpy as np
import pytest


def rand_bytes(rng, alphabet: bytes, n) -> bytes:
    """n random characters of `alphabet` (one byte each: bytes() of an int64 array would be its raw 8-byte image)."""
    return rng.choice(np.frombuffer(alphabet, dtype=np.uint8), size=int(n)).tobytes()

pytestmark = pytest.mark.gpu

SORT_TILE = 512 * 16     # keys per CTA of radix_onesweep_kernel
RLE_TILE = 128 * 16      # keys per CTA of rle_heads_kernel


@pytest.fixture(scope="module")
def g():
    import genometester4_b200 as g
    g.init(0)
    return g


def check(g, oracle, words, k):
    res = g.count_words(words, k)
    exp = oracle.count_words(words, k)
    w, c = res.to_host()
    assert res.n_words == len(exp.words)
    assert res.total_count == len(words)
    assert np.array_equal(w, exp.words)
    assert np.array_equal(c, exp.counts)
    res.free()


def random_words(rng, n, k, distinct=None):
    hi = 4 ** k
    if distinct is None:
        return rng.integers(0, hi, size=n, dtype=np.uint64) if hi < 2 ** 63 else \
            rng.integers(0, 2 ** 63, size=n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=n, dtype=np.uint64)
    pool = random_words(rng, distinct, k)
    return pool[rng.integers(0, distinct, size=n)]


@pytest.mark.parametrize("k", [1, 4, 11, 16, 25, 31, 32])
def test_count_words_random(g, oracle, k):
    rng = np.random.default_rng(100 + k)
    for n in (1, 2, 33, SORT_TILE - 1, SORT_TILE, SORT_TILE + 1, 3 * RLE_TILE + 7, 100_003):
        check(g, oracle, random_words(rng, n, k), k)


def test_count_words_duplicates_and_order(g, oracle):
    rng = np.random.default_rng(7)
    k = 25
    n = 5 * SORT_TILE + 123
    check(g, oracle, np.full(n, 12345, dtype=np.uint64), k)                    # one run spanning many tiles
    check(g, oracle, random_words(rng, n, k, distinct=3), k)
    check(g, oracle, random_words(rng, n, k, distinct=1000), k)
    w = np.sort(random_words(rng, n, k))
    check(g, oracle, w, k)
    check(g, oracle, w[::-1].copy(), k)
    # runs that end exactly on tile boundaries of the run-length kernel
    w = np.repeat(np.arange(10, dtype=np.uint64) * np.uint64(977), RLE_TILE)
    check(g, oracle, rng.permutation(w), k)
    # extreme words: 0 and 4^32 - 1
    w = np.array([0, 2 ** 64 - 1, 0, 2 ** 64 - 1, 5, 2 ** 63, 2 ** 63], dtype=np.uint64)
    check(g, oracle, w, 32)


def test_count_words_empty_and_bad_args(g):
    res = g.count_words(np.zeros(0, dtype=np.uint64), 16)
    assert res.n_words == 0 and res.total_count == 0
    for k in (0, 33):
        with pytest.raises(g.GT4GPUError) as e:
            g.count_words(np.zeros(4, dtype=np.uint64), k)
        assert e.value.code == 1


def test_count_words_device_input_medium(g, oracle):
    """2e7 words already in HBM: compared with torch.unique (checker only) and, on a prefix, with the oracle."""
    import torch
    k = 25
    gen = torch.Generator(device="cuda").manual_seed(5)
    w = torch.randint(0, 4 ** 12, (20_000_000,), generator=gen, device="cuda", dtype=torch.int64)   # ~1.2 words per key
    torch.cuda.synchronize()
    res = g.count_words(w.data_ptr(), k, n_words=w.numel())
    uw, uc = torch.unique(w, return_counts=True)
    tw, tc = res.as_torch()
    assert res.n_words == uw.numel() and res.total_count == w.numel()
    assert torch.equal(tw, uw) and torch.equal(tc.to(torch.int64), uc)
    res.free()
    check(g, oracle, w[:300_000].cpu().numpy().astype(np.uint64), k)


def test_listmaker_pipeline_matches_glistmaker_golden(g, oracle):
    """sequence file -> words (host reader) -> tables -> count_words -> union_multi == the list glistmaker wrote."""
    import json
    from pathlib import Path
    gold_dir = Path(__file__).parent / "golden" / "maker"
    gold = json.loads((gold_dir / "maker_golden.json").read_text())
    for case in gold["cases"]:
        text = (gold_dir / case["input"]).read_bytes()
        k = case["k"]
        words = g.sequence_words(text, k)
        # several small tables, collated like collate_files (gt4_write_union, cutoff 1)
        n_tables = 3
        parts = [words[i::n_tables] for i in range(n_tables)]
        tables = [g.count_words(p, k) for p in parts if len(p)]
        if not tables:
            assert case["n_words"] == 0
            continue
        lists = [g.WordList.from_device(*t.device_ptrs, t.n_words, k, keepalive=t) for t in tables]
        res = g.union_multi(lists, cutoff=1)
        assert res.n_words == case["n_words"] and res.total_count == case["total_count"]
        import hashlib
        assert hashlib.sha256(res.list_bytes()).hexdigest() == case["sha256"], case


def test_listmaker_cli_against_glistmaker_golden(tmp_path):
    """gt4gpu-listmaker leaves byte for byte the <out>_<k>.list the unmodified glistmaker wrote (committed digests);
    small --table_size values force many GPU tables and the final collation."""
    import hashlib
    import json
    import subprocess
    from pathlib import Path
    from genometester4_b200 import _lib
    gold_dir = Path(__file__).parent / "golden" / "maker"
    gold = json.loads((gold_dir / "maker_golden.json").read_text())
    for i, case in enumerate(gold["cases"]):
        if i % 3 and case["k"] not in (1, 32):       # every process start pays the CUDA context creation: keep the list short
            continue
        # "--table_size N" swallows the token after N as well (src/glistmaker.c:214), hence the filler
        extra = [[], ["--table_size", "1000", "filler"], ["--table_size", "37", "filler", "-D"]][(i // 3) % 3]
        # FastA text goes to the GPU reader in blocks that end where a record ends: tiny blocks exercise the splitting
        env = {**os.environ, "GT4GPU_FASTA_BLOCK": "700"} if i % 2 else None
        r = subprocess.run([str(_lib.listmaker_cli_path()), str(gold_dir / case["input"]), "-w", str(case["k"]), "-o", "t", *extra],
                           cwd=tmp_path, capture_output=True, env=env)
        assert r.returncode == 0, (case, r.stderr)
        data = (tmp_path / f"t_{case['k']}.list").read_bytes()
        assert len(data) == case["bytes"] and hashlib.sha256(data).hexdigest() == case["sha256"], (case, extra)
        assert not (tmp_path / f"t_{case['k']}.list.tmp").exists()
        (tmp_path / f"t_{case['k']}.list").unlink()
    # several input files = one list of their words together
    r = subprocess.run([str(_lib.listmaker_cli_path()), str(gold_dir / "plain.fa"), str(gold_dir / "multi.fa"), str(gold_dir / "reads.fq"),
                        "-w", "16", "-o", "all", "--table_size", "5000", "x"], cwd=tmp_path, capture_output=True)
    assert r.returncode == 0, r.stderr
    from oracle import oracle as O
    words = np.concatenate([O.sequence_words((gold_dir / f).read_bytes(), 16) for f in ("plain.fa", "multi.fa", "reads.fq")])
    exp = O.count_words(words, 16)
    assert (tmp_path / "all_16.list").read_bytes() == O.header_bytes(16, len(exp.words), int(exp.counts.sum())) + exp.records().tobytes()


def test_device_fasta_reader_matches_host_reader(g, oracle):
    """gt4gpu_fasta_words_device (parallel line-state + compaction kernels) == the byte-serial reader, word for word."""
    import json
    from pathlib import Path
    gold_dir = Path(__file__).parent / "golden" / "maker"
    texts = [(gold_dir / f).read_bytes() for f in ("plain.fa", "multi.fa", "messy.fa", "repeats.fa", "short.fa")]
    texts += [b">x", b">x\n", b">x\nACGT", b">x\nAC\nGT\n>y\nTTTT", b">a\nACGTN\x00ACGT", b">a\r\nAC\r\nGT\r\n", b">a\nacgu\n",
              b">a>b\nAC>GT\nACGT\n", b">\n" + b"ACGT" * 3000 + b"\n", b">n\n" + b"\n" * 5000 + b"ACGTACGT",
              b">long name " + b"x" * 9000 + b"\nACGTTGCA\n>" + b">" * 5000 + b"\nGGGGCCCC"]
    rng = np.random.default_rng(21)
    for _ in range(6):            # random images whose names, lines and N runs straddle the 4 KiB chunks in every way
        parts = []
        for r in range(int(rng.integers(1, 60))):
            name = rand_bytes(rng, b"abc >XYZ", rng.integers(0, 300))
            seq = rand_bytes(rng, b"ACGTACGTACGTNacgtn\n\n\r -", rng.integers(0, 9000))
            parts.append(b">" + name.replace(b"\n", b"") + b"\n" + seq)
            if rng.random() < 0.2:
                parts.append(b">inline" * int(rng.integers(1, 4)))
        texts.append(b"".join(parts))
    for text in texts:
        for k in (1, 2, 7, 16, 25, 31, 32):
            dev = g.fasta_words_device(text, k)
            host = oracle.sequence_words(text, k)
            assert dev.n_words == host.size, (text[:60], k)
            assert np.array_equal(dev.to_host(), host), (text[:60], k)
            dev.free()
    for bad in (b"x", b"ACGT\n"):
        with pytest.raises(g.GT4GPUError) as e:
            g.fasta_words_device(bad, 4)
        assert e.value.code == 3
    assert g.fasta_words_device(b"", 4).n_words == 0


def test_device_fasta_to_list_medium(g, oracle):
    """2 Mbp random genome with N runs: device reader -> count_words (device words) == oracle list."""
    rng = np.random.default_rng(8)
    seq = rng.choice(list(b"ACGT"), size=2_000_000).astype(np.uint8)
    seq[rng.integers(0, seq.size, size=200)] = ord("N")
    lines = b"\n".join(seq[i:i + 70].tobytes() for i in range(0, seq.size, 70))
    text = b">chr\n" + lines + b"\n"
    k = 21
    dev = g.fasta_words_device(text, k)
    res = g.count_words(dev.ptr, k, n_words=dev.n_words)
    exp = oracle.count_words(oracle.sequence_words(text, k), k)
    w, c = res.to_host()
    assert np.array_equal(w, exp.words) and np.array_equal(c, exp.counts)


def test_device_fastq_reader_matches_host_reader(g, oracle):
    """Four-line FastQ records on the GPU (state = line number mod 4) == the byte-serial reader; every image the
    reference's reader gives up on is refused with GT4GPU_ERR_FORMAT instead."""
    from pathlib import Path
    gold_dir = Path(__file__).parent / "golden" / "maker"
    rng = np.random.default_rng(31)

    def record(i, n, crlf=False):
        seq = rand_bytes(rng, b"ACGTACGTACGTNacgt", n)
        qual = rand_bytes(rng, b"IJK>@+#!~", n)
        nl = b"\r\n" if crlf else b"\n"
        return b"@read%d some text" % i + nl + seq + nl + b"+" + (b"read%d" % i if i % 3 == 0 else b"") + nl + qual + nl

    good = [(gold_dir / "reads.fq").read_bytes(),
            b"@r\nACGT\n+\nIIII\n", b"@r\nACGT\n+\nIIII", b"@r\nACGT\n+\n", b"@r\nACGT", b"@r\n", b"@r",
            b"@r\nAC>GT@AC\n+r\n>>@@\n@s\nGT\n+\nII\n",
            b"".join(record(i, int(rng.integers(1, 400))) for i in range(300)),
            b"".join(record(i, int(rng.integers(3000, 9000))) for i in range(12)),           # lines longer than a chunk
            b"".join(record(i, 150, crlf=True) for i in range(200)),
            b"".join(record(i, 100) for i in range(100))[:-37]]                              # stops inside a quality line
    for text in good:
        for k in (1, 4, 16, 25, 32):
            host = oracle.sequence_words(text, k)
            dev = g.fasta_words_device(text, k)
            assert dev.n_words == host.size and np.array_equal(dev.to_host(), host), (text[:40], k)
            dev.free()
    bad = [b"@r\nACGT\nIIII\n", b"@r\nACGT\n+\nIIII\nX", b"@r\nACGT\n", b"@r\nACGT\n+abc", b"@r\nACGT\n+\nIIII\n\n",
           b"".join(record(i, 80) for i in range(50)) + b"ACGT\n" + record(99, 80)]
    for text in bad:
        with pytest.raises(ValueError):
            oracle.sequence_words(text, 2)
        with pytest.raises(g.GT4GPUError) as e:
            g.fasta_words_device(text, 2)
        assert e.value.code == 3, text[:40]


def test_device_readers_fuzz_against_oracle(g, oracle):
    """Random images over a hostile alphabet: the device readers accept exactly what the reference's state machine
    accepts (FastA always once the start tag is right; FastQ unless a record is malformed) and agree on every word."""
    rng = np.random.default_rng(1618)
    alphabet = np.frombuffer(b"ACGTacgtNn>@+\n\n\n\r \x00IJ-", dtype=np.uint8)
    n_ok = n_err = 0
    for case in range(2500):
        n = int(rng.integers(0, 200)) if case % 5 else int(rng.integers(4000, 9000))      # some span several chunks
        text = (b">" if case % 2 else b"@") + rng.choice(alphabet, size=n).tobytes()
        k = int(rng.integers(1, 9))
        try:
            want = oracle.sequence_words(text, k)
        except ValueError:
            want = None
        try:
            dev = g.fasta_words_device(text, k)
            got = dev.to_host()
            dev.free()
        except g.GT4GPUError as e:
            assert e.code == 3
            got = None
        assert (want is None) == (got is None), (text[:80], k, len(text))
        if want is not None:
            assert np.array_equal(want, got), (text[:80], k)
            n_ok += 1
        else:
            n_err += 1
    assert n_ok > 500 and n_err > 300
