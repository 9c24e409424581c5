"""The N > 1 path on CPU: world_size-2 (and 3) gloo process groups run the key-range sharding plan, the
all-gather of per-rank output counts and the parallel file assembly.  The per-rank merge is injected (the
oracle stands in for the GPU, which this container does not have); everything else is the product code of
genometester4_b200/sharded.py and libgt4gpu's host-side gt4gpu_plan_splitters."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from tests import refrun  # noqa: E402
from tests.util import make_multi, make_pair  # noqa: E402


def _oracle_pair(paths, ranges, k, kwargs, stream):
    from genometester4_b200 import api
    from oracle import oracle as O
    a, b = O.read_list(paths[0]), O.read_list(paths[1])
    sa = O.SList(a.words[ranges[0][0]:ranges[0][1]], a.counts[ranges[0][0]:ranges[0][1]], k)
    sb = O.SList(b.words[ranges[1][0]:ranges[1][1]], b.counts[ranges[1][0]:ranges[1][1]], k)
    rule = kwargs["rule"]
    res = O.compare2(sa, sb, union=kwargs["find_union"], intrsec=kwargs["find_intrsec"], diff=kwargs["find_diff"] or kwargs["find_ddiff"],
                     ddiff=kwargs["find_ddiff"], subtract=kwargs["subtract"], cutoff=kwargs["cutoff"], rule=rule,
                     count_override=kwargs["count_override"])
    out = {}
    for s, r in res.items():
        rec = np.empty(r.n_words, dtype=api.RECORD)
        rec["word"], rec["count"] = r.words, r.counts
        out[s] = rec
    return out


def _oracle_multi(paths, ranges, k, kwargs, stream):
    from genometester4_b200 import api
    from oracle import oracle as O
    lists = []
    for p, (lo, hi) in zip(paths, ranges):
        l = O.read_list(p)
        lists.append(O.SList(l.words[lo:hi], l.counts[lo:hi], l.word_length))
    op = kwargs["op"]
    fn = O.union_multi if op == "union" else O.intersect_multi
    rc, r = fn(lists, cutoff=kwargs["cutoff"], rule=kwargs["rule"], count_override=kwargs["count_override"])
    assert rc == 0
    rec = np.empty(r.n_words, dtype=api.RECORD)
    rec["word"], rec["count"] = r.words, r.counts
    return {("union" if op == "union" else "intrsec"): rec}


def _worker(rank, world, tmp, port):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from genometester4_b200 import sharded
    tmp = Path(tmp)
    tot = sharded.compare_files(tmp / "A.list", tmp / "B.list", str(tmp / "sh"), find_union=1, find_intrsec=1, find_ddiff=1,
                                cutoff=3, rule="default", merge_fn=_oracle_pair)
    co = sharded.compare_files(tmp / "A.list", tmp / "B.list", str(tmp / "co"), find_union=1, countonly=1, merge_fn=_oracle_pair)
    tm = sharded.multi_files([tmp / f"M{j}.list" for j in range(5)], str(tmp / "shm"), op="union", cutoff=2, merge_fn=_oracle_multi)
    ti = sharded.multi_files([tmp / f"M{j}.list" for j in range(5)], str(tmp / "shm"), op="intersect", rule="add", merge_fn=_oracle_multi)
    if rank == 0:
        np.save(tmp / "totals.npy", np.array([tot["union"], tot["intrsec"], tot["diff1"], tot["diff2"], co["union"], tm["union"], ti["intrsec"]], dtype=np.uint64))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_files_equal_single_process(world, tmp_path, oracle):
    from genometester4_b200 import _lib
    if not _lib.lib_path().exists():
        pytest.skip("libgt4gpu.so not built")
    a, b = make_pair(61, 30_000, 20_000, 9_000, 20, "tail")
    oracle.write_list(tmp_path / "A.list", *a, 20)
    oracle.write_list(tmp_path / "B.list", *b, 20)
    multi = make_multi(62, 5, 8_000, 20_000, 20, "tail")
    multi[3] = (multi[3][0][:0], multi[3][1][:0])        # one empty list
    for j, (w, c) in enumerate(multi):
        oracle.write_list(tmp_path / f"M{j}.list", w, c, 20)
    port = 29500 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, str(tmp_path), port), nprocs=world, join=True)

    want = oracle.compare2(oracle.SList(*a, 20), oracle.SList(*b, 20), union=True, intrsec=True, diff=True, ddiff=True, cutoff=3)
    for s, tag in (("union", "union"), ("intrsec", "intrsec"), ("diff1", "0_diff1"), ("diff2", "0_diff2")):
        assert (tmp_path / f"sh_20_{tag}.list").read_bytes() == refrun.list_bytes(want[s], 20), s
    assert not list(tmp_path.glob("co_*")) and not list(tmp_path.glob("*.tmp"))
    ol = [oracle.SList(w, c, 20) for w, c in multi]
    rc, wu = oracle.union_multi(ol, cutoff=2)
    assert (tmp_path / "shm_20_union.list").read_bytes() == refrun.list_bytes(wu, 20)
    rc, wi = oracle.intersect_multi(ol, rule="add")
    assert wi.n_words == 0 and (tmp_path / "shm_20_intrsec.list").read_bytes() == refrun.list_bytes(wi, 20)
    tot = np.load(tmp_path / "totals.npy")
    cu = oracle.compare2(oracle.SList(*a, 20), oracle.SList(*b, 20), union=True)["union"]
    assert tot.tolist() == [[want[s].n_words, want[s].total_count] for s in ("union", "intrsec", "diff1", "diff2")] + \
        [[cu.n_words, cu.total_count], [wu.n_words, wu.total_count], [0, 0]]


def test_splitter_plan_properties():
    from genometester4_b200 import _lib, api
    if not _lib.lib_path().exists():
        pytest.skip("libgt4gpu.so not built")
    rng = np.random.default_rng(5)
    lists = [np.unique(rng.integers(0, 1 << 62, size=n, dtype=np.uint64)) for n in (50_000, 1, 0, 30_000, 7)]
    lists.append(lists[0][::3].copy())                    # many keys shared with list 0
    rec = np.zeros(lists[3].size, dtype=api.RECORD)
    rec["word"] = lists[3]
    arrays = lists[:3] + [rec["word"]] + lists[4:]        # one strided (12-byte) view, like a mapped file
    for parts in (1, 2, 3, 8):
        bounds, split = api.plan_splitters(arrays, parts)
        assert bounds.shape == (len(arrays), parts + 1) and len(split) == parts - 1
        assert (bounds[:, 0] == 0).all() and (bounds[:, -1] == [a.size for a in arrays]).all()
        assert (np.diff(bounds.astype(np.int64), axis=1) >= 0).all()
        total = sum(a.size for a in arrays)
        sizes = np.diff(bounds.astype(np.int64), axis=1).sum(axis=0)
        assert np.abs(sizes - total / parts).max() <= len(arrays) + 1          # balanced up to key multiplicity
        for p in range(1, parts):                                                 # equal keys land in the same part
            for a, brow in zip(arrays, bounds):
                i = int(brow[p])
                assert (i == 0 or a[i - 1] < split[p - 1]) and (i == a.size or a[i] >= split[p - 1])
