"""Key-range sharded set operations over list FILES, one process per GPU.

    splitters (host, exact co-rank on key values)  ->  every rank loads its record range of every list
    ->  independent merge on its GPU  ->  all-gather of {n_out, sum} per output  ->  exclusive scan
    ->  every rank pwrites its slice at 48 + 12 * offset, rank 0 writes the header and renames.

No payload crosses GPUs; the only collective is the all-gather of two integers per output stream
(``torch.distributed``: NCCL on GPUs, gloo in the CPU tests).  Output files are byte-identical to what the
reference's single-process run writes (`<o>_<k>_union.list` ..., /root/reference/src/glistcompare.c:816-831).

``merge_fn`` exists so that the sharding logic can be exercised without a GPU (the tests inject the oracle);
the default is the CUDA library and nothing else.
"""
from __future__ import annotations

import ctypes as C
import os
import time
from pathlib import Path

import numpy as np

from . import _lib, api

STREAM_TAGS = {"union": "union", "intrsec": "intrsec", "diff1": "0_diff1", "diff2": "0_diff2"}


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


def read_header(path, stream: bool = False) -> _lib.Header:
    h = _lib.Header()
    api._check(_lib.load().gt4gpu_list_read_header(os.fsencode(path), int(stream), C.byref(h)))
    return h


INDEX_RECORD = np.dtype([("word", "<u8"), ("loc", "<u8")])      # GT4I k-mer table, /root/reference/src/index-map.c:122-139


def map_records(path, header) -> np.ndarray:
    """The file's records as a (read-only) numpy view: 12-byte list records, or 16-byte GT4I index records
    (gt4gpu_list_read_header marks those with count_bytes == 8).  Only the keys are used (splitter planning)."""
    dtype = INDEX_RECORD if header.count_bytes == 8 else api.RECORD
    if header.n_words == 0:
        return np.zeros(0, dtype=dtype)
    return np.memmap(path, dtype=dtype, mode="r", offset=header.list_start, shape=(header.n_words,))


def plan(paths, n_parts: int, stream: bool = False):
    """bounds[list, part] record indices so that every part holds ~1/n_parts of all records and equal keys share a part."""
    headers = [read_header(p, stream) for p in paths]
    maps = [map_records(p, h) for p, h in zip(paths, headers)]
    bounds, splitters = api.plan_splitters(maps, n_parts)
    return headers, bounds, splitters


def _gpu_merge_pair(paths, ranges, k, kwargs, stream, phases=None):
    la = api.WordList.open(paths[0], stream=stream, first=ranges[0][0], count=ranges[0][1] - ranges[0][0])
    lb = api.WordList.open(paths[1], stream=stream, first=ranges[1][0], count=ranges[1][1] - ranges[1][0])
    if phases is not None:
        phases.mark("load")
    return api.compare_wordmaps(la, lb, **kwargs)


def _gpu_merge_multi(paths, ranges, k, kwargs, stream, phases=None):
    lists = [api.WordList.open(p, stream=stream, first=r[0], count=r[1] - r[0]) for p, r in zip(paths, ranges)]
    if phases is not None:
        phases.mark("load")
    op = kwargs.pop("op")
    fn = api.union_multi if op == "union" else api.intersect_multi
    return {("union" if op == "union" else "intrsec"): fn(lists, **kwargs)}


def _exchange(local, streams):
    """all-gather {n_out, sum} of every stream -> (my record offset, global n, global sum) per stream."""
    import torch
    dist, rank, world = _dist()
    mine = torch.tensor([[local[s][0], local[s][1]] for s in streams], dtype=torch.int64).reshape(-1, 2)
    if world == 1:
        return {s: (0, int(mine[i, 0]), int(mine[i, 1])) for i, s in enumerate(streams)}
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    mine = mine.to(dev)
    allv = torch.zeros(world, len(streams), 2, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(allv.view(-1), mine.view(-1))
    allv = allv.cpu()
    out = {}
    for i, s in enumerate(streams):
        n = allv[:, i, 0]
        out[s] = (int(n[:rank].sum()), int(n.sum()), int(allv[:, i, 1].sum()))
    return out


def _barrier():
    dist, _, world = _dist()
    if world > 1:
        dist.barrier()


def _assemble(out_prefix, k, streams, results, totals, countonly, mode):
    """Every rank writes its slice of every output; rank 0 creates the files, writes the headers and renames."""
    _, rank, _ = _dist()
    names = {s: (f"{out_prefix}_{k}_{STREAM_TAGS[s]}.list.tmp", f"{out_prefix}_{k}_{STREAM_TAGS[s]}.list") for s in streams}
    if countonly:
        return
    if rank == 0:
        for s in streams:
            fd = os.open(names[s][0], os.O_WRONLY | os.O_CREAT | os.O_TRUNC, mode)
            h = _lib.Header()
            _lib.load().gt4gpu_header_init(C.byref(h), k)
            h.n_words, h.total_count = totals[s][1], totals[s][2]
            os.pwrite(fd, bytes(h), 0)
            os.ftruncate(fd, 48 + 12 * totals[s][1])
            os.close(fd)
    _barrier()
    for s in streams:
        r = results[s]
        fd = os.open(names[s][0], os.O_WRONLY)
        if isinstance(r, api.Result):
            if r.n_words:
                r.write_records_at(fd, totals[s][0])
        else:                       # numpy records from an injected merge_fn
            os.pwrite(fd, np.ascontiguousarray(r, dtype=api.RECORD).tobytes(), 48 + 12 * totals[s][0])
        os.close(fd)
    _barrier()
    if rank == 0:
        for s in streams:
            os.rename(*names[s])
    _barrier()


def _local_totals(results):
    out = {}
    for s, r in results.items():
        if isinstance(r, api.Result):
            out[s] = (r.n_words, r.total_count)
        else:
            out[s] = (len(r), int(np.asarray(r["count"], dtype=np.uint64).sum()))
    return out


class _Phases:
    """Wall-clock seconds of the phases of one sharded call on this rank (plan / load / merge / exchange / write)."""

    def __init__(self, sink):
        self.sink, self.t = sink, time.perf_counter()

    def mark(self, name):
        if self.sink is not None:
            now = time.perf_counter()
            self.sink[name] = self.sink.get(name, 0.0) + (now - self.t)
            self.t = now


def compare_files(path_a, path_b, out_prefix="out", *, find_union=0, find_intrsec=0, find_diff=0, find_ddiff=0, subtract=0,
                  countonly=0, cutoff=1, rule=api.RULE_DEFAULT, count_override=1, stream=False, merge_fn=None, mode=0o666,
                  timings=None):
    """Sharded `glistcompare A B ...` (two lists).  Returns {stream: (n_words, total_count)} (global).
    timings: optional dict that receives this rank's seconds per phase."""
    _, rank, world = _dist()
    ph = _Phases(timings)
    paths = [Path(path_a), Path(path_b)]
    headers, bounds, _ = plan(paths, world, stream)
    ph.mark("plan")
    if headers[0].word_length != headers[1].word_length:
        raise ValueError(f"File {path_b} has different word length ({headers[1].word_length} != {headers[0].word_length})")
    k = headers[0].word_length
    ranges = [(int(bounds[j, rank]), int(bounds[j, rank + 1])) for j in range(2)]
    kwargs = dict(find_union=find_union, find_intrsec=find_intrsec, find_diff=find_diff, find_ddiff=find_ddiff, subtract=subtract,
                  countonly=countonly, cutoff=cutoff, rule=rule, count_override=count_override)
    results = _gpu_merge_pair(paths, ranges, k, kwargs, stream, ph) if merge_fn is None else merge_fn(paths, ranges, k, kwargs, stream)
    ph.mark("merge")
    streams = [s for s in api.STREAM_NAMES if s in results]
    totals = _exchange(_local_totals(results), streams)
    ph.mark("exchange")
    _assemble(out_prefix, k, streams, results, totals, countonly, mode)
    ph.mark("write")
    return {s: (totals[s][1], totals[s][2]) for s in streams}


def multi_files(paths, out_prefix="out", *, op="union", countonly=0, cutoff=1, rule=api.RULE_DEFAULT, count_override=1,
                stream=False, merge_fn=None, mode=0o644, timings=None):
    """Sharded `glistcompare L0 L1 L2 ... -u|-i` (N lists).  Returns {stream: (n_words, total_count)}."""
    _, rank, world = _dist()
    ph = _Phases(timings)
    paths = [Path(p) for p in paths]
    headers, bounds, _ = plan(paths, world, stream)
    ph.mark("plan")
    # header word length as in the reference: first non-empty list for the union (glistcompare.c:535), list 0 for the intersection (:639)
    k = headers[0].word_length
    if op == "union":
        k = next((h.word_length for h in headers if h.n_words), headers[-1].word_length)
    ranges = [(int(bounds[j, rank]), int(bounds[j, rank + 1])) for j in range(len(paths))]
    if op != "union" and any(h.n_words == 0 for h in headers):
        # an empty list anywhere empties the intersection (:631-636); a rank whose RANGE of some list is empty is not that
        ranges = [(0, 0)] * len(paths)
    kwargs = dict(op=op, cutoff=cutoff, rule=rule, count_override=count_override, countonly=countonly)
    results = _gpu_merge_multi(paths, ranges, k, kwargs, stream, ph) if merge_fn is None else merge_fn(paths, ranges, k, kwargs, stream)
    ph.mark("merge")
    streams = list(results)
    totals = _exchange(_local_totals(results), streams)
    ph.mark("exchange")
    _assemble(out_prefix, k, streams, results, totals, countonly, mode)
    ph.mark("write")
    return {s: (totals[s][1], totals[s][2]) for s in streams}
