#!/usr/bin/env python
"""bench.py -- headline benchmark of the set-operation path (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, one process per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU glistcompare

Workload (configs[1] of BASELINE.json): ``glistcompare -u`` (rule add, cut-off 1) of two sorted
25-mer lists of ~1e9 k-mers each, 50 % overlap, on every GPU (weak scaling: rank r owns the r-th
key range of a world-size-times-larger pair of lists, so the shards are exactly what the key-range
splitters of section 8(e) would hand out).  A step = one pass of the hot path over the rank's shard:
partition kernel + tile kernel (+ the allgather of per-rank output counts when N > 1).

One JSON line is printed by rank 0; see README/DESIGN.md section 6 for the fields.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "input k-mers/sec merged (glistcompare -u, rule add, 2 x 25-mer lists)"
UNIT = "k-mers/s"
FALLBACK_HBM_GBS = 6650.0     # /opt/skills/guides/B200_PROFILING.md, used only if MEASURED_PEAKS.json is absent


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["gt4gpu", "reference"], default="gt4gpu")
    ap.add_argument("--n-per-list", type=float, default=1e9, help="k-mers per list per GPU")
    ap.add_argument("--overlap", type=float, default=0.5, help="|A and B| / |A|")
    ap.add_argument("--k", type=int, default=25)
    ap.add_argument("--op", choices=["union", "intersect", "diff"], default="union")
    ap.add_argument("--cutoff", type=int, default=1)
    ap.add_argument("--count-only", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the end-to-end leg (0: --steps, at most 20)")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE.json configurations (3, 4, 5, sharded files)")
    ap.add_argument("--configs-scale", type=float, default=1.0, help="shrink the extra configurations (debugging)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=float, default=1e8, help="k-mers per list of the CPU baseline sample")
    ap.add_argument("--ref-sample", type=float, default=4e7, help="k-mers per list per step of --impl reference")
    ap.add_argument("--tile", type=str, default=None, help="multi-output kernel tile, e.g. 256x9")
    ap.add_argument("--stream-shape", type=str, default=None, help="single-output kernel: consumers x items, e.g. 512x11")
    ap.add_argument("--no-stream-kernel", action="store_true", help="route the merge through setop2_tile_kernel")
    return ap.parse_args()


def universe_for(n_per_list: float, overlap: float):
    """|A| = |B| = n with |A & B| = overlap * n  ->  universe M and the (a_only, b_only) shares."""
    n_both = overlap * n_per_list
    m = int(round(2 * n_per_list - n_both))
    p_only = (n_per_list - n_both) / m
    return m, p_only, p_only


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per launch of the tile kernel from the committed ncu summary of this workload, if any."""
    p = ROOT / "profiles" / "ncu_summary.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get("setop2_stream_kernel", {}).get("dram_bytes_per_launch")
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi sampled every 100 ms from before the warm-up; only samples whose timestamp falls inside
    the timed region count (the recipe's clocks line, /opt/skills/guides/B200_PROFILING.md)."""
    QUERY = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            self.path = tempfile.NamedTemporaryFile(prefix="clocks_", suffix=".csv", delete=False).name
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self, t0: float, t1: float):
        import datetime
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        inside, everything, smax, reasons, power = [], [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in Path(self.path).read_text().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                sm = float(f[1])
                smax = float(f[2])
            except ValueError:
                continue
            everything.append(sm)
            if t0 - 0.05 <= ts <= t1 + 0.05:
                inside.append(sm)
                try:
                    power.append(float(f[3]))
                except ValueError:
                    pass
                for nme, val in zip(names, f[4:8]):
                    if val.lower().startswith("active"):
                        reasons.add(nme)
        os.unlink(self.path)
        use = inside if inside else everything
        return {"sm_mhz": statistics.median(use) if use else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples_in_timed_region": len(inside), "samples": len(everything),
                "power_w_max": max(power) if power else None}


# ------------------------------------------------------------------------------------------ reference arm

def build_sample_lists(tmpdir: Path, n_per_list: float, overlap: float, k: int, seed: int = 42):
    """First `n_per_list` k-mers per list of the workload, generated on the host (numpy twin of the
    device generator) and written as .list files the way glistmaker would."""
    import numpy as np
    from genometester4_b200 import synth
    from genometester4_b200.api import RECORD
    m, pa, pb = universe_for(n_per_list, overlap)
    paths = []
    (wa, ca), (wb, cb) = synth.pair_numpy(seed, k, m, 0, m, pa, pb)
    for name, w, c in (("A", wa, ca), ("B", wb, cb)):
        rec = np.empty(w.size, dtype=RECORD)
        rec["word"], rec["count"] = w, c
        hdr = np.zeros(1, dtype=[("code", "<u4"), ("major", "<u4"), ("minor", "<u4"), ("k", "<u4"), ("n", "<u8"),
                                 ("total", "<u8"), ("start", "<u8"), ("wb", "<u4"), ("cb", "<u4")])
        hdr["code"], hdr["major"], hdr["minor"], hdr["k"] = 0x47543443, 4, 2, k
        hdr["n"], hdr["total"], hdr["start"], hdr["wb"], hdr["cb"] = w.size, int(c.astype(np.uint64).sum()), 48, 8, 4
        p = tmpdir / f"sample_{name}.list"
        with open(p, "wb") as f:
            f.write(hdr.tobytes())
            f.write(rec.tobytes())
        paths.append(p)
    return paths, wa.size + wb.size


def reference_binary():
    p = ROOT / "oracle" / "_ref" / "glistcompare"
    if not p.exists() and Path("/root/reference/src").exists():
        subprocess.run(["make", "-C", str(ROOT / "oracle"), "ref"], check=False, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return p if p.exists() else None


def time_reference(paths, op_flag: str, cutoff: int, count_only: bool, cwd: Path) -> float:
    exe = reference_binary()
    args = [str(exe), str(paths[0]), str(paths[1]), op_flag, "-c", str(cutoff), "-o", str(cwd / "ref")]
    if count_only:
        args.append("--count_only")
    t0 = time.perf_counter()
    subprocess.run(args, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return time.perf_counter() - t0


def time_oracle_port(paths, op: str, cutoff: int) -> float:
    from oracle import oracle as O
    a, b = O.read_list(paths[0]), O.read_list(paths[1])
    t0 = time.perf_counter()
    O.compare2(a, b, union=op == "union", intrsec=op == "intersect", diff=op == "diff", cutoff=cutoff)
    return time.perf_counter() - t0


OP_FLAG = {"union": "-u", "intersect": "-i", "diff": "-d"}


def cpu_arm(n_per_list: float, overlap: float, k: int, op: str, cutoff: int, count_only: bool, repeats: int, warmup: int, parity_with=None):
    """Times the reference's CPU implementation on a bounded sample.  Returns (input k-mers, all times, info dict).
    parity_with: callable (path_a, path_b) -> bytes of the list file the GPU path produces for the same inputs; when
    given (and the reference binary wrote a file) the two are compared byte for byte and the verdict is returned in info."""
    import hashlib
    shm = Path("/dev/shm") if Path("/dev/shm").is_dir() else None
    parity = None
    with tempfile.TemporaryDirectory(dir=shm) as td:
        td = Path(td)
        paths, n_in = build_sample_lists(td, n_per_list, overlap, k)
        kind = "reference" if reference_binary() is not None else "port"
        times = []
        for it in range(warmup + repeats):
            t = time_reference(paths, OP_FLAG[op], cutoff, count_only, td) if kind == "reference" else time_oracle_port(paths, op, cutoff)
            if it >= warmup:
                times.append(t)
        best = min(times)
        if parity_with is not None and kind == "reference" and not count_only:
            tag = {"union": "union", "intersect": "intrsec", "diff": "0_diff1"}[op]
            ref_file = td / f"ref_{k}_{tag}.list"
            ref_bytes = ref_file.read_bytes()
            ours = parity_with(paths[0], paths[1])
            parity = {"records": int(n_in), "output_records": (len(ref_bytes) - 48) // 12,
                      "reference_sha256": hashlib.sha256(ref_bytes).hexdigest(), "gt4gpu_sha256": hashlib.sha256(ours).hexdigest(),
                      "identical": ref_bytes == ours,
                      "what": f"unmodified glistcompare {OP_FLAG[op]} -c {cutoff} vs gt4gpu on the same two {n_in // 2}-k-mer list files, whole output file"}
            del ref_bytes, ours
    info = {"value": n_in / best, "unit": UNIT, "kind": kind,
            "cores": 3 if kind == "reference" else 1,
            "sample": (f"{OP_FLAG[op]}{' --count_only' if count_only else ''} of the first {n_in} k-mers of the workload "
                       f"({n_in // 2} per list, page-cache-hot .list files in /dev/shm, best of {len(times)}); "
                       + ("unmodified glistcompare from oracle/_ref: 1 merge thread + 2 mmap scout threads (the reference has no multi-threaded merge)"
                          if kind == "reference" else "oracle C port, 1 thread"))}
    if parity is not None:
        info["parity_check"] = parity
    return n_in, times, info


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ref_warmup = min(args.warmup, 1)          # one untimed run pages the sample files in; more would only repeat it
    n_in, times, info = cpu_arm(args.ref_sample, args.overlap, args.k, args.op, args.cutoff, args.count_only,
                                repeats=args.steps, warmup=ref_warmup)
    ms = 1000.0 * sum(times) / len(times)
    value = n_in / (ms / 1000.0)
    info["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": ref_warmup, "warmup_requested": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64 keys / u32 counts", "data": "synthetic",
            "config": workload_config(args, n_in // 2, n_in // 2, sample=True),
            "cpu_baseline": info,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(args, na, nb, sample=False):
    return {"workload": f"glistcompare {OP_FLAG[args.op]} (rule default, cutoff {args.cutoff}"
                        f"{', --count_only' if args.count_only else ''}) of two sorted {args.k}-mer lists, "
                        f"~{args.n_per_list:.0e} k-mers each per GPU, {int(args.overlap * 100)}% overlap (BASELINE.json configs[1])",
            "k": args.k, "n_a_per_gpu": int(na), "n_b_per_gpu": int(nb), "overlap": args.overlap,
            "sharding": "key-range, one shard per GPU, no payload exchange",
            "l2_policy": "inputs (>= 24 GB per step) exceed the 126 MB L2; no flush needed" if not sample else "cpu sample",
            "cpu_sample": bool(sample)}


# ------------------------------------------------------------------------------------------ our arm

def bind_to_gpu_numa_node(local_rank: int):
    """Best effort: run this rank (and allocate its pinned host buffers) on the NUMA node its GPU hangs off, so that the
    end-to-end path does not cross sockets.  Returns the node or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:          # 00000000:1b:00.0 -> 0000:1b:00.0
            bus = bus[4:]
        node = int(Path(f"/sys/bus/pci/devices/{bus}/numa_node").read_text())
        if node < 0:
            return None
        cpus = set()
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def run_gt4gpu_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import genometester4_b200 as g
    from genometester4_b200 import api, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; genometester4_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    numa_node = bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    g.init(local)
    if args.tile:
        nt, vt = args.tile.split("x")
        g.set_tile(int(nt), int(vt))
    if args.stream_shape:
        nc, vt = args.stream_shape.split("x")
        g.set_option("stream_shape", int(nc) * 100 + int(vt))
    if args.no_stream_kernel:
        g.set_option("use_stream_kernel", 0)
    g.set_stream(torch.cuda.current_stream().cuda_stream)

    # ---- this rank's shard of the synthetic lists, generated in HBM
    m_local, pa, pb = universe_for(args.n_per_list, args.overlap)
    universe = m_local * world
    (wa, ca), (wb, cb) = synth.pair_torch(42, args.k, universe, rank * m_local, (rank + 1) * m_local, pa, pb, device="cuda")
    torch.cuda.synchronize()
    na, nb = wa.numel(), wb.numel()
    la = g.WordList.from_device(wa.data_ptr(), ca.data_ptr(), na, args.k, keepalive=(wa, ca))
    lb = g.WordList.from_device(wb.data_ptr(), cb.data_ptr(), nb, args.k, keepalive=(wb, cb))
    stream_name = {"union": "union", "intersect": "intrsec", "diff": "diff1"}[args.op]
    kw = {"union": dict(find_union=1), "intersect": dict(find_intrsec=1), "diff": dict(find_diff=1)}[args.op]
    cap = {"union": na + nb, "intersect": min(na, nb), "diff": na}[args.op]
    out_buffers = None
    if not args.count_only:
        ow = torch.empty(cap, dtype=torch.int64, device="cuda")
        oc = torch.empty(cap, dtype=torch.int32, device="cuda")
        out_buffers = {stream_name: (ow.data_ptr(), oc.data_ptr(), cap)}
    torch.cuda.empty_cache()
    gather = torch.zeros(world, 2, dtype=torch.int64, device="cuda") if world > 1 else None

    def step():
        r = g.compare_wordmaps(la, lb, cutoff=args.cutoff, countonly=int(args.count_only), out_buffers=out_buffers, **kw)[stream_name]
        if world > 1:
            # the path's only collective: per-rank {n_out, sum} -> global record offsets + header totals
            mine = torch.tensor([r.n_words, r.total_count], dtype=torch.int64, device="cuda")
            dist.all_gather_into_tensor(gather.view(-1), mine)
        return r

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        res = step()
    n_out = res.n_words
    part_ms, merge_ms, launches = [], [], 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    wall0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        step()
        p_ms, m_ms, nl = g.last_timing()
        part_ms.append(p_ms)
        merge_ms.append(m_ms)
        launches += nl
    ev1.record()
    barrier()
    wall1 = time.time()
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    ms_step = ev0.elapsed_time(ev1) / args.steps
    stats = torch.tensor([ms_step, statistics.mean(merge_ms), statistics.mean(part_ms)], dtype=torch.float64, device="cuda")
    sizes = torch.tensor([na + nb, n_out], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(sizes, op=dist.ReduceOp.SUM)
    ms_step, merge_ms_max, part_ms_max = (float(x) for x in stats.tolist())
    total_in, total_out = (int(x) for x in sizes.tolist())
    value = total_in / (ms_step / 1000.0)

    # ---- roofline of the dominant kernel (setop2_tile_kernel), per launch on this rank
    peak, peak_src = peaks()
    algo_bytes = 12 * (na + nb) + (0 if args.count_only else 12 * n_out)
    achieved = algo_bytes / (statistics.mean(merge_ms) / 1000.0) / 1e9
    kernel_name = "setop2_tile_kernel" if args.no_stream_kernel else "setop2_stream_kernel"
    roofline = {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": ncu_traffic(), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": statistics.mean(merge_ms),
                "partition_kernel_ms": statistics.mean(part_ms),
                "step_frac_of_peak": (algo_bytes / (ms_step / 1000.0) / 1e9) / peak}

    # ---- end to end through the C ABI with HOST buffers (packed 12-byte records, pinned), copies timed
    e2e = None
    host_ceiling = None
    if not args.no_e2e:
        host_ceiling = measure_host_ceiling(torch, dist, world)
        e2e = run_e2e(args, g, api, torch, dist, world, rank, la, lb, wa, ca, wb, cb, na, nb, n_out, cap)
        if e2e and host_ceiling:
            # full duplex: a step cannot be shorter than its larger direction at the measured concurrent rate
            floor_s = max(e2e["h2d_bytes_per_step"] * world / (host_ceiling["h2d_gbs_concurrent"] * 1e9),
                          e2e["d2h_bytes_per_step"] * world / (host_ceiling["d2h_gbs_concurrent"] * 1e9))
            e2e["host_ceiling"] = host_ceiling
            e2e["frac_of_host_ceiling"] = floor_s / (e2e["ms_per_step"] / 1000.0)

    # the headline lists are no longer needed: give their HBM to the other configurations
    del la, lb, wa, ca, wb, cb, res
    if not args.count_only:
        del ow, oc
    out_buffers = None
    torch.cuda.empty_cache()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        def gpu_list_bytes(path_a, path_b):
            a, b = g.WordList.open(path_a), g.WordList.open(path_b)
            r = g.compare_wordmaps(a, b, cutoff=args.cutoff, **kw)[stream_name]
            return r.list_bytes()
        try:
            _, _, cpu = cpu_arm(args.cpu_sample, args.overlap, args.k, args.op, args.cutoff, args.count_only, repeats=1, warmup=0,
                                parity_with=gpu_list_bytes)
        except Exception as exc:  # pragma: no cover
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(exc)}
    parity = cpu.pop("parity_check", None) if isinstance(cpu, dict) else None

    configs = None
    if not args.no_configs:
        configs = run_other_configs(args, g, api, synth, torch, dist, world, rank, peaks()[0])

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u64 keys / u32 counts (integer compare, add mod 2^32)", "data": "synthetic",
                "config": workload_config(args, na, nb), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
                "gpu_launches": launches, "clocks": clocks, "parity_check": parity,
                "output_kmers": total_out, "input_kmers": total_in, "numa_node_rank0": numa_node, "kernel_config": {"kernel": kernel_name, "stream_shape": args.stream_shape or os.environ.get("GT4GPU_STREAM_SHAPE", "512x9"),
                                  "tile": args.tile or os.environ.get("GT4GPU_TILE", "256x9")},
                "configs": configs}
        print(json.dumps(line), flush=True)
        if parity is not None and not parity["identical"]:
            raise SystemExit("bench.py: the GPU output differs from the reference's on the parity sample")
    if world > 1:
        dist.destroy_process_group()


def measure_host_ceiling(torch, dist, world, nbytes=1 << 30, reps=3):
    """What the host can feed: pinned host <-> device copies of 1 GiB on every rank at once, each direction alone and
    both together (the e2e path overlaps them).  Returns aggregate GB/s over all ranks (slowest rank's time)."""
    host_in = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    host_out = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    dev_in = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dev_out = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def timed(do_in, do_out):
        best = None
        for _ in range(reps + 1):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if do_in:
                with torch.cuda.stream(s1):
                    dev_in.copy_(host_in, non_blocking=True)
            if do_out:
                with torch.cuda.stream(s2):
                    host_out.copy_(dev_out, non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
            best = dt if best is None else min(best, dt)
        return world * nbytes / best / 1e9

    out = {"h2d_gbs_alone": timed(True, False), "d2h_gbs_alone": timed(False, True)}
    both = timed(True, True)
    out["h2d_gbs_concurrent"] = both
    out["d2h_gbs_concurrent"] = both
    out["what"] = f"{world} rank(s) x 1 GiB pinned host <-> device copies at once, aggregate GB/s per direction (best of {reps})"
    del host_in, host_out, dev_in, dev_out
    return out


def run_other_configs(args, g, api, synth, torch, dist, world, rank, peak):
    """The other BASELINE.json configurations, device resident, STRONG scaling (the global lists are fixed; rank r
    generates and merges the r-th key range, which is what the splitters hand out), plus the real file -> file sharded path
    (splitter planning, range loads, merge, all-gather, parallel pwrite) on list files in /dev/shm."""
    scale = args.configs_scale
    out = {}
    gather = torch.zeros(world, 2, dtype=torch.int64, device="cuda") if world > 1 else None

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, reps=3, warm=1):
        """ms per call (CUDA events around `reps` calls, max over ranks), kernel ms of the last call, last result."""
        r = None
        for _ in range(warm):
            del r                     # (hand the previous result's buffers back to the pool before the next call needs them)
            r = fn()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        ev0.record()
        for _ in range(reps):
            del r
            r = fn()
            if world > 1:
                mine = torch.tensor([r.n_words, r.total_count], dtype=torch.int64, device="cuda")
                dist.all_gather_into_tensor(gather.view(-1), mine)
        ev1.record()
        sync_all()
        p_ms, m_ms, _ = g.last_timing()
        t = torch.tensor([ev0.elapsed_time(ev1) / reps, p_ms + m_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), r

    def total(x):
        t = torch.tensor([x], dtype=torch.int64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return int(t.item())

    def entry(what, n_in, n_out, ms, kernel_ms, countonly, **extra):
        b = 12 * n_in + (0 if countonly else 12 * n_out)
        e = {"what": what, "scaling": "strong", "input_kmers": n_in, "output_kmers": n_out, "ms_per_step": ms, "kernels_ms": kernel_ms,
             "kmers_per_s": n_in / (ms / 1e3), "algorithmic_gb": b / 1e9, "frac_of_aggregate_hbm_peak": b / (ms / 1e3) / 1e9 / (peak * world)}
        e.update(extra)
        return e

    def pair_shard(k, n_a, n_b, n_both, seed=42):
        m = int(n_a + n_b - n_both)
        u0, u1 = m * rank // world, m * (rank + 1) // world
        (wa, ca), (wb, cb) = synth.pair_torch(seed, k, m, u0, u1, (n_a - n_both) / m, (n_b - n_both) / m)
        la = g.WordList.from_device(wa.data_ptr(), ca.data_ptr(), wa.numel(), k, keepalive=(wa, ca))
        lb = g.WordList.from_device(wb.data_ptr(), cb.data_ptr(), wb.numel(), k, keepalive=(wb, cb))
        return la, lb

    # ---- configs[1], the other operations on the headline lists (per-GPU weak-scaling shards as in the headline)
    n1 = args.n_per_list * scale
    m_local, pa, pb = universe_for(n1, args.overlap)
    (wa, ca), (wb, cb) = synth.pair_torch(42, args.k, m_local * world, rank * m_local, (rank + 1) * m_local, pa, pb)
    la = g.WordList.from_device(wa.data_ptr(), ca.data_ptr(), wa.numel(), args.k, keepalive=(wa, ca))
    lb = g.WordList.from_device(wb.data_ptr(), cb.data_ptr(), wb.numel(), args.k, keepalive=(wb, cb))
    for name, kw, stream in (("2_intersect", dict(find_intrsec=1), "intrsec"), ("2_diff", dict(find_diff=1), "diff1"),
                             ("2_union_intersect_diff", dict(find_union=1, find_intrsec=1, find_diff=1), None)):
        def run():
            r = g.compare_wordmaps(la, lb, cutoff=args.cutoff, **kw)
            return r[stream] if stream else r["union"]
        ms, kms, r = timed(run)
        n_in = total(len(la) + len(lb))
        if stream:
            n_out = total(r.n_words)
        else:
            rr = g.compare_wordmaps(la, lb, cutoff=args.cutoff, **kw)
            n_out = total(sum(x.n_words for x in rr.values()))
            del rr
        out[name] = entry(f"glistcompare {' '.join('-' + c for c in ('u' if 'find_union' in kw else '') + ('i' if 'find_intrsec' in kw else '') + ('d' if 'find_diff' in kw else ''))} "
                          f"on the headline lists ({n1:.0e} k-mers per list per GPU, weak scaling)", n_in, n_out, ms, kms, False, scaling="weak")
        del r
    del la, lb, wa, ca, wb, cb
    torch.cuda.empty_cache()

    # ---- configs[2]: glistcompare -d -c 5, 32-mers, 3e9 / 1e9 k-mers, key-range sharded over the GPUs
    la, lb = pair_shard(32, 3e9 * scale, 1e9 * scale, 0.8e9 * scale)
    for co in (0, 1):
        ms, kms, r = timed(lambda: g.compare_wordmaps(la, lb, find_diff=1, cutoff=5, countonly=co)["diff1"])
        out["3" + ("_count_only" if co else "")] = entry(
            f"glistcompare -d -c 5{' --count_only' if co else ''}, 32-mers, {3e9 * scale:.0e} / {1e9 * scale:.0e} k-mers, key-range sharded over {world} GPU(s)",
            total(len(la) + len(lb)), total(r.n_words), ms, kms, bool(co))
        del r
    del la, lb
    torch.cuda.empty_cache()

    # ---- configs[3]: counts-only intersect / union sweep
    sweep = []
    for n, ovs in ((1e6, (0.01, 0.5, 0.99)), (1e7, (0.5,)), (1e8, (0.01, 0.5, 0.99)), (1e9, (0.5,)), (3e9, (0.01, 0.5, 0.99))):
        n = n * scale
        for ov in ovs:
            la, lb = pair_shard(25, n, n, ov * n)
            for op, kw, stream in (("intersect", dict(find_intrsec=1), "intrsec"), ("union", dict(find_union=1), "union")):
                ms, kms, r = timed(lambda: g.compare_wordmaps(la, lb, countonly=1, **kw)[stream], reps=3 if n >= 1e8 else 10)
                e = entry(f"--count_only {op}", total(len(la) + len(lb)), total(r.n_words), ms, kms, True, n_per_list=int(n), overlap=ov, op=op)
                del e["what"]
                sweep.append(e)
                del r
            del la, lb
            torch.cuda.empty_cache()
    out["4"] = {"what": f"counts-only intersect / union sweep, 25-mers, lists key-range sharded over {world} GPU(s)", "points": sweep}

    # ---- configs[4]: union of 8 lists of 5e8 32-mers drawn from one universe (MakeUnion.pl), single-pass k-way kernel
    n_each = int(5e8 * scale)
    m = n_each * 3
    u0, u1 = m * rank // world, m * (rank + 1) // world
    lists, keep = [], []
    for j in range(8):
        w, c = synth.list_torch(5, 32, m, u0, u1, j, 1 / 3)
        keep.append((w, c))
        lists.append(g.WordList.from_device(w.data_ptr(), c.data_ptr(), w.numel(), 32))
    n_in = total(sum(len(x) for x in lists))
    for name, fn in (("5", g.union_multi), ("5_intersect", g.intersect_multi)):
        for co in (0, 1):
            ms, kms, r = timed(lambda: fn(lists, cutoff=1, countonly=co), reps=2)
            out[name + ("_count_only" if co else "")] = entry(
                f"{'union' if fn is g.union_multi else 'intersection'} of 8 lists of {n_each:.0e} 32-mers{' --count_only' if co else ''} (single-pass k-way kernel), key-range sharded over {world} GPU(s)",
                n_in, total(r.n_words), ms, kms, bool(co))
            del r
    del lists, keep
    torch.cuda.empty_cache()

    # ---- the real sharded file path (genometester4_b200/sharded.py) on config-3 shaped list files in /dev/shm
    try:
        out["sharded_files"] = run_sharded_files(args, g, api, synth, torch, dist, world, rank, scale)
    except Exception as exc:  # pragma: no cover
        out["sharded_files"] = {"error": repr(exc)}
    return out


def run_sharded_files(args, g, api, synth, torch, dist, world, rank, scale):
    """glistcompare -d -c 5 of two 32-mer list FILES (3e8 / 1e8 k-mers, a tenth of configs[2]) through
    sharded.compare_files: host-side splitter planning on the mmaps, every rank loads its record range, merges, the
    ranks all-gather {n_out, sum}, every rank pwrites its slice of the output file.  Whole call timed per phase."""
    import hashlib
    import shutil

    import numpy as np

    from genometester4_b200 import sharded
    n_a, n_b, n_both, k = int(3e8 * scale), int(1e8 * scale), int(0.8e8 * scale), 32
    shm = Path("/dev/shm") if Path("/dev/shm").is_dir() else Path(tempfile.gettempdir())
    td = shm / "gt4gpu_bench_sharded"
    if rank == 0:
        shutil.rmtree(td, ignore_errors=True)
        td.mkdir(parents=True)
        m = n_a + n_b - n_both
        (wa, ca), (wb, cb) = synth.pair_torch(7, k, m, 0, m, (n_a - n_both) / m, (n_b - n_both) / m)
        for name, w, c in (("A", wa, ca), ("B", wb, cb)):
            n = w.numel()
            dev = torch.empty(n * 12, dtype=torch.uint8, device="cuda")
            assert api._lib.load().gt4gpu_interleave(w.data_ptr(), c.data_ptr(), n, dev.data_ptr()) == 0
            rec = dev.cpu().numpy()
            hdr = np.zeros(1, dtype=[("code", "<u4"), ("major", "<u4"), ("minor", "<u4"), ("k", "<u4"), ("n", "<u8"),
                                     ("total", "<u8"), ("start", "<u8"), ("wb", "<u4"), ("cb", "<u4")])
            hdr["code"], hdr["major"], hdr["minor"], hdr["k"] = 0x47543443, 4, 2, k
            hdr["n"], hdr["total"], hdr["start"], hdr["wb"], hdr["cb"] = n, int(c.to(torch.int64).sum().item()), 48, 8, 4
            with open(td / f"{name}.list", "wb") as f:
                f.write(hdr.tobytes())
                rec.tofile(f)
            del dev, rec
        del wa, ca, wb, cb
        torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
    best, phases_best, tot = None, None, None
    for it in range(3):
        timings = {}
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tot = sharded.compare_files(td / "A.list", td / "B.list", str(td / "out"), find_diff=1, cutoff=5, timings=timings)
        dt = time.perf_counter() - t0
        names = ["plan", "load", "merge", "exchange", "write"]
        t = torch.tensor([dt] + [timings.get(x, 0.0) for x in names], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if best is None or float(t[0]) < best:
            best = float(t[0])
            phases_best = {x: float(t[i + 1]) for i, x in enumerate(names)}
    res = None
    if rank == 0:
        out_file = td / f"out_{k}_0_diff1.list"
        sha = hashlib.sha256()
        with open(out_file, "rb") as f:
            while True:
                blk = f.read(1 << 26)
                if not blk:
                    break
                sha.update(blk)
        res = {"what": f"sharded.compare_files: glistcompare -d -c 5 of two 32-mer list files ({n_a} / {n_b} k-mers) in /dev/shm, {world} rank(s): "
                       "splitter planning, range loads (H2D), merge, all-gather of output counts, parallel pwrite; best of 3, phases = slowest rank",
               "input_kmers": n_a + n_b, "output_kmers": tot["diff1"][0], "seconds": best, "kmers_per_s": (n_a + n_b) / best,
               "phase_seconds": phases_best, "output_sha256": sha.hexdigest(),
               "bytes_in_plus_out": 12 * (n_a + n_b + tot["diff1"][0])}
        exe = reference_binary()
        if world == 1 and exe is not None and scale == 1.0:
            # the unmodified reference on the very same files: byte-for-byte parity at 4e8 input k-mers
            t0 = time.perf_counter()
            subprocess.run([str(exe), str(td / "A.list"), str(td / "B.list"), "-d", "-c", "5", "-o", str(td / "ref")], check=True,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            res["reference_seconds"] = time.perf_counter() - t0
            ref_sha = hashlib.sha256()
            with open(td / f"ref_{k}_0_diff1.list", "rb") as f:
                while True:
                    blk = f.read(1 << 26)
                    if not blk:
                        break
                    ref_sha.update(blk)
            res["reference_sha256"] = ref_sha.hexdigest()
            res["identical_to_reference"] = ref_sha.hexdigest() == sha.hexdigest()
        shutil.rmtree(td, ignore_errors=True)
    if world > 1:
        dist.barrier()
    return res


def run_e2e(args, g, api, torch, dist, world, rank, la, lb, wa, ca, wb, cb, na, nb, n_out, cap):
    """Same step through gt4gpu_compare2_host_aos: inputs are packed .list records in pinned host memory
    (what an mmap of the files holds), outputs land in pinned host memory; H2D, de-interleave, merge,
    interleave and D2H all inside the timed region.  Shrinks the sample if host memory is short."""
    import numpy as np
    avail = 0
    for line in Path("/proc/meminfo").read_text().splitlines():
        if line.startswith("MemAvailable:"):
            avail = int(line.split()[1]) * 1024
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    need = 12 * (na + nb + cap)
    frac = 1.0
    budget = 0.5 * avail / max(local_world, 1)
    if need > budget:
        frac = budget / need
    ea, eb = int(na * frac), int(nb * frac)
    ecap = ea + eb if args.op == "union" else (min(ea, eb) if args.op == "intersect" else ea)

    def pinned_records(words_t, counts_t, n):
        host = torch.empty(n * 12, dtype=torch.uint8, pin_memory=True)
        dev = torch.empty(n * 12, dtype=torch.uint8, device="cuda")
        rc = api._lib.load().gt4gpu_interleave(words_t.data_ptr(), counts_t.data_ptr(), n, dev.data_ptr())
        assert rc == 0
        host.copy_(dev)
        torch.cuda.synchronize()
        del dev
        return host

    ha = pinned_records(wa, ca, ea)
    hb = pinned_records(wb, cb, eb)
    hout = None if args.count_only else torch.empty(max(ecap, 1) * 12, dtype=torch.uint8, pin_memory=True)
    torch.cuda.empty_cache()
    ops = {"union": api.OP_UNION, "intersect": api.OP_INTRSEC, "diff": api.OP_DIFF}[args.op]
    sidx = {"union": 0, "intersect": 1, "diff": 2}[args.op]
    outs = [None] * 4
    if hout is not None:
        outs[sidx] = (hout.data_ptr(), ecap)

    def one():
        n_o, t_o = api.compare2_host_records((ha.data_ptr(), ea), (hb.data_ptr(), eb), args.k, ops, cutoff=args.cutoff,
                                             countonly=int(args.count_only), out_records=outs)
        return n_o[sidx]

    e2e_steps = args.e2e_steps if args.e2e_steps > 0 else min(args.steps, 20)
    one()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e_out = one()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / e2e_steps
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    tot = torch.tensor([ea + eb], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dt = float(t.item())
    return {"value": int(tot.item()) / dt, "unit": UNIT, "h2d_bytes_per_step": 12 * (ea + eb),
            "d2h_bytes_per_step": 0 if args.count_only else 12 * int(e_out), "ms_per_step": dt * 1000.0, "steps": e2e_steps,
            "api": "gt4gpu_compare2_host_aos (packed 12-byte records in pinned host memory in and out)",
            "sample_fraction_of_workload": frac}


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gt4gpu_arm(args)


if __name__ == "__main__":
    main()
