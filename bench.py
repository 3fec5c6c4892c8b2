#!/usr/bin/env python
"""bench.py -- string pairs/sec of the row-wise similarity hot path on B200 (BASELINE.json metric).

A step = one pass of every measure of the workload (default: BASELINE config C2, 10M ASCII name pairs
of length 4..24 per GPU, all five measures; seeds in SURVEY.md 8(d)) over one batch of synthetic pairs.

  value     pair evaluations (rows x measures) per second with the two columns already resident in
            HBM: ONE fused kernel launch per step (`strsim_b200_compute_device_multi`), CUDA events on
            the launching stream.
  e2e       the same step the way a Polars query runs it (/root/reference/polars_strsim/__init__.py:
            8-60: one expression = one plugin call): one `_polars_plugin_<measure>` call per measure
            over PAGEABLE host Arrow buffers, results landing in the Arrow buffers the plugin allocates;
            the plugin's column cache is emptied before every step, so every step uploads both columns
            (host->device) and downloads every measure's results (device->host) inside the timed region.
  e2e_pinned  the same step through ONE `strsim_b200_compute_host_multi` call with pinned buffers (what
            a caller that controls its memory gets); secondary.
  roofline  the fused kernel against measured HBM bandwidth (+ the column-statistics pre-pass that an
            upload runs once, timed next to it), and against the integer issue rates measured in the run.
  cpu_baseline / --impl reference
            the reference's CPU algorithm (the oracle port: no cargo in this image)
            with the reference's static row-range threading on all host cores, same rows, same metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C2|C3|C4|C5|L1|M1|N1|T1] [--rows R]
    python bench.py --impl reference ...
    python bench.py --strong --gpus N      # ONE plugin call sharded over N GPUs inside the library
    torchrun --nproc-per-node N bench.py --gpus N ...   # one process per GPU, rows sharded by range

Multi-GPU is row-range sharding with no data-path collective: under torchrun every rank
owns `rows` rows (weak scaling); NCCL is used only for the barrier and the max-over-ranks of the time.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
for p in (ROOT, ROOT / "polars-strsim_b200"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

MEASURES = ("levenshtein", "jaro", "jaro_winkler", "jaccard", "sorensen_dice")
WORKLOADS = {
    "C2": dict(config=2, rows=10_000_000, measures=MEASURES,
               desc="C2: 10M synthetic ASCII name pairs, lengths 4-24, all five measures"),
    "C3": dict(config=3, rows=100_000_000, measures=MEASURES,
               desc="C3: mixed-Unicode name pairs (Latin-1 diacritics + CJK), 5% nulls, uneven chunks"),
    "C4": dict(config=4, rows=1_000_000, measures=("levenshtein",),
               desc="C4: long-text pairs, 200-4000 codepoints, Levenshtein"),
    "L1": dict(config=6, rows=10_000_000, measures=MEASURES,
               desc="L1 (not a BASELINE config): the Latin rows of C3 alone -- names of 4-24 codepoints, 15 % of the "
                    "characters from U+00C0-U+00FF, no nulls"),
    "M1": dict(config=7, rows=10_000_000, measures=MEASURES,
               desc="M1 (not a BASELINE config): medium ASCII strings of 20-60 characters (addresses) -- a third of the "
                    "rows fit the 32-byte kernels, the rest run the 64-bit instantiation"),
    "N1": dict(config=8, rows=10_000_000, measures=MEASURES,
               desc="N1 (not a BASELINE config): C2 with real-name spelling -- capitalised words, spaces, hyphens and "
                    "apostrophes (any ASCII: the 7-plane instantiation of the short kernel)"),
    "T1": dict(config=9, rows=10_000_000, measures=MEASURES,
               desc="T1 (not a BASELINE config): M1 with a tail -- one row in ten has 100-300 characters and takes the "
                    "long-row kernels (warp-cooperative Jaro / multiset, multi-word Myers)"),
    "C5": dict(config=5, rows=125_000_000, measures=("jaro_winkler", "sorensen_dice"),
               desc="C5: record-linkage pairs (C2 generator, seed 0xC5), Jaro-Winkler + Sorensen-Dice"),
}
UNIT = "pairs/s"
# ONE metric string for both arms (the driver forms the ratio of the two lines only when they agree)
METRIC = "string pairs/sec (pair evaluations = rows x measures of the workload, whole job)"
DTYPE = "u8 bytes / u32 codepoints, u32/u64 bit-vectors, f64 results"


def run_config(args, wl, world):
    """`config` of the JSON line: identical in both arms (same workload, rows, measures, sharding)."""
    return {"workload": wl["desc"], "rows_per_gpu": args.rows, "measures": list(wl["measures"]),
            "parallelism": (f"one call sharded over {args.gpus} GPUs by row range, no collective" if args.strong
                            else f"row-range x{world}, no collective")}


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def kernel_source_hash() -> str:
    """sha256 over the device code (csrc/*.cuh: every kernel lives there; host.cu and arrow_plugin.cpp are the
    host driver): profiles/traffic.json carries the hash of the kernels its ncu capture was taken on, and a
    capture of other kernels is refused instead of silently quoted."""
    h = hashlib.sha256()
    src = ROOT / "polars-strsim_b200" / "csrc"
    for f in sorted(src.glob("*.cuh")):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    return h.hexdigest()[:16]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for name, val in zip(names, r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return world, rank, local


def flat(col):
    import pyarrow as pa

    return col.combine_chunks() if isinstance(col, pa.ChunkedArray) else col


def oracle_rate(A, B, measures, threads, probe=50_000):
    """pair evaluations/s of the oracle port on a small prefix (sizes the bounded samples)"""
    from oracle import oracle

    probe = min(len(A), probe)
    t0 = time.perf_counter()
    for m in measures:
        oracle.batch_views(m, A.slice(0, probe), B.slice(0, probe), n_threads=threads)
    return probe * len(measures) / max(time.perf_counter() - t0, 1e-6)


def cpu_baseline(A, B, measures, budget_s: float = 15.0):
    """The oracle (a port of the reference's algorithms with its static row-range threading,
    strsim.rs:21-39,72-100) on this box's host cores, on a bounded sample of the same workload."""
    from oracle import oracle

    threads = oracle.n_host_threads()
    n = len(A)
    A, B = flat(A), flat(B)
    rate = oracle_rate(A, B, measures, threads)
    sample = int(min(n, max(50_000, rate * budget_s / len(measures))))
    per = {}
    t_all = 0.0
    for m in measures:
        t0 = time.perf_counter()
        oracle.batch_views(m, A.slice(0, sample), B.slice(0, sample), n_threads=threads)
        dt = time.perf_counter() - t0
        per[m] = sample / dt
        t_all += dt
    return {"value": sample * len(measures) / t_all, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"first {sample} rows of the workload x {len(measures)} measures, "
                      f"C restatement of polars-strsim's rayon path, {threads} threads",
            "per_measure": per}


def run_reference(args, wl, world, rank):
    """--impl reference: the reference's own CPU implementation of the path.  The reference is Rust and
    cannot be built in this image (no cargo), so this times the oracle port (kind "port") with all
    host threads -- the reference's static row-range partition, one thread per range (strsim.rs:21-39,
    72-100).  A step covers the workload's rows whenever the whole run then ends within a few minutes
    (C2: 10M rows x 5 measures is about a second per step on 16 cores); otherwise a bounded prefix of the
    same rows, stated in `cpu_baseline.sample`."""
    if rank != 0:
        return
    from bench_support import workloads
    from oracle import oracle

    measures = wl["measures"]
    threads = oracle.n_host_threads()
    probe_rows = min(args.rows, 50_000 if wl["config"] != 4 else 200)
    Ap, Bp = workloads.make_pairs(wl["config"], probe_rows)
    rate = oracle_rate(flat(Ap), flat(Bp), measures, threads, probe=probe_rows)
    budget_s = 150.0  # whole run: warm-up + timed steps
    per_step = rate * budget_s / max(1, args.steps + args.warmup) / len(measures)
    sample = int(min(args.rows, max(probe_rows, per_step)))
    if sample > 0.8 * args.rows:
        sample = args.rows
    A, B = workloads.make_pairs(wl["config"], sample)
    A, B = flat(A), flat(B)
    for _ in range(args.warmup):
        for m in measures:
            oracle.batch_views(m, A, B, n_threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for m in measures:
            oracle.batch_views(m, A, B, n_threads=threads)
    dt = time.perf_counter() - t0
    value = sample * len(measures) * args.steps / dt
    whole = sample == args.rows
    line = {
        "impl": "reference", "metric": METRIC,
        "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong" if args.strong else "weak",
        "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": run_config(args, wl, world),
        "rows_per_step": sample,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": (f"all {sample} rows of the workload" if whole else
                                    f"first {sample} of the workload's {args.rows} rows") +
                                   f" x {len(measures)} measures per step, C restatement of polars-strsim's "
                                   f"rayon path, {threads} threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def input_bytes(A, B) -> int:
    """bytes one upload of the two columns moves host -> device (chunks made by slicing share buffers:
    the library uploads each distinct buffer once)"""
    seen, total = set(), 0
    for col in (A, B):
        for ch in (col.chunks if hasattr(col, "chunks") else [col]):
            bufs = ch.buffers()
            total += 16 * len(ch) + ((len(ch) + 7) // 8 if bufs[0] is not None else 0)
            for b in bufs[2:]:
                if b is not None and b.address not in seen:
                    seen.add(b.address)
                    total += b.size
    return total


def run_strong(args, wl):
    """--strong: ONE process drives all --gpus devices.  `value`: the workload's rows cut into one contiguous
    range per GPU (the library's own rule: split_offsets on 64-row units), every range resident on its GPU,
    one fused launch per GPU and step from one host thread per GPU, time = max over the GPUs (CUDA events).
    `e2e`: every `_polars_plugin_<measure>` call is sharded over the GPUs INSIDE the library
    (STRSIM_B200_DEVICES): each device uploads its rows' views and its stretch of the data buffers, computes,
    and downloads into its range of the one result buffer."""
    import numpy as np
    import torch

    from bench_support import plugin_driver, workloads
    from polars_strsim import _native

    G = args.gpus
    if torch.cuda.device_count() < G:
        raise SystemExit(f"--strong --gpus {G}: only {torch.cuda.device_count()} devices visible")
    measures = wl["measures"]
    n = args.rows
    A, B = workloads.make_pairs(wl["config"], n, pinned=False, uneven_b=(wl["config"] == 3))
    alg_bytes = workloads.algorithmic_bytes(A, B)
    units, per = (n + 63) // 64, ((n + 63) // 64) // G
    cuts = [g * per * 64 for g in range(G)] + [n]
    barrier = threading.Barrier(G)
    ms, sums, errors = [0.0] * G, [None] * G, []

    def worker(g):
        try:
            torch.cuda.set_device(g)
            _native.set_device(g)
            lo, hi = cuts[g], cuts[g + 1]
            colA, colB = _native.DeviceColumn(A.slice(lo, hi - lo)), _native.DeviceColumn(B.slice(lo, hi - lo))
            outs = [torch.empty(hi - lo, dtype=torch.float64, device=f"cuda:{g}") for _ in measures]
            ptrs = [o.data_ptr() for o in outs]
            nulls = A.null_count + B.null_count > 0
            val = torch.zeros((hi - lo + 31) // 32, dtype=torch.int32, device=f"cuda:{g}") if nulls else None
            stream = torch.cuda.Stream(device=g)  # not the legacy default stream (see main())
            torch.cuda.set_stream(stream)

            def step():
                if len(measures) > 1:
                    _native.compute_device_multi(measures, colA, colB, ptrs, val.data_ptr() if nulls else 0, None,
                                                 stream.cuda_stream)
                else:
                    _native.compute_device(measures[0], colA, colB, ptrs[0], val.data_ptr() if nulls else 0, 0,
                                           stream.cuda_stream)

            for _ in range(args.warmup):
                step()
            torch.cuda.synchronize(g)
            barrier.wait()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(args.steps):
                step()
            e1.record(stream)
            torch.cuda.synchronize(g)
            ms[g] = e0.elapsed_time(e1)
            sums[g] = [float(o.sum().item()) for o in outs]
        except Exception as exc:  # noqa: BLE001
            errors.append((g, exc))
            try:
                barrier.abort()
            except Exception:  # noqa: BLE001
                pass

    sampler = ClockSampler(0)
    sampler.start()
    launches0 = _native.kernel_launches()
    threads = [threading.Thread(target=worker, args=(g,)) for g in range(G)]
    t_wall = time.perf_counter()
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    t_wall = time.perf_counter() - t_wall
    if errors:
        raise SystemExit(f"--strong: device worker failed: {errors}")
    launches = _native.kernel_launches() - launches0
    elapsed_ms = max(ms)
    checksums = {m: float(sum(sums[g][i] for g in range(G))) for i, m in enumerate(measures)}

    # ---- end to end: the sharded plugin calls
    def plugin_step(keep=False):
        plugin_driver.cache_clear()
        results = []
        for m in measures:
            r = plugin_driver.call(m, A, B)
            if keep:
                results.append(r)
            else:
                r.release()
        return results

    for k in range(4):  # see main(): the plugin learns the measures, the pinned result pool grows
        plugin_step()
        if k < 2:
            time.sleep(0.6)
    e2e_steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        plugin_step()
    dt = time.perf_counter() - t0
    res = plugin_step(keep=True)
    e2e_sums = [float(sum(float(v.sum()) for v in r.values())) for r in res]
    for r in res:
        r.release()
    plugin_driver.cache_clear()
    clocks = sampler.stop()
    has_nulls = A.null_count + B.null_count > 0
    peak, peak_src = peaks()
    dom_bytes = alg_bytes + 8 * n * (len(measures) - 1)
    dom_s = elapsed_ms / args.steps * 1e-3
    line = {
        "metric": METRIC, "value": n * len(measures) * args.steps / (elapsed_ms * 1e-3), "unit": UNIT, "n_gpus": G,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": run_config(args, wl, 1),
        "notes": {"value_is": "device-resident: one row range per GPU, one fused launch per GPU and step, max over GPUs",
                  "ms_per_gpu": [m / args.steps for m in ms], "rows_per_gpu": [cuts[g + 1] - cuts[g] for g in range(G)]},
        "roofline": {"bound": "hbm", "kernel": "short_kernel<fused>", "achieved": dom_bytes / dom_s / 1e9,
                     "peak": peak * G, "unit": "GB/s", "frac": dom_bytes / dom_s / 1e9 / (peak * G), "traffic": None,
                     "peak_source": peak_src + f" x {G} GPUs", "algorithmic_bytes_per_launch": dom_bytes / G,
                     "launch_ms": dom_s * 1e3},
        "clocks": clocks, "gpu_launches": launches, "checksums": checksums,
        "e2e": {"value": n * len(measures) * e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": input_bytes(A, B),
                "d2h_bytes_per_step": (8 * n + ((n + 7) // 8 if has_nulls else 0)) * len(measures),
                "ms_per_step": dt / e2e_steps * 1e3, "steps": e2e_steps, "devices": os.environ["STRSIM_B200_DEVICES"],
                "api": "one _polars_plugin_<measure> call per measure, each sharded by the library over the devices "
                       "(row ranges, no collective); pageable Arrow buffers in, one result buffer out; cache cleared "
                       "before every step",
                "checksum_matches_device": bool(all(
                    abs(s - checksums[m]) <= 1e-9 * max(1.0, abs(checksums[m])) for s, m in zip(e2e_sums, measures)))},
    }
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(A, B, measures)
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--rows", type=int, default=None, help="rows per GPU (default: the workload's size)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-fuse", action="store_true", help="step = one single-measure launch per measure")
    ap.add_argument("--strong", action="store_true",
                    help="ONE process: every plugin call of the e2e leg is sharded by the library over --gpus devices "
                         "(STRSIM_B200_DEVICES); strong scaling of a single call")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.rows is None:
        args.rows = wl["rows"]
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    world, rank, local = dist_setup()

    if args.impl == "reference":
        run_reference(args, wl, world, rank)
        return

    if args.strong:
        if world > 1:
            raise SystemExit("--strong is one process that drives all GPUs: run it without torchrun")
        os.environ["STRSIM_B200_DEVICES"] = ",".join(str(d) for d in range(args.gpus))
        run_strong(args, wl)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    from bench_support import plugin_driver, workloads
    from polars_strsim import _native

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: polars-strsim_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    _native.set_device(local)
    # several ranks on one host: keep this rank's generator threads and pinned buffers on the NUMA node of
    # its GPU, so that its copies do not cross the socket interconnect
    numa_node = _native.bind_thread_near_device(local) if world > 1 else -1
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    measures = wl["measures"]
    n = args.rows
    # ---- data: this rank's row range [rank*n, (rank+1)*n), generated into PAGEABLE host memory (numpy),
    # which is what a Polars column lives in
    A, B = workloads.make_pairs(wl["config"], n, row_base=rank * n, pinned=False, uneven_b=(wl["config"] == 3))
    alg_bytes = workloads.algorithmic_bytes(A, B)  # per launch of one measure
    colA, colB = _native.DeviceColumn(A), _native.DeviceColumn(B)
    has_nulls = A.null_count + B.null_count > 0
    fused = len(measures) > 1 and not args.no_fuse
    if args.no_fuse:
        os.environ["STRSIM_B200_NO_FUSE"] = "1"  # read once by the library: the e2e call follows suit
    outs = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in measures]
    out = outs[-1]
    out_ptrs = [o.data_ptr() for o in outs]
    val = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda") if has_nulls else None
    vptr = val.data_ptr() if val is not None else 0
    # a stream of our own: torch's current stream is the legacy default stream (handle 0), which the library
    # would read as "use your per-thread stream" -- and events recorded on the default stream do not order
    # with the library's non-blocking streams, so kernels launched after a call's last host synchronisation
    # (validity bitmap, long-row kernels) would fall outside the bracket
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sptr = stream.cuda_stream
    assert sptr != 0

    def step():
        if fused:
            _native.compute_device_multi(measures, colA, colB, out_ptrs, vptr, None, sptr)
        else:
            for i, m in enumerate(measures):
                _native.compute_device(m, colA, colB, out_ptrs[i], vptr, 0, sptr)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _native.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    launches = _native.kernel_launches() - launches0
    overflow = _native.last_overflow()
    elapsed_ms = max_over_ranks(e0.elapsed_time(e1))
    checksum = float(out.sum().item())
    checksums = {m: float(o.sum().item()) for m, o in zip(measures, outs)}
    # every rank owns other rows: the sum over the ranks of each measure's checksum stands for the whole job
    checksums_all_ranks = None
    if world > 1:
        t = torch.tensor([checksums[m] for m in measures], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        checksums_all_ranks = {m: float(v) for m, v in zip(measures, t.tolist())}

    # ---- the column-statistics pre-pass (an upload runs it once per column; it picks the kernel
    # instantiation): the same kernels again over the resident columns, CUDA events on the same stream
    colA.restat(sptr)
    colB.restat(sptr)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stat_reps = max(3, min(args.steps, 10))
    s0.record(stream)
    for _ in range(stat_reps):
        colA.restat(sptr)
        colB.restat(sptr)
    s1.record(stream)
    barrier()
    stats_ms = s0.elapsed_time(s1) / stat_reps

    # ---- each measure's own kernel (single-measure launches), same rules: CUDA events on the launching
    # stream, averaged over the steps; the clock sampler keeps running
    per_events = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                   for _ in measures] for _ in range(args.steps)]
    scratch = torch.empty(n, dtype=torch.float64, device="cuda")
    for i, m in enumerate(measures):
        _native.compute_device(m, colA, colB, scratch.data_ptr(), vptr, 0, sptr)
    for k in range(args.steps):
        for i, m in enumerate(measures):
            per_events[k][i][0].record(stream)
            _native.compute_device(m, colA, colB, scratch.data_ptr(), vptr, 0, sptr)
            per_events[k][i][1].record(stream)
    barrier()
    fused_matches_single = bool(abs(float(scratch.sum().item()) - checksum) == 0.0)
    del scratch
    clocks = sampler.stop() if rank == 0 else None
    cells = None
    if wl["config"] == 4:
        # long-string Levenshtein work unit: DP cells = sum la*lb over pairs with a != b,
        # codepoint lengths taken from the kernel's own debug record
        dbg = torch.zeros((n, 6), dtype=torch.int32, device="cuda")
        _native.compute_device("levenshtein", colA, colB, out.data_ptr(), 0, dbg.data_ptr(), sptr)
        torch.cuda.synchronize()
        d64 = dbg.to(torch.int64)
        cells = int((d64[:, 1] * d64[:, 2] * (d64[:, 0] == 0)).sum().item())
        del dbg, d64

    per_measure_ms = {m: float(np.mean([per_events[k][i][0].elapsed_time(per_events[k][i][1])
                                        for k in range(args.steps)])) for i, m in enumerate(measures)}

    # ---- end to end --------------------------------------------------------------------------------------
    e2e = e2e_pinned = None
    d2h_bytes = (8 * n + ((n + 7) // 8 if has_nulls else 0)) * len(measures)
    if not args.no_e2e:
        # (1) the product surface: one `_polars_plugin_<measure>` call per measure, pageable Arrow buffers in,
        # the plugin's own Arrow buffers out.  The cache of uploaded columns is emptied before every step:
        # every step pays its own host->device upload (the first call of the step) and every call its own
        # device->host download; the other calls of the step find the columns in HBM, as the five
        # expressions of one Polars query do (README.md:47-51).
        def plugin_step(keep=False):
            plugin_driver.cache_clear()
            results = []
            for m in measures:
                r = plugin_driver.call(m, A, B)
                if keep:
                    results.append(r)
                else:
                    r.release()
            return results

        import ctypes

        L = _native.lib()
        L.strsim_b200_speculation_stats.argtypes = [ctypes.POINTER(ctypes.c_int64)]
        L.strsim_b200_speculation_stats.restype = None

        def served():
            o = (ctypes.c_int64 * 3)()
            L.strsim_b200_speculation_stats(o)
            return int(o[0])

        # warm-up: the plugin learns the query's measures from the first step, and its pool of pinned result
        # buffers grows in a background thread after the first requests (a one-time cost per process)
        for k in range(4):
            plugin_step()
            if k < 2:
                time.sleep(0.6)
        barrier()
        e2e_steps = max(1, min(args.steps, 5))
        served0 = served()
        plugin_launches0 = _native.kernel_launches()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            plugin_step()
        dt = max_over_ranks(time.perf_counter() - t0)
        plugin_launches = _native.kernel_launches() - plugin_launches0
        res = plugin_step(keep=True)
        sums = [float(sum(float(v.sum()) for v in r.values())) for r in res]
        chunks_out = [len(r.arrays) for r in res]
        for r in res:
            r.release()
        plugin_driver.cache_clear()
        e2e = {"value": n * len(measures) * world * e2e_steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": input_bytes(A, B), "d2h_bytes_per_step": d2h_bytes,
               "ms_per_step": dt / e2e_steps * 1e3, "steps": e2e_steps,
               "api": "one _polars_plugin_<measure> call per measure (what a Polars query issues), pageable Arrow "
                      "buffers in, results in the plugin's own Arrow buffers; the column cache is cleared before "
                      "every step, so each step uploads both columns once and downloads every measure",
               "kernel_launches_per_step": plugin_launches / e2e_steps,
               # companion measures (include/strsim_b200.h: strsim_b200_speculation): the call that uploads the
               # columns computed and downloaded the measures the previous step's calls had asked for; this
               # many calls per step took their result ready-made
               "calls_served_from_results_computed_with_the_upload_per_step": (served() - served0) / (e2e_steps + 1),
               "warmup_steps": 4,
               "result_chunks_per_call": chunks_out,
               "checksum_matches_device": bool(all(
                   abs(s - checksums[m]) <= 1e-9 * max(1.0, abs(checksums[m])) for s, m in zip(sums, measures)))}
        if args.strong:
            e2e["devices"] = os.environ["STRSIM_B200_DEVICES"]

        # (2) one host call for all measures over pinned buffers (secondary): needs a pinned copy of the inputs
        if input_bytes(A, B) + d2h_bytes < (6 << 30) and not args.strong:
            Ap, Bp = workloads.make_pairs(wl["config"], n, row_base=rank * n, pinned=True, uneven_b=(wl["config"] == 3))
            prepared = _native.prepare(Ap, Bp)
            host_outs = [torch.empty(n, dtype=torch.float64, pin_memory=True) for _ in measures]
            host_val = torch.zeros((n + 7) // 8 + 8, dtype=torch.uint8, pin_memory=True)
            ovs, ob = [t.numpy() for t in host_outs], host_val.numpy()

            def pinned_step():
                _native.compute_host_multi(measures, None, None, out_values=ovs, out_validity=ob, prepared=prepared)

            pinned_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                pinned_step()
            dtp = max_over_ranks(time.perf_counter() - t0)
            e2e_pinned = {"value": n * len(measures) * world * e2e_steps / dtp, "unit": UNIT,
                          "ms_per_step": dtp / e2e_steps * 1e3,
                          "api": "strsim_b200_compute_host_multi: one upload of the step's two columns, one fused "
                                 "pass per row slice, results downloaded; pinned host buffers on both sides",
                          "checksum_matches_device": bool(abs(float(host_outs[-1].sum().item()) - checksum)
                                                          < 1e-6 * max(1.0, abs(checksum)))}
            del Ap, Bp, host_outs, host_val, prepared

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    if fused:
        # the step IS one launch of the fused kernel: views + payload read once, one f64 per measure written
        dominant, dom_key = "fused:" + "+".join(measures), "fused"
        dom_s = elapsed_ms / args.steps * 1e-3
        dom_bytes = alg_bytes + 8 * n * (len(measures) - 1)
    else:
        dominant = dom_key = max(per_measure_ms, key=per_measure_ms.get)
        dom_s = per_measure_ms[dominant] * 1e-3
        dom_bytes = alg_bytes
    traffic = warp_inst = None
    traffic_note = None
    tp = ROOT / "profiles" / "traffic.json"
    src_hash = kernel_source_hash()
    prof = json.loads(tp.read_text()) if tp.exists() else {}
    if n == wl["rows"]:  # the ncu capture was taken at the workload's full size
        if prof.get("source_hash") == src_hash:
            traffic = prof.get(args.workload, {}).get(dom_key)
            warp_inst = prof.get(args.workload + "_warp_instructions", {}).get(dom_key)
        else:
            traffic_note = (f"profiles/traffic.json was captured on sources {prof.get('source_hash')}, this run's "
                            f"sources hash to {src_hash}: the ncu figures are not quoted")
    roofline = {"bound": "hbm", "kernel": f"short_kernel<{dominant}>", "achieved": dom_bytes / dom_s / 1e9,
                "peak": peak, "unit": "GB/s", "frac": dom_bytes / dom_s / 1e9 / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes,
                "algorithmic_bytes_per_pair": dom_bytes / n, "launch_ms": dom_s * 1e3,
                # SURVEY.md 8(d) counts every kernel of the measure incl. the pre-pass: the column statistics
                # run once per upload (not per step), so they are timed separately and BOTH fractions are given
                "stats_prepass_ms": stats_ms,
                "frac_incl_stats_prepass": dom_bytes / (dom_s + stats_ms * 1e-3) / 1e9 / peak,
                "kernel_source_hash": src_hash}
    if traffic_note:
        roofline["traffic_note"] = traffic_note
    if warp_inst and clocks and clocks.get("sm_mhz"):
        # what actually bounds the kernel (ncu: math-pipe throttle / not-selected stalls): the issue rate.
        # Ceilings are MEASURED here with bench_support/peaks.cu (MEASURED_PEAKS.json has no integer peak):
        # LOP3-only = the alu pipe (LOP3/SHF/PRMT/IADD3), LOP3+IMAD mix = both integer-capable pipes
        ipc = warp_inst / (148 * dom_s * clocks["sm_mhz"] * 1e6)
        roofline["issue"] = {"warp_instructions_per_launch": warp_inst, "ipc_per_sm": ipc,
                             "note": "SM integer pipes bound, not HBM; warp instructions per launch from the "
                                     "ncu capture in profiles/ (same kernel sources, see kernel_source_hash)"}
        try:
            from bench_support import peaks as int_peaks

            pk = int_peaks.measure(clocks["sm_mhz"])
            roofline["issue"].update({
                "ipc_peak_alu_pipe_measured": pk["lop3"]["per_clk_per_sm"],
                "ipc_peak_alu_plus_fma_measured": pk["lop3_imad_mix"]["per_clk_per_sm"],
                "frac_of_alu_plus_fma_peak": ipc / pk["lop3_imad_mix"]["per_clk_per_sm"]})
        except Exception as exc:  # the microbenchmark is evidence, not the product
            roofline["issue"]["peak_error"] = str(exc)
    per_measure = {m: {"ms": ms, "pairs_per_s": n / (ms * 1e-3), "gbps": alg_bytes / (ms * 1e-3) / 1e9,
                       "hbm_frac": alg_bytes / (ms * 1e-3) / 1e9 / peak} for m, ms in per_measure_ms.items()}
    line = {
        "metric": METRIC,
        "value": n * len(measures) * world * args.steps / (elapsed_ms * 1e-3),
        "unit": UNIT, "n_gpus": args.gpus if args.strong else world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if args.strong else "weak", "vs_baseline": None,
        "dtype": DTYPE, "data": "synthetic",
        "config": run_config(args, wl, world),
        "notes": {"host_numa_node_rank0": numa_node,
                  "l2": "inputs per launch (views+payload %.0f MB) exceed the 126 MB L2; no flush needed" % (alg_bytes / 1e6),
                  "value_is": "device-resident: both columns (and their byte statistics) already in HBM"},
        "per_measure": per_measure, "roofline": roofline, "clocks": clocks, "gpu_launches": launches,
        "overflow_rows_last_call": {"to_64bit_kernel": overflow[0], "to_long_kernel": overflow[1]},
        "checksum": checksum, "checksums": checksums, "checksums_all_ranks": checksums_all_ranks,
        "step": ("one fused launch for all measures (strsim_b200_compute_device_multi)" if fused
                 else "one single-measure launch per measure"),
        "fused_matches_single_measure_kernel": fused_matches_single,
        "per_measure_note": "each measure's own single-measure kernel, timed separately after the step loop",
    }
    if cells is not None:
        ms = per_measure_ms["levenshtein"]
        line["long_levenshtein"] = {"cells_per_launch": cells, "gcups": cells / (ms * 1e-3) / 1e9,
                                    "note": "cells = sum la*lb (codepoints) over pairs with a != b"}
        # the roofline of this kernel is the SM integer issue rate, not HBM: warp
        # instructions per cell from the ncu capture in profiles/, ceilings measured here (peaks.cu)
        per_cell = prof.get("C4_warp_instructions_per_cell", {}).get("long_lev_kernel") \
            if prof.get("source_hash") == src_hash else None
        if per_cell and clocks and clocks.get("sm_mhz"):
            ipc = cells * per_cell / (148 * ms * 1e-3 * clocks["sm_mhz"] * 1e6)
            issue = {"warp_instructions_per_cell": per_cell, "ipc_per_sm": ipc}
            try:
                from bench_support import peaks as int_peaks

                pk = int_peaks.measure(clocks["sm_mhz"])
                issue.update({"ipc_peak_alu_pipe_measured": pk["lop3"]["per_clk_per_sm"],
                              "ipc_peak_alu_plus_fma_measured": pk["lop3_imad_mix"]["per_clk_per_sm"],
                              "frac_of_alu_plus_fma_peak": ipc / pk["lop3_imad_mix"]["per_clk_per_sm"]})
            except Exception as exc:
                issue["peak_error"] = str(exc)
            line["long_levenshtein"]["issue"] = issue
    if e2e:
        line["e2e"] = e2e
    if e2e_pinned:
        line["e2e_pinned"] = e2e_pinned
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(A, B, measures)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
