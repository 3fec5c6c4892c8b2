#!/usr/bin/env python
"""bench.py -- string pairs/sec of the row-wise similarity hot path on B200 (BASELINE.json metric).

A step = one pass of ALL FIVE measures (levenshtein, jaro, jaro_winkler, jaccard, sorensen_dice)
over one batch of synthetic pairs (default: BASELINE config C2, 10M ASCII name pairs of length
4..24 per GPU, seeds in SURVEY.md 8(d)), evaluated by ONE fused kernel launch
(`strsim_b200_compute_device_multi`: views and bytes read once, one position mask per character
feeding all measures).  `value` counts pair evaluations (rows x 5) per second with the two columns
already resident in HBM; `per_measure` times each measure's own single-measure kernel the same way
(the BASELINE metric is quoted per measure); `e2e` is the step through the host-buffer C ABI call
(`strsim_b200_compute_host_multi`): pinned host memory in, ONE H2D upload of the step's two
columns, the fused kernel, and the D2H of every measure's results inside the timed region.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C2|C3|C4] [--rows R]
    python bench.py --impl reference ...     # the reference's CPU algorithm (oracle port) on host cores
    torchrun --nproc-per-node N bench.py --gpus N ...   # one process per GPU, rows sharded by range

Multi-GPU is row-range sharding with no data-path collective: every rank owns
`rows` rows (weak scaling); NCCL is used only for the barrier and the max-over-ranks of the time.
"""
from __future__ import annotations

import argparse
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
    "C5": dict(config=5, rows=125_000_000, measures=("jaro_winkler", "sorensen_dice"),
               desc="C5: record-linkage pairs (C2 generator, seed 0xC5), Jaro-Winkler + Sorensen-Dice"),
}
UNIT = "pairs/s"


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


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


def dist_setup(n_gpus: int):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return world, rank, local


def cpu_baseline(A, B, measures, budget_s: float = 15.0):
    """The oracle (a port of the reference's algorithms with its static row-range threading,
    strsim.rs:21-39,72-100) on this box's host cores, on a bounded sample of the same workload."""
    import pyarrow as pa
    from oracle import oracle

    threads = oracle.n_host_threads()
    n = len(A)
    flatA = A.combine_chunks() if isinstance(A, pa.ChunkedArray) else A
    flatB = B.combine_chunks() if isinstance(B, pa.ChunkedArray) else B
    probe = min(n, 50_000)
    t0 = time.perf_counter()
    for m in measures:
        oracle.batch_views(m, flatA.slice(0, probe), flatB.slice(0, probe), n_threads=threads)
    rate = probe * len(measures) / max(time.perf_counter() - t0, 1e-6)
    sample = int(min(n, max(probe, rate * budget_s / len(measures))))
    per = {}
    t_all = 0.0
    for m in measures:
        t0 = time.perf_counter()
        oracle.batch_views(m, flatA.slice(0, sample), flatB.slice(0, sample), n_threads=threads)
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
    host threads.  Each step is a bounded sample of the workload."""
    if rank != 0:
        return
    from bench_support import workloads
    from oracle import oracle

    measures = wl["measures"]
    threads = oracle.n_host_threads()
    sample = min(args.rows, 2_000_000 if wl["config"] != 4 else 2_000)
    A, B = workloads.make_pairs(wl["config"], sample)
    import pyarrow as pa

    if isinstance(A, pa.ChunkedArray):
        A, B = A.combine_chunks(), B.combine_chunks()
    for _ in range(args.warmup):
        for m in measures:
            oracle.batch_views(m, A, B, n_threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for m in measures:
            oracle.batch_views(m, A, B, n_threads=threads)
    dt = time.perf_counter() - t0
    value = sample * len(measures) * args.steps / dt
    line = {
        "impl": "reference", "metric": "string pairs/sec (mean over the measures of the workload)",
        "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/u32 codepoints, f64 results", "data": "synthetic",
        "config": {"workload": wl["desc"], "rows_per_step": sample, "measures": list(measures)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} rows x {len(measures)} measures per step"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
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
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.rows is None:
        args.rows = wl["rows"]
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    world, rank, local = dist_setup(args.gpus)

    if args.impl == "reference":
        run_reference(args, wl, world, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    from bench_support import workloads
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
    # ---- data: this rank's row range [rank*n, (rank+1)*n), generated into pinned host memory --------
    A, B = workloads.make_pairs(wl["config"], n, row_base=rank * n, pinned=not args.no_e2e,
                                uneven_b=(wl["config"] == 3))
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
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream

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
    elapsed_ms = e0.elapsed_time(e1)
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    checksum = float(out.sum().item())
    checksums = {m: float(o.sum().item()) for m, o in zip(measures, outs)}

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

    # ---- end to end through the host-buffer C ABI (pinned in, pinned out) -------------------------------
    e2e = None
    if not args.no_e2e:
        prepared = _native.prepare(A, B)
        host_outs = [torch.empty(n, dtype=torch.float64, pin_memory=True) for _ in measures]
        host_val = torch.zeros((n + 7) // 8 + 8, dtype=torch.uint8, pin_memory=True)
        ovs, ob = [t.numpy() for t in host_outs], host_val.numpy()
        host_out = host_outs[-1]

        def e2e_step():
            # ONE host->device upload of this step's two columns, then every measure of the step;
            # each measure's 8n result bytes come back to (pinned) host memory
            _native.compute_host_multi(measures, None, None, out_values=ovs, out_validity=ob, prepared=prepared)

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        e2e_steps = max(1, min(args.steps, 3))
        for _ in range(e2e_steps):
            e2e_step()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        seen, in_bytes = set(), 0  # chunks made by slicing share buffers: the library uploads each once
        for col in (A, B):
            for ch in (col.chunks if hasattr(col, "chunks") else [col]):
                bufs = ch.buffers()
                in_bytes += 16 * len(ch) + ((len(ch) + 7) // 8 if bufs[0] is not None else 0)
                for b in bufs[2:]:
                    if b is not None and b.address not in seen:
                        seen.add(b.address)
                        in_bytes += b.size
        e2e = {"value": n * len(measures) * world * e2e_steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": in_bytes,
               "d2h_bytes_per_step": (8 * n + ((n + 7) // 8 if has_nulls else 0)) * len(measures),
               "ms_per_step": dt / e2e_steps * 1e3,
               "api": "strsim_b200_compute_host_multi: one upload of the step's two columns, all measures of the "
                      "step (one fused pass per row slice), results downloaded (pinned host buffers)",
               "checksum_matches_device": bool(abs(float(host_out.sum().item()) - checksum) < 1e-6 * max(1.0, abs(checksum)))}

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
    tp = ROOT / "profiles" / "traffic.json"
    if tp.exists() and n == wl["rows"]:  # the ncu capture was taken at the workload's full size
        prof = json.loads(tp.read_text())
        traffic = prof.get(args.workload, {}).get(dom_key)
        warp_inst = prof.get(args.workload + "_warp_instructions", {}).get(dom_key)
    roofline = {"bound": "hbm", "kernel": f"short_kernel<{dominant}>", "achieved": dom_bytes / dom_s / 1e9,
                "peak": peak, "unit": "GB/s", "frac": dom_bytes / dom_s / 1e9 / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes,
                "algorithmic_bytes_per_pair": dom_bytes / n, "launch_ms": dom_s * 1e3}
    if warp_inst and clocks and clocks.get("sm_mhz"):
        # what actually bounds the kernel (ncu: math-pipe throttle / not-selected stalls): the issue rate.
        # Ceilings are MEASURED here with bench_support/peaks.cu (MEASURED_PEAKS.json has no integer peak):
        # LOP3-only = the alu pipe (LOP3/SHF/PRMT/IADD3), LOP3+IMAD mix = both integer-capable pipes
        ipc = warp_inst / (148 * dom_s * clocks["sm_mhz"] * 1e6)
        roofline["issue"] = {"warp_instructions_per_launch": warp_inst, "ipc_per_sm": ipc,
                             "note": "SM integer pipes bound, not HBM; warp instructions per launch from the "
                                     "ncu capture in profiles/"}
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
        "metric": "string pairs/sec (pair evaluations = rows x measures; per-measure rates in per_measure)",
        "value": n * len(measures) * world * args.steps / (elapsed_ms * 1e-3),
        "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/u32 codepoints, u32/u64 bit-vectors, f64 results", "data": "synthetic",
        "config": {"workload": wl["desc"], "rows_per_gpu": n, "measures": list(measures),
                   "parallelism": f"row-range x{world}, no collective",
                   "host_numa_node_rank0": numa_node,
                   "l2": "inputs per launch (views+payload %.0f MB) exceed the 126 MB L2; no flush needed" % (alg_bytes / 1e6)},
        "per_measure": per_measure, "roofline": roofline, "clocks": clocks, "gpu_launches": launches,
        "overflow_rows_last_call": {"to_64bit_kernel": overflow[0], "to_long_kernel": overflow[1]},
        "checksum": checksum, "checksums": checksums,
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
        per_cell = json.loads(tp.read_text()).get("C4_warp_instructions_per_cell", {}).get("long_lev_kernel") \
            if tp.exists() else None
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
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(A, B, measures)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
