#!/usr/bin/env python
"""Markdown summary + traffic.json entries from an `ncu --set full` report.

    python tools/summarize_ncu.py REPORT.ncu-rep WORKLOAD OUT.md [profiles/traffic.json]
"""
import csv
import io
import json
import re
import subprocess
import sys
from pathlib import Path

MEASURES = ("levenshtein", "jaro", "jaro_winkler", "jaccard", "sorensen_dice")
KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed.avg.per_cycle_elapsed", "IPC per SM (max 4)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / instruction"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "resident warps % of 64"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM (smem limit)"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank-conflict wavefronts"),
    ("sm__cycles_elapsed.max", "SM cycles"),
]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    rep, workload, out_md = sys.argv[1:4]
    traffic_path = Path(sys.argv[4]) if len(sys.argv) > 4 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    lines = [f"# ncu --set full summary: {Path(rep).name} ({workload})", "",
             "Numbers under the profiler are cold-cache and serialised: use them for shares and counters,",
             "not as bench values (bench.py times with CUDA events outside the profiler).", ""]
    traffic = {}
    insts = {}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]]
        lines.append(f"## `{name}`")
        lines.append("")
        lines.append("| metric | value |")
        lines.append("|---|---|")
        for k, label in KEYS:
            if k in ix:
                lines.append(f"| {label} (`{k}`) | {r[ix[k]]} {units[ix[k]]} |")
        lines.append("")
        m = re.search(r"short_kernel<unsigned int, (\d+)", name)
        if m and "dram__bytes_read.sum" in ix:
            t = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]]) + \
                to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
            # template value 0..4: one measure; 8 + group mask: the fused kernel (pair_algos.cuh MULTI_BASE)
            key = MEASURES[int(m.group(1))] if int(m.group(1)) < 5 else "fused"
            traffic[key] = t
            insts[key] = float(r[ix["smsp__inst_executed.sum"]].replace(",", ""))
    Path(out_md).write_text("\n".join(lines) + "\n")
    if traffic_path and len(sys.argv) > 5:  # C4: argv[5] = DP cells of the captured launch
        cells = float(sys.argv[5])
        for r in rows[2:]:
            if "long_lev_kernel" in r[ix["Kernel Name"]]:
                sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
                from bench import kernel_source_hash

                cur = json.loads(traffic_path.read_text()) if traffic_path.exists() else {}
                if cur.get("source_hash") != kernel_source_hash():
                    cur = {"source_hash": kernel_source_hash()}
                inst = float(r[ix["smsp__inst_executed.sum"]].replace(",", ""))
                t = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]]) + \
                    to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
                cur["C4_warp_instructions_per_cell"] = {"long_lev_kernel": inst / cells,
                                                        "_note": "smsp__inst_executed.sum / DP cells of the captured launch"}
                cur["C4_dram_bytes_per_cell"] = {"long_lev_kernel": t / cells}
                traffic_path.write_text(json.dumps(cur, indent=1) + "\n")
    if traffic_path and traffic:
        cur = json.loads(traffic_path.read_text()) if traffic_path.exists() else {}
        # the capture belongs to the kernel sources it was taken on: bench.py quotes it only for the same hash
        sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
        from bench import kernel_source_hash

        if cur.get("source_hash") != kernel_source_hash():
            cur = {"source_hash": kernel_source_hash()}
        cur.setdefault(workload, {}).update(traffic)
        cur.setdefault(workload + "_warp_instructions", {}).update(insts)
        cur["_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch of short_kernel<measure>, "
                        "from one ncu --set full capture of `python bench.py --workload <W> --steps 1 --warmup 3`")
        traffic_path.write_text(json.dumps(cur, indent=1) + "\n")
    print("wrote", out_md, traffic)


if __name__ == "__main__":
    main()
