#!/usr/bin/env python
"""Key counters + warp-state (stall) breakdown from an `ncu --page raw --csv` dump.   python tools/ncu_raw_summary.py raw.csv [raw2.csv ...]"""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__warps_active.avg.per_cycle_active"]
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    H, U, V = rows[0], rows[1], rows[2]
    d = {h: (V[i], U[i]) for i, h in enumerate(H)}
    print("==", path, d.get("Kernel Name", ("", ""))[0][:100])
    for k in KEYS:
        if k in d:
            print(f"  {k:75s} {d[k][0]:>16s} {d[k][1]}")
    stalls = []
    for h, (v, u) in d.items():
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
            try:
                stalls.append((float(v), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    tot = sum(s for s, _ in stalls) or 1.0
    print("  stalls (warps per issue-active cycle):", ", ".join(f"{n}={s:.2f}" for s, n in sorted(stalls, reverse=True)[:10]))
