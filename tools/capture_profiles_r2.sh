#!/bin/bash
# ncu evidence of one round, summarised ON the GPU box (an .ncu-rep with sources is ~60 MB and gpurun brings
# back at most 64 MiB):   gpurun --timeout 1500 -- 'bash tools/capture_profiles_r2.sh r2'
#   * launch lists of the bench commands (cold-cache, serialised times: shares only)
#   * --set full counters of the fused C2 launch + the five single-measure launches  -> <R>_ncu_short_kernel_C2.md,
#     profiles/traffic.json (stamped with the hash of the kernel sources), per-phase / per-line shares
#   * the same for the two launches of a C3 segment and for long_lev_kernel on C4 (40,000 pairs)
set -u
R=${1:-r2}; O=gpurun_out; T=/tmp/rep
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p $O $T
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/${R}_launches_C2.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/${R}_bench_under_ncu_C2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 90 --csv --log-file $O/${R}_launches_C3.csv python bench.py --workload C3 --rows 20000000 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $O/${R}_bench_under_ncu_C3.log 2>&1
# C2: the timed fused launch and the per-measure pass behind it
ncu --set full --clock-control none --import-source on -k regex:short_kernel -s 3 -c 6 -f -o $T/c2 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $O/${R}_ncu_full_C2.log 2>&1
python tools/summarize_ncu.py $T/c2.ncu-rep C2 $O/${R}_ncu_short_kernel_C2.md profiles/traffic.json
python tools/ncu_regions.py $T/c2.ncu-rep "(int)15, (int)256" short_kernelIjLi15ELi256ELi2ELb0ELi32ELb1ELb1ELb0ELb0E polars-strsim_b200/csrc/host.o > $O/${R}_ncu_regions_C2_fused.txt 2>&1
python tools/ncu_lines.py $T/c2.ncu-rep 0 short_kernelIjLi15ELi256ELi2ELb0ELi32ELb1ELb1ELb0ELb0E polars-strsim_b200/csrc/host.o "(int)15, (int)256" > $O/${R}_ncu_lines_C2_fused.txt 2>&1
ncu -i $T/c2.ncu-rep --page raw --csv > $O/${R}_ncu_raw_C2.csv 2>/dev/null
python tools/ncu_raw_summary.py $O/${R}_ncu_raw_C2.csv > $O/${R}_ncu_stalls_C2.txt 2>&1
# C3: the Latin-1 launch and the register-compare (gather) launch of the first segment
ncu --set full --clock-control none --import-source on -k regex:short_kernel -c 2 -f -o $T/c3 python tools/prof_one.py C3 fused 20000000 1 > $O/${R}_ncu_full_C3.log 2>&1
python tools/summarize_ncu.py $T/c3.ncu-rep C3 $O/${R}_ncu_short_kernel_C3.md
ncu -i $T/c3.ncu-rep --page raw --csv > $O/${R}_ncu_raw_C3.csv 2>/dev/null
python tools/ncu_raw_summary.py $O/${R}_ncu_raw_C3.csv > $O/${R}_ncu_stalls_C3.txt 2>&1
python tools/ncu_regions.py $T/c3.ncu-rep "(bool)1, (bool)1>" short_kernelIjLi15ELi256ELi2ELb0ELi128ELb0ELb0ELb1ELb1E polars-strsim_b200/csrc/host.o > $O/${R}_ncu_regions_C3_latin1_launch.txt 2>&1
# C4: long_lev_kernel on 40,000 pairs (cells from the bench line of the same rows)
python bench.py --workload C4 --rows 40000 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $O/${R}_bench_C4_40k.json 2> $O/${R}_c4.err
CELLS=$(python -c "import json; print(json.load(open('$O/${R}_bench_C4_40k.json'))['long_levenshtein']['cells_per_launch'])")
ncu --set full --clock-control none --import-source on -k regex:long_lev_kernel -c 1 -f -o $T/c4 python tools/prof_one.py C4 levenshtein 40000 1 > $O/${R}_ncu_full_C4.log 2>&1
python tools/summarize_ncu.py $T/c4.ncu-rep C4 $O/${R}_ncu_long_lev_C4.md profiles/traffic.json $CELLS
ncu -i $T/c4.ncu-rep --page raw --csv > $O/${R}_ncu_raw_C4.csv 2>/dev/null
python tools/ncu_raw_summary.py $O/${R}_ncu_raw_C4.csv > $O/${R}_ncu_stalls_C4.txt 2>&1
python tools/ncu_lines.py $T/c4.ncu-rep 0 long_lev_kernel polars-strsim_b200/csrc/host.o > $O/${R}_ncu_lines_C4_long_lev.txt 2>&1
# T1: the ten-word plane launch over the 65-320-byte rows of an ASCII column (4 M rows, a tenth of them long)
ncu --set full --clock-control none --import-source on -k regex:short_kernel -s 2 -c 1 -f -o $T/t1 python tools/prof_one.py T1 fused 4000000 1 > $O/${R}_ncu_full_T1.log 2>&1
python tools/summarize_ncu.py $T/t1.ncu-rep T1 $O/${R}_ncu_wide_rows_T1.md
ncu -i $T/t1.ncu-rep --page raw --csv > $O/${R}_ncu_raw_T1.csv 2>/dev/null
python tools/ncu_raw_summary.py $O/${R}_ncu_raw_T1.csv > $O/${R}_ncu_stalls_T1.txt 2>&1
rm -f $O/${R}_ncu_raw_T1.csv
cp profiles/traffic.json $O/traffic.json
rm -f $O/${R}_ncu_raw_C2.csv $O/${R}_ncu_raw_C3.csv $O/${R}_ncu_raw_C4.csv
ls -la $O | tail -30; cat $O/traffic.json
