#!/bin/bash
# Round evidence on the GPU box (run under gpurun): ncu captures first (their counters feed
# profiles/traffic.json, which bench.py reads into roofline.traffic / roofline.issue), then the bench
# lines.  Everything lands in gpurun_out/ and is copied into profiles/ by hand afterwards.
#   gpurun --timeout 1800 -- 'bash tools/capture_round.sh r1'
set -u
R=${1:-r1}
O=gpurun_out
mkdir -p $O
# 1. launch list of the default bench command (cold-cache, serialised times: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/${R}_launches_C2.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
# 2. full counters: the timed fused launch (4th short_kernel launch) + the five single-measure kernels after it
ncu --set full --clock-control none --import-source on -k regex:short_kernel -s 3 -c 6 -f -o $O/${R}_prof_short_C2 \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_full.log 2>&1
python tools/summarize_ncu.py $O/${R}_prof_short_C2.ncu-rep C2 $O/${R}_ncu_short_kernel_C2.md profiles/traffic.json
cp profiles/traffic.json $O/traffic.json
# 3. bench lines (outside the profiler)
python bench.py > $O/${R}_bench_C2.json 2> $O/bench_err.log
python bench.py --impl reference --steps 3 --warmup 1 > $O/${R}_bench_C2_reference.json 2>> $O/bench_err.log
python bench.py --workload C3 --rows 20000000 --steps 5 > $O/${R}_bench_C3_20M.json 2>> $O/bench_err.log
python bench.py --workload C4 --rows 200000 --steps 3 --no-e2e > $O/${R}_bench_C4_200k.json 2>> $O/bench_err.log
python bench.py --workload C5 --rows 50000000 --steps 5 --no-cpu-baseline > $O/${R}_bench_C5_50M.json 2>> $O/bench_err.log
# 4. end-to-end timeline of one host call (upload landed / kernels / download done per row slice)
STRSIM_B200_TRACE=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "strsim trace" | tail -6 > $O/${R}_e2e_timeline_C2.txt
tail -5 $O/bench_err.log
for f in $O/${R}_bench_*.json; do echo $f; head -c 400 $f; echo; done
# 5. the README scenario through the plugin symbols: five separate calls, with / without the column cache
python tools/plugin_e2e.py > $O/${R}_plugin_e2e.jsonl 2>> $O/bench_err.log
STRSIM_B200_CACHE=0 python tools/plugin_e2e.py >> $O/${R}_plugin_e2e.jsonl 2>> $O/bench_err.log
STRSIM_B200_STAGED_D2H=0 python tools/plugin_e2e.py >> $O/${R}_plugin_e2e.jsonl 2>> $O/bench_err.log
cat $O/${R}_plugin_e2e.jsonl
python tools/plugin_e2e.py --pageable >> $O/${R}_plugin_e2e.jsonl 2>> $O/bench_err.log
STRSIM_B200_CACHE=0 python tools/plugin_e2e.py --pageable >> $O/${R}_plugin_e2e.jsonl 2>> $O/bench_err.log
