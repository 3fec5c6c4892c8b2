#!/bin/bash
# Round evidence, bench lines (outside any profiler) on the GPU box:
#   gpurun --timeout 1500 -- 'bash tools/capture_round.sh r1'
# The ncu captures (launch lists, --set full reports, profiles/traffic.json) come from
# tools/capture_profiles.sh, one workload per gpurun call; run that FIRST so that bench.py reads the
# current profiles/traffic.json into roofline.traffic / roofline.issue.
set -u
R=${1:-r1}
O=gpurun_out
mkdir -p $O
python bench.py > $O/${R}_bench_C2.json 2> $O/bench_err.log
python bench.py --impl reference --steps 3 --warmup 1 > $O/${R}_bench_C2_reference.json 2>> $O/bench_err.log
python bench.py --workload C3 --steps 5 > $O/${R}_bench_C3_100M.json 2>> $O/bench_err.log
python bench.py --workload C4 --steps 3 --no-e2e > $O/${R}_bench_C4_1M.json 2>> $O/bench_err.log
python bench.py --workload C5 --steps 5 --no-cpu-baseline > $O/${R}_bench_C5_125M.json 2>> $O/bench_err.log
python bench.py --workload L1 --steps 10 --no-cpu-baseline --no-e2e > $O/${R}_bench_L1.json 2>> $O/bench_err.log
# end-to-end timeline of one host call (upload landed / kernels / download done per row slice)
STRSIM_B200_TRACE=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "strsim trace" | tail -6 > $O/${R}_e2e_timeline_C2.txt
# the README scenario through the plugin symbols: five separate calls, with / without the column cache
python tools/plugin_e2e.py > $O/${R}_plugin_e2e.jsonl 2>> $O/bench_err.log
STRSIM_B200_CACHE=0 python tools/plugin_e2e.py >> $O/${R}_plugin_e2e.jsonl 2>> $O/bench_err.log
STRSIM_B200_STAGED_D2H=0 python tools/plugin_e2e.py >> $O/${R}_plugin_e2e.jsonl 2>> $O/bench_err.log
python tools/plugin_e2e.py --pageable >> $O/${R}_plugin_e2e.jsonl 2>> $O/bench_err.log
STRSIM_B200_CACHE=0 python tools/plugin_e2e.py --pageable >> $O/${R}_plugin_e2e.jsonl 2>> $O/bench_err.log
tail -5 $O/bench_err.log
for f in $O/${R}_bench_*.json; do echo $f; head -c 300 $f; echo; done
cat $O/${R}_plugin_e2e.jsonl
