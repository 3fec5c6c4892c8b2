#!/bin/bash
# ncu evidence on the GPU box (run under gpurun, ONE workload per call: gpurun brings back at most 64 MiB
# and a --set full capture of six launches with sources is ~50 MB):
#   gpurun --timeout 900 -- 'bash tools/capture_profiles.sh C2'
# launch list of the bench command (cold-cache, serialised times: shares only), then full counters of the
# timed fused launch and the single-measure launches after it.
set -u
W=${1:-C2}; R=${2:-r1}; O=gpurun_out
mkdir -p $O
if [ "$W" = C2 ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/${R}_launches_C2.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:short_kernel -s 3 -c 6 -f -o $O/${R}_prof_short_C2 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_full.log 2>&1
  python tools/summarize_ncu.py $O/${R}_prof_short_C2.ncu-rep C2 $O/${R}_ncu_short_kernel_C2.md profiles/traffic.json
  cp profiles/traffic.json $O/traffic.json
elif [ "$W" = C3 ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/${R}_launches_C3.csv python bench.py --workload C3 --rows 20000000 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_under_ncu3.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:short_kernel -s 8 -c 3 -f -o $O/${R}_prof_short_C3 python bench.py --workload C3 --rows 20000000 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_full3.log 2>&1
elif [ "$W" = C4 ]; then
  ncu --set full --clock-control none --import-source on -k regex:long_lev_kernel -c 1 -f -o $O/${R}_prof_long_C4 python bench.py --workload C4 --rows 40000 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_full4.log 2>&1
fi
ls -la $O | tail -8
