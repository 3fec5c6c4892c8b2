#!/bin/bash
# Bench lines of one round (outside any profiler) on a 1-GPU box:
#   gpurun --timeout 2400 -- 'bash tools/capture_round_r2.sh r2'
# Run tools/capture_profiles_r2.sh FIRST so that bench.py finds profiles/traffic.json of the same kernel sources.
set -u
R=${1:-r2}; O=gpurun_out
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p $O
python bench.py --steps 20 --warmup 5 > $O/${R}_bench_C2.json 2> $O/${R}_bench_err.log
python bench.py --impl reference --steps 5 --warmup 1 > $O/${R}_bench_C2_reference.json 2>> $O/${R}_bench_err.log
python bench.py --workload C3 --steps 5 > $O/${R}_bench_C3_100M.json 2>> $O/${R}_bench_err.log
python bench.py --workload C4 --steps 3 --no-e2e > $O/${R}_bench_C4_1M.json 2>> $O/${R}_bench_err.log
python bench.py --workload C5 --steps 5 > $O/${R}_bench_C5_125M.json 2>> $O/${R}_bench_err.log
for w in L1 M1 N1 T1; do
  python bench.py --workload $w --steps 10 --no-cpu-baseline --no-e2e > $O/${R}_bench_$w.json 2>> $O/${R}_bench_err.log
done
# the README scenario through the plugin symbols, call by call (cache / companion measures on and off)
python tools/plugin_e2e.py --pageable > $O/${R}_plugin_e2e.jsonl 2>> $O/${R}_bench_err.log
STRSIM_B200_SPECULATE=0 python tools/plugin_e2e.py --pageable >> $O/${R}_plugin_e2e.jsonl 2>> $O/${R}_bench_err.log
STRSIM_B200_CACHE=0 python tools/plugin_e2e.py --pageable >> $O/${R}_plugin_e2e.jsonl 2>> $O/${R}_bench_err.log
python tools/plugin_e2e.py >> $O/${R}_plugin_e2e.jsonl 2>> $O/${R}_bench_err.log
STRSIM_B200_TRACE=1 python tools/plugin_e2e.py --pageable 2>&1 >/dev/null | grep "strsim trace" | tail -8 > $O/${R}_e2e_timeline_C2.txt
tail -5 $O/${R}_bench_err.log
for f in $O/${R}_bench_*.json; do python - "$f" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
e = d.get("e2e") or {}
print(sys.argv[1].split("/")[-1], "value %.4g" % d["value"], "ms/step %.3f" % d["ms_per_step"], "frac", round((d.get("roofline") or {}).get("frac", 0), 4),
      "e2e %.4g (%.2f ms)" % (e.get("value", 0), e.get("ms_per_step", 0)), "cpu %.4g" % (d.get("cpu_baseline") or {}).get("value", 0),
      (d.get("long_levenshtein") or {}).get("gcups", ""))
PY
done
cat $O/${R}_plugin_e2e.jsonl | cut -c1-260
