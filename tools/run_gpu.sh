set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q -k "long_rows or medium_and_long or boundaries or fallback or fused_measures_sub" > gpurun_out/r2g_pytest.log 2>&1; tail -15 gpurun_out/r2g_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for w in T1 C2; do
python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2g_bench_$w.json 2> gpurun_out/r2g_bench_$w.err; python - <<PY
import json
d=json.load(open('gpurun_out/r2g_bench_$w.json'))
print('$w', d['ms_per_step'], {k:round(v['ms'],3) for k,v in d['per_measure'].items()}, d['overflow_rows_last_call'], d['roofline']['stats_prepass_ms'])
PY
done
