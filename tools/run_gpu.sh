cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q -k "not full_size and not c4_by" > gpurun_out/r2p_pytest.log 2>&1; tail -4 gpurun_out/r2p_pytest.log
for w in C3 M1; do
python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2p_bench_$w.json 2> gpurun_out/r2p_err.log; python - <<PY
import json
d=json.load(open('gpurun_out/r2p_bench_$w.json'))
print('$w', round(d['ms_per_step'],4), {k:round(v['ms'],3) for k,v in d['per_measure'].items()}, d['overflow_rows_last_call'], d['roofline']['frac'])
PY
done
tail -3 gpurun_out/r2p_err.log
