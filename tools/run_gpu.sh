cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q -k "not full_size and not c4_by" > gpurun_out/r2m_pytest.log 2>&1; tail -6 gpurun_out/r2m_pytest.log
python bench.py --workload C3 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2m_bench_C3.json 2> gpurun_out/r2m_err.log; python - <<PY
import json
d=json.load(open('gpurun_out/r2m_bench_C3.json'))
print('C3', round(d['ms_per_step'],4), {k:round(v['ms'],3) for k,v in d['per_measure'].items()}, d['overflow_rows_last_call'], d['roofline']['frac'])
PY
tail -3 gpurun_out/r2m_err.log
