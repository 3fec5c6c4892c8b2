set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q -k "plugin or pageable or concurrent or nulls or progressive or abi" > gpurun_out/r2c_pytest.log 2>&1; tail -5 gpurun_out/r2c_pytest.log
STRSIM_B200_TRACE=1 python tools/plugin_e2e.py 10000000 --pageable > gpurun_out/r2c_trace.json 2> gpurun_out/r2c_trace.log
tail -c 400 gpurun_out/r2c_trace.json
tail -48 gpurun_out/r2c_trace.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_C2.json 2> gpurun_out/r2c_bench_err.log; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c_bench_C2.json'))
print(d['ms_per_step'], d['roofline']['stats_prepass_ms'], d['e2e']['ms_per_step'], d['e2e_pinned']['ms_per_step'], d['e2e']['checksum_matches_device'])
PY
tail -3 gpurun_out/r2c_bench_err.log
