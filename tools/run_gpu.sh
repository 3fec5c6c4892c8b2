cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/r2r_pytest.log 2>&1; tail -3 gpurun_out/r2r_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for lib in "" exp/libvariant_ulat_rpt2.so; do
STRSIM_B200_LIB=$lib python bench.py --workload C3 --rows 50000000 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2r_c3.json 2> gpurun_out/r2r_err.log; python - <<PY
import json
d=json.load(open('gpurun_out/r2r_c3.json'))
print('C3 50M lib=[$lib]', round(d['ms_per_step'],4))
PY
done
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C2', d['ms_per_step'], d['roofline']['traffic'])"
