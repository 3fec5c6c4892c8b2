cd $GRAFT_REPO_ROOT
for r in 4 6 8; do
STRSIM_B200_WT_RPT=$r python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2i_bench_C2_rpt$r.json 2> gpurun_out/r2i_err.log; python - <<PY
import json
d=json.load(open('gpurun_out/r2i_bench_C2_rpt$r.json'))
print('C2 wt rpt=$r', round(d['ms_per_step'],4), d['overflow_rows_last_call'])
PY
done
tail -2 gpurun_out/r2i_err.log
