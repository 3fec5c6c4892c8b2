cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q -k "not c4_by" > gpurun_out/r2n_pytest.log 2>&1; tail -6 gpurun_out/r2n_pytest.log
for rb in 0 1; do
STRSIM_B200_READBACK=$rb python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2n_bench_C2_rb$rb.json 2> gpurun_out/r2n_err.log; python - <<PY
import json
d=json.load(open('gpurun_out/r2n_bench_C2_rb$rb.json'))
print('C2 readback=$rb', round(d['ms_per_step'],4), {k:round(v['ms'],3) for k,v in d['per_measure'].items()}, d['roofline']['stats_prepass_ms'], d['gpu_launches'])
PY
done
tail -2 gpurun_out/r2n_err.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2n_launches_C2.csv python tools/prof_one.py C2 fused 10000000 3 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2n_launches_C2.csv')) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
H=rows[hdr]; ki=H.index('Kernel Name'); vi=H.index('Metric Value')
for r in rows[hdr+1:]:
    print(r[0], r[ki][:95], round(float(r[vi].replace(',',''))/1e3,1))
PY
