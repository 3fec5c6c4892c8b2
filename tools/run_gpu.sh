cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q -k "wide_rows or long_rows or long_lev or medium or M1 or fused_measures_subsets" 2>&1 | tail -5
for w in T1 M1; do python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/err_$w.log | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$w', round(d['ms_per_step'],3), [round(v['ms'],3) for v in d.get('per_measure',{}).values()], d.get('gpu_launches'))
"; done
STRSIM_B200_WIDE_ROWS=0 python bench.py --workload T1 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/err_T1b.log | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('T1-nowide', round(d['ms_per_step'],3), [round(v['ms'],3) for v in d.get('per_measure',{}).values()], d.get('gpu_launches'))
"
