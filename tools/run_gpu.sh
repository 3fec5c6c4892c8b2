cd $GRAFT_REPO_ROOT
O=gpurun_out; mkdir -p $O
show() { python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$1', round(d['ms_per_step'],4), [round(v['ms'],3) for v in d.get('per_measure',{}).values()], d.get('gpu_launches'), d['checksum'], (d.get('long_levenshtein') or {}).get('gcups',''))
"; }
timeout 300 python -m pytest tests -m gpu -x -q -k "wide_rows or medium or long_rows or tiles_from" 2>&1 | tail -2
timeout 300 python bench.py --workload T1 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>$O/err_T1.log | show T1
timeout 300 python bench.py --workload M1 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>$O/err_M1.log | show M1
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e 2>$O/err_C2.log | show C2
