cd $GRAFT_REPO_ROOT
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/s7_pytest.log; tail -3 $O/s7_pytest.log
show() { python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$1', round(d['ms_per_step'],4), [round(v['ms'],3) for v in d.get('per_measure',{}).values()], d.get('gpu_launches'), d['checksum'], (d.get('long_levenshtein') or {}).get('gcups',''))
"; }
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e 2>$O/err_C2.log | show C2
timeout 300 python bench.py --workload C3 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>$O/err_C3.log | show C3
timeout 300 python bench.py --workload C5 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>$O/err_C5.log | show C5
for w in N1 L1 M1 T1; do
timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>$O/err_$w.log | show $w
done
