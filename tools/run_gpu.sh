cd $GRAFT_REPO_ROOT
O=gpurun_out; mkdir -p $O
show() { python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$1', round(d['ms_per_step'],4), [round(v['ms'],3) for v in d.get('per_measure',{}).values()], d.get('gpu_launches'), d['checksum'], (d.get('long_levenshtein') or {}).get('gcups',''))
"; }
for v in head nshead nscur; do
 if [ $v = main ]; then unset STRSIM_B200_LIB; else export STRSIM_B200_LIB=$PWD/exp/variants/lib$v.so; fi
 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e 2>$O/err_C2.log | show C2-$v
 timeout 300 python bench.py --workload L1 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>$O/err_L1.log | show L1-$v
 timeout 300 python bench.py --workload C3 --rows 50000000 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>$O/err_C3.log | show C3-$v
 timeout 300 python bench.py --workload M1 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>$O/err_M1.log | show M1-$v
done
