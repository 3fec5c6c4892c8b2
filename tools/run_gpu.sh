cd $GRAFT_REPO_ROOT
O=gpurun_out; mkdir -p $O
show() { python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$1', round(d['ms_per_step'],4), [round(v['ms'],3) for v in d.get('per_measure',{}).values()], d.get('gpu_launches'), d['checksum'], (d.get('long_levenshtein') or {}).get('gcups',''))
"; }
for v in m128_2 m128_1 m64_2 m64_3; do
 if [ $v = main ]; then unset STRSIM_B200_LIB; else export STRSIM_B200_LIB=$PWD/exp/variants/lib$v.so; fi
 timeout 300 python bench.py --workload M1 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>$O/err_M1.log | show M1-$v
done
