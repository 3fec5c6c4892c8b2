cd $GRAFT_REPO_ROOT
python - <<'PY'
import sys, time, json, os
sys.path[:0]=['.','polars-strsim_b200']
from bench_support import plugin_driver, workloads
from polars_strsim import _native
M=("levenshtein","jaro","jaro_winkler","jaccard","sorensen_dice")
A,B=workloads.make_pairs(2,10_000_000)
L=_native.lib()
def step():
    plugin_driver.cache_clear()
    t=[]
    for m in M:
        t0=time.perf_counter(); r=plugin_driver.call(m,A,B); t.append((time.perf_counter()-t0)*1e3); r.release()
    return t
for spec in (1,0):
    L.strsim_b200_speculation(spec)
    for i in range(12):
        t=step()
        if i>=8: print('spec',spec,'step',i,'total %.2f'%sum(t),['%.2f'%x for x in t])
    time.sleep(1.0)
PY
