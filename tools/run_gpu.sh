cd $GRAFT_REPO_ROOT
python - <<'PY'
import sys, time, ctypes
sys.path[:0]=['.','polars-strsim_b200']
from bench_support import plugin_driver, workloads
from polars_strsim import _native
M=("levenshtein","jaro","jaro_winkler","jaccard","sorensen_dice")
A,B=workloads.make_pairs(2,10_000_000)
L=_native.lib()
t_start=time.perf_counter()
def step(tag):
    plugin_driver.cache_clear()
    t=[]
    for m in M:
        t0=time.perf_counter(); r=plugin_driver.call(m,A,B); t.append((time.perf_counter()-t0)*1e3); r.release()
    print('%6.2f s' % (time.perf_counter()-t_start), tag,'total %.2f'%sum(t),['%.2f'%x for x in t], flush=True)
for k in range(12):
    step('step %d'%k)
    if k<2: time.sleep(0.6)
PY
