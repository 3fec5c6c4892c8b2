set -x
cd $GRAFT_REPO_ROOT
for t in 8 16; do
STRSIM_B200_COPY_THREADS=$t STRSIM_B200_TRACE=1 python tools/plugin_e2e.py 10000000 --pageable > gpurun_out/r2b_trace_t$t.json 2> gpurun_out/r2b_trace_t$t.log
done
tail -c 400 gpurun_out/r2b_trace_t8.json; tail -c 400 gpurun_out/r2b_trace_t16.json
tail -60 gpurun_out/r2b_trace_t8.log
python exp/bw2.py > gpurun_out/r2b_bw2.txt 2>&1; cat gpurun_out/r2b_bw2.txt
