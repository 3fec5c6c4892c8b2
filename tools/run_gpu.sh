cd $GRAFT_REPO_ROOT
R=r2
K="random_short_mixed or length_boundaries or wide_rows or fused_measures_subsets or scattered_views"
for tool in racecheck memcheck; do
  timeout 500 compute-sanitizer --tool $tool --log-file gpurun_out/${R}_sanitizer_$tool.log \
      python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$K" > gpurun_out/${R}_sanitizer_${tool}_pytest.log 2>&1
  echo "$tool: exit $?"; tail -3 gpurun_out/${R}_sanitizer_${tool}_pytest.log; tail -4 gpurun_out/${R}_sanitizer_$tool.log
done
