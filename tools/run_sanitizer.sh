#!/bin/bash
# compute-sanitizer passes over the parity tests that exercise every kernel family:
#   gpurun --timeout 1500 -- 'bash tools/run_sanitizer.sh r2'
R=${1:-r2}
cd ${GRAFT_REPO_ROOT:-.}
K="random_short_mixed or long_levenshtein_multiword or long_rows_one_pair or dictionary_encoded or length_boundaries or nulls_slices or wide_rows or fused_measures_subsets or scattered_views"
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --log-file gpurun_out/${R}_sanitizer_$tool.log \
      python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$K" > gpurun_out/${R}_sanitizer_${tool}_pytest.log 2>&1
  echo "$tool: exit $?"; tail -3 gpurun_out/${R}_sanitizer_${tool}_pytest.log; tail -5 gpurun_out/${R}_sanitizer_$tool.log
done
