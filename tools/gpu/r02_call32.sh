#!/bin/bash
# round 2: sanitizer over the kernels added late in the round (tile search, grid_range, device sort, long parallel Swing sums)
mkdir -p gpurun_out
O=gpurun_out
echo "== memcheck"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_round2.py tests/test_multivariate_plan.py tests/test_gpu_operators.py -m gpu -q -x -k "grid_range or tile_kernels or device_sort or device_plan or time_predicate or stream" > $O/r02y_memcheck.txt 2>&1; echo "rc=$?"; tail -4 $O/r02y_memcheck.txt
echo "== racecheck"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_round2.py tests/test_multivariate_plan.py -m gpu -q -x -k "(tile_kernels and search and sine) or (device_sort and 5000) or grid_range_on_long" > $O/r02y_racecheck.txt 2>&1; echo "rc=$?"; tail -4 $O/r02y_racecheck.txt
echo "== suite"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
