#!/bin/bash
# Round 2, GPU call 1: whole GPU suite on the product build; the MacaqueV run decoder (first time on a GPU): its tests
# and the lossless random-walk bench, next to the product build on the same box; sanitizer over the MacaqueV kernels.
mkdir -p gpurun_out
O=gpurun_out
RUNS=$PWD/modelardb_rs_b200/libmodelardb_cuda_runs.so
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r02a_smi.txt 2>&1
echo "== suite (product build)"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $O/r02a_gputests.txt
echo "== macaque tests, run decoder build"
MODELARDB_CUDA_LIB=$RUNS timeout 600 python -m pytest tests -m gpu -x -q -k "macaque or golden or medium or walk or lane or config1" 2>&1 | tail -8 | tee $O/r02a_gputests_runs.txt
echo "== bench walk lossless: product, then run decoder"
timeout 200 python bench.py --kind walk --eb lossless --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > $O/r02a_walk_base.json 2> $O/r02a_walk_base.err; echo "rc=$?"
MODELARDB_CUDA_LIB=$RUNS timeout 200 python bench.py --kind walk --eb lossless --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > $O/r02a_walk_runs.json 2> $O/r02a_walk_runs.err; echo "rc=$?"
MODELARDB_CUDA_LIB=$RUNS timeout 200 python bench.py --kind sine --eb lossless --series 200 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > $O/r02a_sine_lossless_runs.json 2> $O/r02a_sine_lossless_runs.err; echo "rc=$?"
timeout 200 python bench.py --kind sine --eb lossless --series 200 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > $O/r02a_sine_lossless_base.json 2> $O/r02a_sine_lossless_base.err; echo "rc=$?"
echo "== memcheck over the MacaqueV kernels (warp, lanes with a low switch), product build"
MDBCU_LANE_ROWS_MIN=4 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q -k "macaque or lane or bad_offsets" 2>&1 | tail -12 | tee $O/r02a_memcheck_lanes.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_golden_fixtures.py -m gpu -q -k "macaque or golden or lane" 2>&1 | tail -12 | tee $O/r02a_memcheck_warp.txt
MDBCU_LANE_ROWS_MIN=4 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_golden_fixtures.py tests/test_gpu_parity.py -m gpu -q -k "golden or long_macaque" 2>&1 | tail -12 | tee $O/r02a_racecheck.txt
echo "== memcheck, run decoder build"
MODELARDB_CUDA_LIB=$RUNS timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_golden_fixtures.py -m gpu -q -k "macaque or golden" 2>&1 | tail -12 | tee $O/r02a_memcheck_runs.txt
for f in $O/r02a_walk_base $O/r02a_walk_runs $O/r02a_sine_lossless_base $O/r02a_sine_lossless_runs; do python tools/bench_brief.py $(basename $f) < $f.json 2>/dev/null || tail -3 $f.err; done
