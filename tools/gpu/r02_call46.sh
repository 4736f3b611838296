#!/bin/bash
# quad loads of k_swing_finish with an L2 prefetch size of 128 B (product) and 256 B (variant), then the parity files
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline"
timeout 60 $B > $O/r02an_l2_128.json 2> $O/r02an_l2_128.err; echo "128 rc=$?"; python tools/bench_brief.py l2_128 < $O/r02an_l2_128.json 2>/dev/null || tail -3 $O/r02an_l2_128.err
MODELARDB_CUDA_LIB=$PWD/variants/lib_l2_256.so timeout 60 $B > $O/r02an_l2_256.json 2> $O/r02an_l2_256.err; echo "256 rc=$?"; python tools/bench_brief.py l2_256 < $O/r02an_l2_256.json 2>/dev/null || tail -3 $O/r02an_l2_256.err
echo "== tests"; timeout 100 python -m pytest tests/test_gpu_parity.py tests/test_golden_fixtures.py tests/test_gpu_round2.py tests/test_gpu_fit_engines.py -m gpu -x -q 2>&1 | tail -3
