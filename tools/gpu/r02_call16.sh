#!/bin/bash
# screened fit engine (mdb_fit_screen.cuh): parity first, then A/B against the exact engine
mkdir -p gpurun_out
O=gpurun_out
echo "== fit engine tests"; timeout 900 python -m pytest tests/test_gpu_fit_engines.py -m gpu -x -q 2>&1 | tail -5
echo "== suite"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
B="python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline"
run() { name=$1; shift; timeout 300 $B "$@" > $O/r02q_$name.json 2> $O/r02q_$name.err; echo "$name rc=$?"; python tools/bench_brief.py $name < $O/r02q_$name.json 2>/dev/null || tail -3 $O/r02q_$name.err; }
run cfg2_auto
run cfg2_exact --fit-engine 3
MODELARDB_CUDA_LIB=$PWD/modelardb_rs_b200/libmodelardb_cuda_s4.so run cfg2_s4
run cfg4_auto --config cfg4
run cfg4_exact --config cfg4 --fit-engine 3
run cfg5_auto --config cfg5
run cfg5_screen --config cfg5 --fit-engine 5
run cfg3_auto --config cfg3
nvcc -O3 -arch=sm_100a -o /tmp/fp64 tools/microbench/fp64.cu 2>&1 | tail -2 && /tmp/fp64 > $O/r02q_fp64.txt 2>&1; tail -8 $O/r02q_fp64.txt
