#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline"
run() { name=$1; shift; timeout 300 $B "$@" > $O/r02v_$name.json 2> $O/r02v_$name.err; echo "$name rc=$?"; python tools/bench_brief.py $name < $O/r02v_$name.json 2>/dev/null || tail -3 $O/r02v_$name.err; }
run cfg2_c32 --option chunk_len=32768
run cfg2_c128 --option chunk_len=131072
export MODELARDB_CUDA_LIB=$PWD/modelardb_rs_b200/libmodelardb_cuda_b5.so
run cfg2_b5
export MODELARDB_CUDA_LIB=$PWD/modelardb_rs_b200/libmodelardb_cuda_b6.so
run cfg2_b6
