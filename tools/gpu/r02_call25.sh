#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline"
run() { name=$1; shift; timeout 300 $B "$@" > $O/r02z_$name.json 2> $O/r02z_$name.err; echo "$name rc=$?"; python tools/bench_brief.py $name < $O/r02z_$name.json 2>/dev/null || tail -3 $O/r02z_$name.err; }
run cfg2_g4
MODELARDB_CUDA_LIB=$PWD/modelardb_rs_b200/libmodelardb_cuda_g8.so run cfg2_g8
MODELARDB_CUDA_LIB=$PWD/modelardb_rs_b200/libmodelardb_cuda_g2.so run cfg2_g2
