#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --steps 5 --warmup 2 --no-e2e --no-cpu-baseline"
run() { name=$1; shift; timeout 300 $B "$@" > $O/r02ad_$name.json 2> $O/r02ad_$name.err; echo "$name rc=$?"; python tools/bench_brief.py $name < $O/r02ad_$name.json 2>/dev/null || tail -3 $O/r02ad_$name.err; python -c "
import json; d=json.load(open('$O/r02ad_$name.json')); print({k:round(v,3) for k,v in d['roofline']['all_kernels_ms_per_step'].items() if 'agg' in k})"; }
run a2048
MODELARDB_CUDA_LIB=$PWD/modelardb_rs_b200/libmodelardb_cuda_a512.so run a512
MODELARDB_CUDA_LIB=$PWD/modelardb_rs_b200/libmodelardb_cuda_a256.so run a256
