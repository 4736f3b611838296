#!/bin/bash
# Round 2, GPU call 4: lanes with in-lane error sums, one wave of chunks, wide steps in the stitching engine.
mkdir -p gpurun_out
O=gpurun_out
echo "== suite"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/r02d_gputests.txt
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
run() { name=$1; shift; timeout 300 $B "$@" > $O/r02d_$name.json 2> $O/r02d_$name.err; echo "$name rc=$?"; python tools/bench_brief.py $name < $O/r02d_$name.json 2>/dev/null || tail -3 $O/r02d_$name.err; }
run default
MODELARDB_CUDA_LIB=$PWD/modelardb_rs_b200/libmodelardb_cuda_mb6.so run mb6
MODELARDB_CUDA_LIB=$PWD/modelardb_rs_b200/libmodelardb_cuda_mb4.so run mb4
run warp --fit-engine 3
run W2048 --lane-warmup 2048
run W3072 --lane-warmup 3072
run W6144 --lane-warmup 6144
run rel5 --eb rel:5.0
run rel5_warp --eb rel:5.0 --fit-engine 3
run walk_lossless --kind walk --eb lossless
run cfg5_rel1 --series 100000 --points 10000
run cfg5_lossless --series 100000 --points 10000 --eb lossless
run series3000 --series 3000
