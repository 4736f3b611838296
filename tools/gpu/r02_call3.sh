#!/bin/bash
# Round 2, GPU call 3: lanes with warm-up and the worker-local continuation in k_spec_async.
mkdir -p gpurun_out
O=gpurun_out
echo "== suite"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/r02c_gputests.txt
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
run() { name=$1; shift; timeout 300 $B "$@" > $O/r02c_$name.json 2> $O/r02c_$name.err; echo "$name rc=$?"; python tools/bench_brief.py $name < $O/r02c_$name.json 2>/dev/null || tail -3 $O/r02c_$name.err; }
run default
run warp --fit-engine 3
run L4096 --chunk-len 4096
run L16384 --chunk-len 16384
run L32768 --chunk-len 32768
run W2048 --lane-warmup 2048
run W6144 --lane-warmup 6144
run W8192 --lane-warmup 8192
run L16384_W8192 --chunk-len 16384 --lane-warmup 8192
run rel5 --eb rel:5.0
run walk_lossless --kind walk --eb lossless
run cfg5_rel1 --series 100000 --points 10000
run cfg5_lossless --series 100000 --points 10000 --eb lossless
run series4000 --series 4000
