#!/bin/bash
# Round 2, GPU call 5: lanes with repair rounds (chunks of a unit re-synchronised in parallel), then the stitching.
mkdir -p gpurun_out
O=gpurun_out
echo "== suite"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/r02e_gputests.txt
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
run() { name=$1; shift; timeout 300 $B "$@" > $O/r02e_$name.json 2> $O/r02e_$name.err; echo "$name rc=$?"; python tools/bench_brief.py $name < $O/r02e_$name.json 2>/dev/null || tail -3 $O/r02e_$name.err; }
run W0 --lane-warmup 0
run W1024 --lane-warmup 1024
run W4096 --lane-warmup 4096
run warp --fit-engine 3
run W0_L8192 --lane-warmup 0 --chunk-len 8192
run W0_L16384 --lane-warmup 0 --chunk-len 16384
run rel5_W0 --eb rel:5.0 --lane-warmup 0
run walk_lossless_W0 --kind walk --eb lossless --lane-warmup 0
run cfg5_rel1 --series 100000 --points 10000
run cfg5_lossless --series 100000 --points 10000 --eb lossless
run series3000_W0 --series 3000 --lane-warmup 0
