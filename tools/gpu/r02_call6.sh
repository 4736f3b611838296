#!/bin/bash
# Round 2, GPU call 6: lanes + repair rounds by the cooperative engine; homogeneous units for comparison.
mkdir -p gpurun_out
O=gpurun_out
echo "== suite"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/r02f_gputests.txt
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
run() { name=$1; shift; timeout 300 $B "$@" > $O/r02f_$name.json 2> $O/r02f_$name.err; echo "$name rc=$?"; python tools/bench_brief.py $name < $O/r02f_$name.json 2>/dev/null || tail -3 $O/r02f_$name.err; }
run W4096
run W0 --lane-warmup 0
run W2048 --lane-warmup 2048
run warp --fit-engine 3
run homog_lanes --sine 100:100,5:15,1000:1000
run homog_warp --sine 100:100,5:15,1000:1000 --fit-engine 3
run mid_lanes --sine 80:120,5:15,800:1200
run mid_warp --sine 80:120,5:15,800:1200 --fit-engine 3
run rel5 --eb rel:5.0
run cfg5_rel1 --series 100000 --points 10000
