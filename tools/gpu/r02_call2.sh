#!/bin/bash
# Round 2, GPU call 2: first run of the one-lane-per-chain engine: GPU suite (it is the default engine), then the default
# bench line with it and with the warp engine alone (engine 3), chunk lengths, and the 5 % / lossless / cfg5 workloads.
mkdir -p gpurun_out
O=gpurun_out
echo "== suite"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $O/r02b_gputests.txt
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
run() { name=$1; shift; timeout 300 $B "$@" > $O/r02b_$name.json 2> $O/r02b_$name.err; echo "$name rc=$?"; python tools/bench_brief.py $name < $O/r02b_$name.json 2>/dev/null || tail -3 $O/r02b_$name.err; }
run default_lanes
run default_warp --fit-engine 3
run lanes_chunk2048 --chunk-len 2048
run lanes_chunk8192 --chunk-len 8192
run lanes_chunk16384 --chunk-len 16384
run rel5_lanes --eb rel:5.0
run rel5_warp --eb rel:5.0 --fit-engine 3
run walk_lossless_lanes --kind walk --eb lossless
run walk_lossless_warp --kind walk --eb lossless --fit-engine 3
run cfg5_rel1_lanes --series 100000 --points 10000
run cfg5_rel1_warp --series 100000 --points 10000 --fit-engine 3
