#!/bin/bash
# round 2: the evidence that goes into profiles/ (one GPU)
mkdir -p gpurun_out
O=gpurun_out
echo "== suite"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== default line"; timeout 900 python bench.py > $O/r02_final_default.json 2> $O/r02_final_default.err; python tools/bench_brief.py default < $O/r02_final_default.json
echo "== reference arm"; timeout 600 python bench.py --impl reference > $O/r02_final_reference.json 2> $O/r02_final_reference.err; cut -c1-200 $O/r02_final_reference.json
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
run() { name=$1; shift; timeout 600 $B "$@" > $O/r02_final_$name.json 2> $O/r02_final_$name.err; echo "$name rc=$?"; python tools/bench_brief.py $name < $O/r02_final_$name.json 2>/dev/null || tail -3 $O/r02_final_$name.err; }
run cfg3 --config cfg3 --no-e2e
run cfg4 --config cfg4 --no-e2e
run cfg5 --config cfg5 --no-e2e
echo "== ncu"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_grid_tile_search -c 1 -o $O/r02_grid_tile_search -f python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $O/ncu_tile_search.log 2>&1; tail -1 $O/ncu_tile_search.log | cut -c1-120
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O/r02_launches_bench.log 2>&1; wc -l $O/r02_launches.csv
