#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
run() { name=$1; shift; timeout 300 $B "$@" > $O/r02ag_$name.json 2> $O/r02ag_$name.err; python -c "
import json; d=json.load(open('$O/r02ag_$name.json')); e=d['e2e']; print('$name', round(e['value']/1e9,3), e['workers'], e['gate'], e['series_per_gpu_per_step'], e['steps'])" || tail -3 $O/r02ag_$name.err; }
for i in 1 2; do
run base$i
run new$i --e2e-series 100 --e2e-workers 6 --e2e-steps 6
run new8w$i --e2e-series 100 --e2e-workers 8 --e2e-steps 5
done
