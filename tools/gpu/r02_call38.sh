#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
run() { name=$1; shift; timeout 300 $B "$@" > $O/r02af_$name.json 2> $O/r02af_$name.err; python -c "
import json; d=json.load(open('$O/r02af_$name.json')); e=d['e2e']; print('$name', round(e['value']/1e9,3), e['workers'], e['gate'], e['series_per_gpu_per_step'], e['steps'], {k:round(v,1) for k,v in e['call_ms_mean'].items()})" || tail -3 $O/r02af_$name.err; }
run base
run steps6 --e2e-steps 6
run w6 --e2e-workers 6 --e2e-steps 4
run w6g3 --e2e-workers 6 --e2e-up-gate 3 --e2e-steps 4
run w4g3 --e2e-up-gate 3 --e2e-steps 6
run s400 --e2e-series 400 --e2e-steps 4
run s100w6 --e2e-series 100 --e2e-workers 6 --e2e-steps 8
run w4d2 --e2e-down-gate 2 --e2e-steps 6
