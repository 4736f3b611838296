#!/bin/bash
# GPU suite, then the lossless cfg5 line and the default line (device-resident legs only).
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 120 python bench.py --series 100000 --points 10000 --eb lossless --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r01k_cfg5_lossless.json 2> gpurun_out/r01k_cfg5_lossless.err; echo "cfg5 rc=$?"
timeout 120 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r01k_default_short.json 2> gpurun_out/r01k_default_short.err; echo "default rc=$?"
timeout 120 python bench.py --kind walk --eb lossless --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r01k_walk_lossless.json 2> gpurun_out/r01k_walk_lossless.err; echo "walk rc=$?"
