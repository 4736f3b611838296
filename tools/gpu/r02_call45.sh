#!/bin/bash
# Swing sums with the closed-form denominator: parity files, a short soak, the bench
mkdir -p gpurun_out
O=gpurun_out
echo "== tests"; timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden_fixtures.py tests/test_gpu_round2.py tests/test_gpu_fit_engines.py -m gpu -x -q 2>&1 | tail -3
timeout 200 python tools/gpu_soak.py 25 7000 > $O/r02_gpu_soak4.json 2> $O/r02_gpu_soak4.err; echo "soak rc=$?"; tail -c 600 $O/r02_gpu_soak4.json
B="python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline"
timeout 200 $B > $O/r02am_cfg2.json 2> $O/r02am_cfg2.err; echo "cfg2 rc=$?"; python tools/bench_brief.py cfg2 < $O/r02am_cfg2.json 2>/dev/null || tail -3 $O/r02am_cfg2.err
