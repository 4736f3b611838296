#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
echo "== tests"; timeout 1200 python -m pytest tests/test_gpu_fit_engines.py tests/test_gpu_parity.py tests/test_multivariate_plan.py -m gpu -x -q -k "long or swing or segmentation or medium or device or multivariate" 2>&1 | tail -3
B="python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline"
run() { name=$1; shift; timeout 300 $B "$@" > $O/r02ab_$name.json 2> $O/r02ab_$name.err; echo "$name rc=$?"; python tools/bench_brief.py $name < $O/r02ab_$name.json 2>/dev/null || tail -3 $O/r02ab_$name.err; }
run cfg2
run cfg4 --config cfg4
run units64k --series 15259 --points 65536
run units64k_notail --series 14208 --points 65536
run units1m_2368 --series 2368 --points 400000
