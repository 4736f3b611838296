#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
echo "== suite"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $O/r02j_gputests.txt
echo "== memcheck (MacaqueV kernels)"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_golden_fixtures.py tests/test_gpu_round2.py -m gpu -q -k "macaque or golden or lane or config1 or tma" 2>&1 | tail -4 | tee $O/r02j_memcheck.txt
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
run() { name=$1; shift; timeout 300 $B "$@" > $O/r02j_$name.json 2> $O/r02j_$name.err; echo "$name rc=$?"; python tools/bench_brief.py $name < $O/r02j_$name.json 2>/dev/null || tail -3 $O/r02j_$name.err; }
run cfg3 --config cfg3
run sine_lossless --kind sine --eb lossless --series 200
run cfg5_lossless --config cfg5 --eb lossless
