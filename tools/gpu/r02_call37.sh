#!/bin/bash
# round 2, final: the whole GPU suite, the default line and the reference arm, the other configs, the launch list, memcheck of the late-irregular path
mkdir -p gpurun_out
O=gpurun_out
echo "== suite"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== default line"; timeout 900 python bench.py > $O/r02_final_default.json 2> $O/r02_final_default.err; python tools/bench_brief.py default < $O/r02_final_default.json
echo "== reference arm"; timeout 600 python bench.py --impl reference > $O/r02_final_reference.json 2> $O/r02_final_reference.err; cut -c1-200 $O/r02_final_reference.json
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
run() { name=$1; shift; timeout 600 $B "$@" > $O/r02_final_$name.json 2> $O/r02_final_$name.err; echo "$name rc=$?"; python tools/bench_brief.py $name < $O/r02_final_$name.json 2>/dev/null || tail -3 $O/r02_final_$name.err; }
run cfg3 --config cfg3
run cfg4 --config cfg4
run cfg5 --config cfg5
echo "== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O/r02_launches_bench.log 2>&1; wc -l $O/r02_launches.csv
echo "== memcheck"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "looks_regular" > $O/r02z_memcheck.txt 2>&1; echo "rc=$?"; tail -3 $O/r02z_memcheck.txt
