#!/bin/bash
# round 2, final evidence for the screened engine: sanitizer, ncu launch list + full capture of the default bench, reference arm
mkdir -p gpurun_out
O=gpurun_out
echo "== memcheck (screened engine: fit by fit, and the whole compress path)"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_fit_engines.py -m gpu -q -x -k "sine or (screened and (mixed or lossy))" > $O/r02x_memcheck.txt 2>&1; echo "rc=$?"; tail -4 $O/r02x_memcheck.txt
echo "== racecheck"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_fit_engines.py -m gpu -q -x -k "sine_series_match_oracle and screened" > $O/r02x_racecheck.txt 2>&1; echo "rc=$?"; tail -4 $O/r02x_racecheck.txt
echo "== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O/r02_launches_bench.log 2>&1; wc -l $O/r02_launches.csv
echo "== reference arm"
timeout 600 python bench.py --impl reference > $O/r02x_reference.json 2> $O/r02x_reference.err; cut -c1-300 $O/r02x_reference.json
