#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
SECONDS=0; timeout 900 python bench.py > $O/r02_final_default.json 2> $O/r02_final_default.err; python tools/bench_brief.py default < $O/r02_final_default.json; echo "wall ${SECONDS}s"; tail -3 $O/r02_final_default.err
