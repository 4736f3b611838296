#!/bin/bash
# the last check of round 2: the whole GPU suite, smoke, and the lossless legs that changed last
mkdir -p gpurun_out
O=gpurun_out
echo "== suite"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python tools/gpu_soak.py 150 5000 > $O/r02_gpu_soak3.json 2> $O/r02_gpu_soak3.err; echo "soak rc=$?"; cat $O/r02_gpu_soak3.json
