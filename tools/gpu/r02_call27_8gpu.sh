#!/bin/bash
# Round 2, second 8-GPU call: the bench lines with the screened fit and the tile search kernel (cfg2 with e2e, cfg4 strong
# scaling, cfg5), and the library communicator test.
N=8
mkdir -p gpurun_out
O=gpurun_out
echo "== multi-GPU tests"; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
runN() { name=$1; shift; timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N "$@" > $O/r02q8_$name.json 2> $O/r02q8_$name.err; echo "$name rc=$?"; python tools/bench_brief.py $name < $O/r02q8_$name.json 2>/dev/null || tail -3 $O/r02q8_$name.err; }
runN cfg2_8gpu --steps 3 --warmup 3 --no-cpu-baseline
runN cfg4_8gpu --config cfg4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e
runN cfg5_8gpu --config cfg5 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e
