#!/bin/bash
# scaling line: cfg2 default bench (with e2e) on N GPUs
N=$1
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --no-cpu-baseline > $O/r02_scale_${N}gpu.json 2> $O/r02_scale_${N}gpu.err; echo "rc=$?"; python tools/bench_brief.py scale$N < $O/r02_scale_${N}gpu.json || tail -5 $O/r02_scale_${N}gpu.err
