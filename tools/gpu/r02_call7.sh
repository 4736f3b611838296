#!/bin/bash
# Round 2, GPU call 7 (2 GPUs): the library's own NCCL communicator: tests, cfg2 / cfg4 bench lines at 1 and 2 GPUs.
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out
echo "== multi-GPU tests"; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -6 | tee $O/r02g_gputests_multi.txt
echo "== suite"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $O/r02g_gputests.txt
run1() { name=$1; shift; timeout 400 python bench.py "$@" > $O/r02g_$name.json 2> $O/r02g_$name.err; echo "$name rc=$?"; python tools/bench_brief.py $name < $O/r02g_$name.json 2>/dev/null || tail -3 $O/r02g_$name.err; }
runN() { name=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N "$@" > $O/r02g_$name.json 2> $O/r02g_$name.err; echo "$name rc=$?"; python tools/bench_brief.py $name < $O/r02g_$name.json 2>/dev/null || tail -3 $O/r02g_$name.err; }
run1 cfg2_1gpu --steps 3 --warmup 3 --no-cpu-baseline
runN cfg2_${N}gpu --steps 3 --warmup 3 --no-cpu-baseline
run1 cfg4_1gpu --config cfg4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e
runN cfg4_${N}gpu --config cfg4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e
