#!/bin/bash
# Round 2, 8-GPU call: host-link ceiling with 1 / 4 / 8 GPUs copying at once, the library communicator test on 4 GPUs,
# and the bench lines of cfg2 (with e2e), cfg4 (strong scaling), cfg3 and cfg5 on 8 GPUs.
N=8
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi topo -m > $O/r02_topo_8.txt 2>&1
lscpu | head -20 > $O/r02_lscpu_8.txt 2>&1
free -g >> $O/r02_lscpu_8.txt 2>&1
for k in 1 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $k --master-addr 127.0.0.1 --master-port 29541 tools/pcie_probe_multi.py 1.0 > $O/r02_pcie_probe_${k}of8.json 2> $O/r02_pcie_probe_${k}of8.err
  cut -c1-330 $O/r02_pcie_probe_${k}of8.json
done
echo "== multi-GPU tests"; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
runN() { name=$1; shift; timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N "$@" > $O/r02p_$name.json 2> $O/r02p_$name.err; echo "$name rc=$?"; python tools/bench_brief.py $name < $O/r02p_$name.json 2>/dev/null || tail -3 $O/r02p_$name.err; }
runN cfg2_8gpu --steps 3 --warmup 3 --no-cpu-baseline
runN cfg4_8gpu --config cfg4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e
runN cfg3_8gpu --config cfg3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e
runN cfg5_8gpu --config cfg5 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e
