#!/bin/bash
# Host-link ceiling with N GPUs copying at once, then bench.py's e2e leg alone at the same N (no CPU baseline).
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi topo -m > $O/r02_topo_$N.txt 2>&1
lscpu | head -25 > $O/r02_lscpu_$N.txt 2>&1
numactl -H >> $O/r02_lscpu_$N.txt 2>&1
for k in 1 $N; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $k --master-addr 127.0.0.1 --master-port 29541 tools/pcie_probe_multi.py 1.0 > $O/r02_pcie_probe_${k}of$N.json 2> $O/r02_pcie_probe_${k}of$N.err
  cat $O/r02_pcie_probe_${k}of$N.json
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > $O/r02_bench_e2e_$N.json 2> $O/r02_bench_e2e_$N.err
python tools/e2e_brief.py < $O/r02_bench_e2e_$N.json 2>/dev/null || tail -5 $O/r02_bench_e2e_$N.err
