#!/bin/bash
# 100 000 series x 10 000 points at four bounds; one bench line each into gpurun_out/.
mkdir -p gpurun_out
for eb in lossless rel:1.0 rel:5.0 rel:10.0; do
  name=$(echo "$eb" | tr ':.' '__')
  timeout 120 python bench.py --series 100000 --points 10000 --eb "$eb" --steps 3 --warmup 3 --no-e2e --no-cpu-baseline \
    > "gpurun_out/r01h_cfg5_${name}.json" 2> "gpurun_out/r01h_cfg5_${name}.err"
  echo "$eb rc=$?"
done
