#!/bin/bash
# A/B of chain-kernel builds (variants/lib_*.so, selected with MODELARDB_CUDA_LIB), then the parity files with the fastest one
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline"
for v in v0 v4 v2 v3 v3b3 v3b5; do
  MODELARDB_CUDA_LIB=$PWD/variants/lib_$v.so timeout 200 $B > $O/r02ak_$v.json 2> $O/r02ak_$v.err; echo "$v rc=$?"
  python tools/bench_brief.py $v < $O/r02ak_$v.json 2>/dev/null || tail -3 $O/r02ak_$v.err
done
best=$(python - <<'PY'
import json
best, t = None, 1e9
for v in ("v4", "v2", "v3", "v3b3", "v3b5"):
    try:
        d = json.load(open(f"gpurun_out/r02ak_{v}.json"))
        c = d["stage_ms_median"]["compress"]
        if c < t: best, t = v, c
    except Exception: pass
print(best or "v3")
PY
)
echo "== tests with $best"
MODELARDB_CUDA_LIB=$PWD/variants/lib_$best.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden_fixtures.py tests/test_gpu_round2.py tests/test_gpu_fit_engines.py -m gpu -x -q 2>&1 | tail -3
