#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
echo "== suite"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $O/r02i_gputests.txt
echo "== memcheck"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_golden_fixtures.py tests/test_gpu_operators.py -m gpu -q -k "not full_size and not pageable and not medium and not threads and not thousands and not segmentation" 2>&1 | tail -4 | tee $O/r02i_memcheck.txt
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
run() { name=$1; shift; timeout 300 $B "$@" > $O/r02i_$name.json 2> $O/r02i_$name.err; echo "$name rc=$?"; python tools/bench_brief.py $name < $O/r02i_$name.json 2>/dev/null || tail -3 $O/r02i_$name.err; }
run tma
run plain --option grid_plain_stores=1
run cfg5_tma --config cfg5
python - <<'PY'
import json
for n in ("tma","plain","cfg5_tma"):
    d=json.load(open(f"gpurun_out/r02i_{n}.json"))
    k=d["roofline"]["all_kernels_ms_per_step"]
    print(n, {x:round(k[x],3) for x in k if x.startswith("k_grid") or x.startswith("k_agg")}, {a:round(b["frac"],3) for a,b in d["roofline"].get("stages").items()})
PY
