"""Print the few numbers of a bench.py JSON line (stdin) that matter when comparing builds.

Usage: python bench.py ... | python tools/bench_brief.py [label]
"""
import json
import sys

label = sys.argv[1] if len(sys.argv) > 1 else ""
d = json.loads(sys.stdin.read())
k = d["roofline"]["all_kernels_ms_per_step"]
top = sorted(k.items(), key=lambda kv: -kv[1])[:5]
print(label, f"value {d['value'] / 1e9:.2f} G/s", "stages", {s: round(v, 2) for s, v in d["stage_ms_median"].items()},
      "rounds", d.get("compress_chain_rounds"), "top", {n: round(v, 2) for n, v in top},
      "e2e", round(d["e2e"]["value"] / 1e9, 2) if d.get("e2e") else None)
