"""Print the e2e block of a bench.py JSON line (stdin). Usage: python bench.py ... | python tools/e2e_brief.py [label]"""
import json
import sys

d = json.loads(sys.stdin.read())
e = d["e2e"]
print(sys.argv[1] if len(sys.argv) > 1 else "", "workers", e["workers"], "gate", e.get("gate"), f"{e['value'] / 1e9:.2f} G pts/s",
      {k: round(v) for k, v in e["call_ms_mean"].items()})
