#!/bin/bash
# full bench line (no CPU arm) -> $1, and a short summary of the kernel times
python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null > $1
python - $1 <<PY
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("count", round(d["ms_per_step"], 4), round(d["roofline"]["frac"], 4), "e2e_ms", round(d["e2e"]["ms_per_step"], 2))
for k, v in d["extras"].items():
    print(k, round(v["ms_per_step"], 4), round(v.get("roofline_frac", 0), 4))
PY
