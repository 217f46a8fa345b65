import json
import sys

d = json.loads(sys.stdin.read())
print(round(d["ms_per_step"], 4), d["kernels_us"], "e2e", round(d["e2e"]["ms_per_step"], 3), "seq",
      round(d["e2e"].get("sequential_ms_per_step", 0), 3))
