"""Reads bench.py's JSON line on stdin and prints the numbers one looks at while iterating."""
import json
import sys

d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print("ms/iter", round(d["ms_per_iter"], 3), "| e2e ms/iter", round(d["e2e"]["ms_per_iter"], 3), "| MPix/s", round(d["value"], 1),
      "| step max", round(d.get("ms_per_step_max") or 0, 2), "| vs ref", round(d.get("ref_cuda", {}).get("speedup_device_resident", 0), 2))
print({k: round(v, 3) for k, v in d.get("stages_ms", {}).items()})
