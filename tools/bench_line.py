"""Compact view of bench.py JSON lines read from stdin."""
import json
import sys

for ln in sys.stdin:
    ln = ln.strip()
    if not ln.startswith("{"):
        continue
    l = json.loads(ln)
    r = l.get("roofline") or {}
    print(l["config"]["workload"], "| ms/step", round(l["ms_per_step"], 3), "| q/s", int(l["value"]), "| e2e q/s",
          int(l["e2e"]["value"]), "| scan_ms", round(r.get("kernel_ms", 0), 3), "| TF", round(r.get("achieved", 0), 1),
          "| frac", round(r.get("frac", 0), 3), "| launches", l.get("gpu_launches"), "| uncertified/step", l.get("uncertified_queries_per_step"))
