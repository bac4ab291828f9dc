#!/usr/bin/env python
"""One-line summary of a bench.py JSON line (used by tools/gpu_visit.sh)."""
import json
import sys

try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d["roofline"]
    ph = d.get("phases") or {}
    print(d["config"]["workload"][:44], "| %.3e p-steps/s" % d["value"], "| ms/step %.4f" % d["ms_per_step"],
          {k: round(v, 4) for k, v in r["phase_ms"].items()}, "| frac", round(r["frac"], 3), r["kernel"],
          "| e2e %.3e" % d["e2e"]["value"], "| cpu %.3e" % ((d.get("cpu_baseline") or {}).get("value", 0)))
    late = ph.get("late")
    if late:
        print("   late", late["steps"], "%.3e p-steps/s" % late["value"], "ms/step %.4f" % late["ms_per_step"],
              {k: round(v, 4) for k, v in late["phase_ms"].items()}, "= %.2f of early" % (late["value"] / d["value"]))
except Exception as e:  # noqa: BLE001
    print("FAILED", e)
