#!/usr/bin/env python
"""stdin: PROBE lines of scripts/mgpu_probe.py -> one compact line per rank"""
import json
import sys

for line in sys.stdin:
    d = json.loads(line.split("PROBE ", 1)[1])
    f = lambda k: "%s med %.0f max %.0f cta0 %.0f" % (k, d[k]["median"], d[k]["max"], d[k]["cta0"]) if k in d else ""
    print("rank", d["rank"], "ms/step %.2f cg %.2f us/iter %.2f |" % (d["ms_per_step"], d["cg_ms"], d["us_per_iter"]), f("phase1"), "|", f("phase2"), "|", f("reductions"),
          "| stages", d.get("stage_ms_per_step"))
