#!/usr/bin/env python
"""Prints the few numbers of a bench.py JSON line that matter when reading a gpurun log."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d.get("roofline", {})
print("ms/step %.3f  value %.3g %s  e2e %.3g (%.3f ms)  launches %s" % (d["ms_per_step"], d["value"], d["unit"], d["e2e"]["value"], d["e2e"].get("ms_per_step", 0), d.get("gpu_launches")))
for k in r.get("kernels", []):
    print("   %-22s x%-4d avg %.4f ms  %.0f GB/s  frac %.3f  share %.3f" % (k["kernel"], k["launches"], k["avg_launch_ms"], k["achieved"], k["frac"], k["share_of_step"]))
print("   phases", d.get("phases_ms_last_step"))
