"""prints the headline numbers of a bench.py JSON line (development tool)"""
import json
import sys

d = json.load(open(sys.argv[1]))
r = d["roofline"]
print("ms_per_step %.3f value %.4g prepass %.3f cells %.3f" % (d["ms_per_step"], d["value"], r["prepass_kernel_ms"], r["accum_kernel_ms"]),
      "e2e %.2f" % d["e2e"]["ms_per_step"] if "e2e" in d else "")
