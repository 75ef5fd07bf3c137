"""Print the per-degree table of a bench.py JSON line read from stdin."""
import json, sys
tag = sys.argv[1] if len(sys.argv) > 1 else ""
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    j = json.loads(line)
    r = j["roofline"]
    print("%-10s value %.2f GDoF/s  frac %.3f  e2e %.2f | " % (tag, j["value"], r["frac"], j["e2e"]["value"]) +
          " ".join("p%d:%.1f(%.2f)" % (q["p"], q["gdofs"], q["fp64_frac"]) for q in r["per_p"]))
