"""Summarise an ncu source-page CSV (ncu -i X.ncu-rep --page source --csv --kernel-id :::K): stall-reason totals,
shared-memory wavefront excess, and the hottest SASS instructions.  Usage: python profiles/ncu_stalls.py src.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 15
print(rows[0][1])
hdr = rows[1]
body = [r for r in rows[2:] if len(r) == len(hdr) and r[0].startswith("0x")]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {s: sum(int(r[col[s]] or 0) for r in body) for s in stalls}
allsamp = sum(tot.values())
print("stall samples (all):", allsamp)
for s, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v:
        print("  %-24s %8d  %5.1f%%" % (s, v, 100.0 * v / allsamp))
sw, si = sum(int(r[col["L1 Wavefronts Shared"]] or 0) for r in body), sum(int(r[col["L1 Wavefronts Shared Ideal"]] or 0) for r in body)
print("shared wavefronts %d ideal %d (x%.2f)" % (sw, si, sw / max(si, 1)))
ops = {}
for r in body:
    op = r[col["Source"]].split()[0] if r[col["Source"]].split() else "?"
    if op.startswith("@"):
        op = r[col["Source"]].split()[1]
    op = op.split(".")[0]
    ops[op] = ops.get(op, 0) + int(r[col["Instructions Executed"]] or 0)
tot_i = sum(ops.values())
print("instructions executed by opcode:", ", ".join("%s %.1f%%" % (k, 100.0 * v / tot_i) for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:14]))
print("hottest instructions by samples:")
for r in sorted(body, key=lambda r: -int(r[col["# Samples"]] or 0))[:top]:
    dom = max(stalls, key=lambda s: int(r[col[s]] or 0))
    print("  %6s  %-60s %s  shared wf %s/%s" % (r[col["# Samples"]], r[col["Source"]].strip()[:60], dom, r[col["L1 Wavefronts Shared"]], r[col["L1 Wavefronts Shared Ideal"]]))
