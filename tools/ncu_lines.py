"""Per-source-line view of an ncu capture: joins `ncu -i X.ncu-rep --page source --csv` (SASS rows with samples and
executed-instruction counts) with `nvdisasm -g` line markers of the same cubin (instruction order is identical).
    python tools/ncu_lines.py src.csv file.cubin 'kernel mangled-name substring' [top]"""
import csv
import re
import subprocess
import sys

src, cubin, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(src)))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr) and r[0].startswith("0x")]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
lines, cur, on = [], None, False
for l in dis:
    if l.startswith("//--------------------- .text."):
        on = kname in l
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+", l):
        lines.append(cur)
print("sass rows in profile %d, in cubin %d" % (len(body), len(lines)))
agg = {}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r, ln in zip(body, lines):
    a = agg.setdefault(ln, [0, 0, {}])
    a[0] += int(r[col["# Samples"]] or 0)
    a[1] += int(r[col["Instructions Executed"]] or 0)
    for s in stalls:
        v = int(r[col[s]] or 0)
        if v:
            a[2][s] = a[2].get(s, 0) + v
ts, ti = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    dom = sorted(a[2].items(), key=lambda kv: -kv[1])[:3]
    print("%-22s samples %5.1f%%  instr %5.1f%%  %s" % ("%s:%d" % ln if ln else "?", 100.0 * a[0] / ts, 100.0 * a[1] / ti, " ".join("%s %.0f%%" % (k.replace("stall_", ""), 100.0 * v / max(a[0], 1)) for k, v in dom)))
