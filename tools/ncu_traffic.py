"""DRAM traffic of the headline launches from an `ncu --set full ... --page raw --csv` export of tools/prof_kernels.py lin2 .. lin10:
   python tools/ncu_traffic.py raw.csv "<source note>" > profiles/rNN_traffic.json      (read by bench.py for roofline.traffic)"""
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
SC = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1}


def val(r, m):
    return float(r[col[m]].replace(",", "")) * SC.get(units[col[m]], 1)


out = {}
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    m = re.search(r"ellip_kernel<(\d+), (\d+), 0", r[col["Kernel Name"]])
    if not m:
        continue
    p = int(m.group(1)) - 1
    out[str(p)] = {"dram_read_gb": round(val(r, "dram__bytes_read.sum") / 1e9, 6), "dram_write_gb": round(val(r, "dram__bytes_write.sum") / 1e9, 6),
                   "ms_under_ncu": round(val(r, "gpu__time_duration.sum"), 6), "regs": int(float(r[col["launch__registers_per_thread"]])),
                   "fp64_pipe_pct": round(float(r[col["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]]), 3),
                   "warps_active_pct": round(float(r[col["sm__warps_active.avg.pct_of_peak_sustained_active"]]), 3),
                   "l1tex_pct": round(float(r[col["l1tex__throughput.avg.pct_of_peak_sustained_active"]]), 3)}
print(json.dumps({"source": sys.argv[2] if len(sys.argv) > 2 else sys.argv[1], "per_degree": out}, indent=1))
