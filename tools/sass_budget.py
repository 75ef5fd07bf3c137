"""Static per-quadrature-point instruction budget of the nine affine headline kernels ellip_kernel<N,N,0,true>, from the SASS counts of
profiles/r3_sass_summary.txt (fully unrolled kernels, one thread per grid line of N points) against the measured launch times of
profiles/r3final_bench.json: FP64-pipe instructions (DFMA + DADD/DMUL, 64 per clock and SM) and 64-bit L1 accesses (LDS/STS + LDG/STG/RED,
16 lanes per clock and SM), and the clocks per point both imply, beside the measured clocks per point (148 SMs, 1.965 GHz, 64^3 elements).

    python tools/sass_budget.py > profiles/r3_sass_budget.txt"""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = {}
for ln in open(os.path.join(ROOT, "profiles", "r3_sass_summary.txt")):
    m = re.match(r"ellip_kernel<(\d+), \d+, 0, true>\s+(\d+)\s+(\d+)\s+(\d+)\s+(.*)", ln)
    if m:
        c = [int(v) for v in m.group(5).split()]           # total dfma dadd_dmul mufu lds sts ldg stg red ldgsts ldl stl bar ldc
        rows[int(m.group(1))] = dict(regs=int(m.group(2)), dfma=c[1], dadd=c[2], lds=c[4], sts=c[5], gl=c[6] + c[7] + c[8])
bench = json.load(open(os.path.join(ROOT, "profiles", "r3final_bench.json")))
ms = {e["p"] + 1: e["ms"] for e in bench["roofline"]["per_p"]}
pipe = {e["p"] + 1: e.get("fp64_pipe_pct_ncu") for e in bench["roofline"]["per_p"]}
print("# per quadrature point; clk = clocks per point and SM implied by each pipe alone; measured = launch time * 1.965 GHz * 148 SMs / (64^3 N^3 points)")
print("%2s %4s %9s %9s %8s %8s | %8s %8s %9s | %9s %9s" % ("N", "regs", "DFMA", "DADD/DMUL", "LDS+STS", "LDG/ST/RED", "fp64 clk", "L1 clk", "measured", "bound/meas", "ncu fp64%"))
for n in sorted(ms):
    r = rows[n]
    fp, sh, gl = (r["dfma"] + r["dadd"]) / n, (r["lds"] + r["sts"]) / n, r["gl"] / n
    cf, cl = fp / 64, (sh + gl) / 16
    clk = ms[n] * 1e-3 * 1.965e9 * 148 / (262144 * n ** 3)
    print("%2d %4d %9.1f %9.1f %8.1f %8.1f | %8.2f %8.2f %9.2f | %9.2f %9s" % (n, r["regs"], r["dfma"] / n, r["dadd"] / n, sh, gl, cf, cl, clk, max(cf, cl) / clk,
                                                                            "%.0f" % pipe[n] if pipe[n] else "-"))
