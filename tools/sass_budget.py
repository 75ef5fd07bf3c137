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


# ---- the stash-mode kernels of the other bench entries (same files): static bounds against the measured launch ------------------------------------
def row(pattern):
    for ln in open(os.path.join(ROOT, "profiles", "r3_sass_summary.txt")):
        if ln.startswith(pattern + " "):
            c = [int(v) for v in ln[len(pattern):].split()]       # regs smem local total dfma dadd_dmul mufu lds sts ldg stg red ...
            return dict(regs=c[0], fp=c[4] + c[5], l1=c[7] + c[8] + c[9] + c[10] + c[11])
    raise KeyError(pattern)


oc = bench["other_configs"]
others = [("p-Laplace Jacobian p=4 (stash 40 B/pt)", "ellip_kernel<5, 5, 1, true>", 5, 64 ** 3, oc["ellip_plaplace_p4"]["ms"]),
          ("p-Laplace Jacobian p=8", "ellip_kernel<9, 9, 1, true>", 9, 64 ** 3, oc["ellip_plaplace_p8"]["ms"]),
          ("elast Jacobian 48^3 p=6 (3 components, stash 72 B/pt)", "vec_kernel<7, 6, 7, 0, 1, true>", 7, 48 ** 3, oc["elast_48^3_p6_matmult"]["ms"]),
          ("stokes coupled Jacobian 32^3 p4/p3 (stash 64 B/pt)", "vec_kernel<5, 4, 5, 1, 1, true>", 5, 32 ** 3, oc["stokes_p4p3_32^3_per_gpu_matmult"]["ms"])]
print()
print("# stash-mode kernels: per quadrature point (all components), same conventions")
print("%-58s %4s %8s %8s | %8s %8s %9s | %9s" % ("kernel", "regs", "FP64", "L1 acc", "fp64 clk", "L1 clk", "measured", "bound/meas"))
for label, pat, n, nelem, t in others:
    r = row(pat)
    fp, l1 = r["fp"] / n, r["l1"] / n
    cf, cl = fp / 64, l1 / 16
    clk = t * 1e-3 * 1.965e9 * 148 / (nelem * n ** 3)
    print("%-58s %4d %8.1f %8.1f | %8.2f %8.2f %9.2f | %9.2f" % (label, r["regs"], fp, l1, cf, cl, clk, max(cf, cl) / clk))
