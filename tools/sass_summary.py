"""Static evidence for the shipped library (no GPU needed): per sm_100a kernel of dohp_b200/libdohp_b200.so the registers, static shared
memory and local (spill) bytes from `cuobjdump --dump-resource-usage`, and the instruction mix from `cuobjdump -sass`: DFMA/DADD/DMUL
(FP64 pipe), LDS/STS (shared-memory pipe), LDG/STG/RED (global), LDGSTS (cp.async), LDL/STL (spills), BAR.  Usage:

    python tools/sass_summary.py [path/to/lib.so] > profiles/<round>_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "dohp_b200", "libdohp_b200.so")
res = subprocess.run(["cuobjdump", "--dump-resource-usage", so], capture_output=True, text=True, check=True).stdout
usage = {}
cur = None
for line in res.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
    if m and cur:
        r_, st_, sh_, lo_ = (int(g) for g in m.groups())
        usage[cur] = (r_, sh_, st_ + lo_)
        cur = None

CLASSES = [("dfma", r"^DFMA"), ("dadd_dmul", r"^(DADD|DMUL)"), ("mufu_dsetp", r"^(MUFU|DSETP)"), ("lds", r"^LDS"), ("sts", r"^STS"), ("ldg", r"^LDG(?!STS)"),
           ("stg", r"^STG"), ("red_atom", r"^(RED|ATOM)"), ("ldgsts", r"^LDGSTS"), ("ldl", r"^LDL"), ("stl", r"^STL"), ("bar", r"^BAR"), ("ldc", r"^(LDC|LDCU|ULDC)")]
mix = collections.defaultdict(collections.Counter)
proc = subprocess.Popen(["cuobjdump", "-sass", so], stdout=subprocess.PIPE, text=True)
cur = None
for line in proc.stdout:
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur:
        op = m.group(1)
        mix[cur]["total"] += 1
        for name, pat in CLASSES:
            if re.match(pat, op):
                mix[cur][name] += 1
                break
proc.wait()

names = list(usage)
dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.split("\n")


def short(n):
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"\((dohp::)?\w+(<[^()]*>)?( const)?[&*]?(, .*)?\)$", "", n)      # drop the parameter list
    n = n.replace("dohp::", "").replace("(int)", "").replace("(bool)", "")
    return n[:70]


cols = ["total"] + [c for c, _ in CLASSES]
print("# %s: %d sm_100a kernels; columns = registers, static shared B, stack + local B per thread (spills / stack arrays), then SASS instruction counts" % (os.path.basename(so), len(names)))
print("%-70s %4s %6s %5s " % ("kernel", "regs", "smem", "local") + " ".join("%9s" % c for c in cols))
rows = sorted(zip(dem, names), key=lambda t: t[0])
tot = collections.Counter()
for d, n in rows:
    r, s, l = usage[n]
    print("%-70s %4d %6d %5d " % (short(d), r, s, l) + " ".join("%9d" % mix[n][c] for c in cols))
    tot.update(mix[n])
print("%-70s %4s %6s %5s " % ("ALL", "", "", "") + " ".join("%9d" % tot[c] for c in cols))
spilled = [short(d) for d, n in rows if usage[n][2] or mix[n]["ldl"] or mix[n]["stl"]]
print("# kernels with local memory (spills or stack arrays): %d%s" % (len(spilled), (": " + "; ".join(spilled)) if spilled else ""))
