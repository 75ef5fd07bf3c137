"""Registers / spills per kernel from `nvcc -Xptxas -v` output on stdin (development helper)."""
import re
import subprocess
import sys

cur = None
rows = {}
for line in sys.stdin:
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m and cur:
        rows.setdefault(cur, {}).update(stack=int(m.group(1)), sst=int(m.group(2)), sld=int(m.group(3)))
    m = re.search(r"Used (\d+) registers", line)
    if m and cur:
        rows.setdefault(cur, {}).update(regs=int(m.group(1)))
names = subprocess.run(["c++filt"] + list(rows), capture_output=True, text=True).stdout.split("\n")
for n, (k, v) in zip(names, rows.items()):
    n = re.sub(r"\(dohp::\w+<.*", "", n).replace("void dohp::", "").replace("(int)", "").replace("(bool)", "")
    print("%-60s regs %3d  spill st/ld %4d/%4d" % (n, v.get("regs", -1), v.get("sst", 0), v.get("sld", 0)))
