"""Shared by the CPU and GPU projection tests: the cases of src/fs/tests/CMakeLists.txt:37-42 (fs-ex1 on the two-hex mesh of
ex1.c:53-102) and the parser of what the test prints."""
import os
import re

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refout")

# refout -> (-const_BDeg, -frequency, -proj_qmethod sparse?)
CASES = {
    "fs-mf-projection-v1": (5, (1.0, 1.0, 1.0), False),
    "fs-mf-projection-v3": (5, (1.0, 1.0, 1.0), False),
    "fs-mf-projection-v3-sparse": (5, (1.0, 1.0, 1.0), True),
    "fs-mf-op-proj-0": (9, (5.0, 4.0, 3.0), False),
    "fs-mf-op-proj-1": (6, (2.0, 2.0, 2.0), False),
}


def parse(name):
    """-> dict: expanded, global, sums (|x|,|g|,|y| of useFS), fnorm0 (as printed), pointwise[3], gradient[3] (as printed)"""
    out = {}
    for line in open(os.path.join(GOLDEN, name + ".refout")):
        m = re.match(r"expanded size (\d+), global size (\d+)", line)
        if m:
            out["expanded"], out["global"] = int(m.group(1)), int(m.group(2))
        m = re.match(r"\|x\| = (\S+), \|g\| = (\S+), \|y\| = (\S+)", line)
        if m:
            out["sums"] = tuple(float(m.group(i)) for i in (1, 2, 3))
        m = re.match(r"\s*0 SNES Function norm (\S+)", line)
        if m:
            out["fnorm0"] = m.group(1)
        m = re.match(r"(Pointwise projection|Gradient ptwise proj) residual\s+\|x\|_1\s+(\S+)\s+\|x\|_2\s+(\S+)\s+\|x\|_inf\s+(\S+)", line)
        if m:
            out["pointwise" if m.group(1).startswith("Pointwise") else "gradient"] = tuple(m.group(i) for i in (2, 3, 4))
    return out


def printed_equal(value, printed, fmt):
    """value prints as the golden does, or differs by one unit of the last printed digit"""
    if fmt % value == printed:
        return True
    ref = float(printed)
    digits = len(printed.split("e")[0].replace(".", "").replace("-", "")) if "e" in printed else len(printed.replace(".", "").lstrip("0"))
    ulp = 10.0 ** (np.floor(np.log10(abs(ref))) - (digits - 1))
    return abs(value - ref) <= 1.5 * ulp


def cg(A, b, rtol=1e-14, maxit=5000):
    """plain conjugate gradients on the (SPD) mass operator: the projection solve, converged far below the printed digits"""
    x, r = np.zeros_like(b), b.copy()
    p, rr, b2 = r.copy(), r @ r, b @ b
    for _ in range(maxit):
        Ap = A(p)
        a = rr / (p @ Ap)
        x += a * p
        r -= a * Ap
        rn = r @ r
        if rn <= rtol * rtol * b2:
            break
        p = r + (rn / rr) * p
        rr = rn
    return x
