"""the reference's golden SOLVER HISTORIES replayed line by line.  The CPU oracle (oracle/oracle.c: residual,
matrix-free Jacobian, embedded-Q1 matrix with the reference's quirks) is driven by the PETSc-faithful Newton / GMRES / CG /
block-ILU(0) of tests/petsc_like.py with the options of src/fs/tests/CMakeLists.txt:53-60; every printed SNES function norm and
every printed KSP residual norm of the refout must be reproduced (same number of Newton steps, same number of Krylov iterations in
every solve, every norm equal to the printed one within one unit of its last printed digit).  CPU only; tests/test_gpu_solver_parity.py
replays the same files with the GPU library as the operator."""
import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import oracle_lib as ol
import petsc_like as pl
from dohp_b200.box import build_box_space, q1_pattern

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refout")

# name -> (degree, ellip_p, lambda, (a,b,c), line search) ; all on the default genblock 4x4x4 box, exact 0
# (src/fs/tests/CMakeLists.txt:53-56; the two -sse cases run PETSc's default cubic backtracking line search)
ELLIP = {
    "ellip-e0-b1-p16": (1, 1.6, 0.0, (1.6, 1.7, 1.8), "basic"),
    "ellip-e0-b3-p16": (3, 1.6, 0.0, (1.6, 1.7, 1.8), "basic"),
    "ellip-e0-b3-p14-sse": (3, 1.4, 0.0, (1.6, 1.7, 1.8), "bt"),
    "ellip-e0-b3-p17-lam1-sse": (3, 1.7, 1.0, (1.0, 1.0, 1.0), "bt"),
}


@pytest.fixture(scope="module", autouse=True)
def _build():
    ol.build_oracle(ref=False)
    ol.oracle().or_set_num_threads(os.cpu_count() or 1)


def check_history(name, fnorms, ksp_hists, last_digit=1):
    snes, ksps = pl.parse_refout(os.path.join(GOLDEN, name + ".refout"))
    if snes:
        assert len(fnorms) == len(snes), ("Newton steps", len(fnorms) - 1, len(snes) - 1, [pl.fmt_short(v) for v in fnorms])
        for i, (v, s) in enumerate(zip(fnorms, snes)):
            assert pl.same_as_printed(v, s, last_digit), ("%d SNES Function norm" % i, pl.fmt_short(v), s)
    assert len(ksp_hists) == len(ksps), ("linear solves", len(ksp_hists), len(ksps))
    for n, (h, g) in enumerate(zip(ksp_hists, ksps)):
        assert len(h) == len(g), ("Krylov iterations of solve %d" % n, len(h) - 1, len(g) - 1, [pl.fmt_short(v) for v in h])
        for i, (v, s) in enumerate(zip(h, g)):
            assert pl.same_as_printed(v, s, last_digit), ("solve %d: %d KSP Residual norm" % (n, i), pl.fmt_short(v), s)


def check_error_lines(name, o, x, bs=1, algebraic=False, function=None):
    """The diagnostics every example prints after its solve (ellip.c:748-757, elast.c:537-546): 'Interpolation residual' = scaled norms of
    x - (exact solution at the nodes), 'Pointwise solution error' / 'Pointwise gradient error' = integral norms of the error at the
    quadrature points (XErrorNorms).  Printed with %8.2e; the solves stop at the same iterate as upstream, so all nine numbers must
    agree to the printed digits (one unit of the last one allowed).  'Algebraic residual' (scaled norms of the last residual) is convergence
    noise after a converged Newton solve and only checked against the clipped value 1.00e-99 there; with -snes_max_it 1 (algebraic=True) it is
    the residual after one step and must agree like the others."""
    import re
    from projection_common import printed_equal
    lines = {}
    for ln in open(os.path.join(GOLDEN, name + ".refout")):
        m = re.match(r"(Algebraic residual|Interpolation residual|Pointwise solution error|Pointwise gradient error)\s+\|x\|_1\s+(\S+)\s+\|x\|_2\s+(\S+)\s+\|x\|_inf\s+(\S+)", ln)
        if m:
            lines[m.group(1)] = tuple(m.group(i) for i in (2, 3, 4))
    assert len(lines) == 4
    en, gn = o.error_norms(x)
    inorm = ol.norms_algebraic_scaled(x - o.nodal_solution(), bs)
    for label, mine in (("Interpolation residual", inorm), ("Pointwise solution error", en), ("Pointwise gradient error", gn)):
        for v, pr in zip(mine, lines[label]):
            assert printed_equal(v, pr, "%8.2e"), (name, label, "%8.2e" % v, pr)
    anorm = ol.norms_algebraic_scaled((function or o.function)(x), bs)
    for v, pr in zip(anorm, lines["Algebraic residual"]):
        if algebraic:
            assert printed_equal(v, pr, "%8.2e"), (name, "Algebraic residual", "%8.2e" % v, pr)
        elif pr == "1.00e-99":
            assert v < 1e-12, (name, "Algebraic residual", v)


def lu_pc(vals, rp, ci, n):
    return spla.splu(sp.csr_matrix((vals, ci, rp), shape=(n, n)).tocsc()).solve


@pytest.mark.parametrize("name", sorted(ELLIP))
def test_ellip_newton_gmres_lu_history(name):
    deg, p, lam, (a, b, c), ls = ELLIP[name]
    space = build_box_space((4, 4, 4), deg)
    o = ol.OracleEllip(space, p=p, lam=lam, exact=0, a=a, b=b, c=c)
    rp, ci = q1_pattern(space)
    x, fnorms, hists = pl.newton(o.function, o.mult, lambda x: lu_pc(o.jacobian_q1(x, rp, ci), rp, ci, space.n), np.zeros(space.n), linesearch=ls)
    check_history(name, fnorms, hists)
    check_error_lines(name, o, x)


def test_ellip_morph_cg_history():
    """ellip-morph: -morph 1 -twist 0.4 -snes_max_it 1 -ksp_type cg -ksp_rtol 1e-8 -pc_type lu (bdeg 3, ellip_p 2): 30 CG iterations."""
    space = build_box_space((4, 4, 4), 3, morph=dict(morph=1.0, twist=0.4))
    o = ol.OracleEllip(space, p=2.0, exact=0)
    rp, ci = q1_pattern(space)
    x0 = np.zeros(space.n)
    F = o.function(x0)
    y, hist = pl.cg(o.mult, lu_pc(o.jacobian_q1(x0, rp, ci), rp, ci, space.n), F, rtol=1e-8)
    check_history("ellip-morph", [], [hist])
    check_error_lines("ellip-morph", o, x0 - y, algebraic=True)      # -snes_max_it 1: the iterate after one full Newton step


def test_ellip_morph_qfast_final_lines():
    """ellip-morph-qfast (-ellip_jac_qmethod fast -ksp_type richardson -pc_type lu -snes_max_it 1): the FAST-rule assembled Jacobian is exact,
    so the step is one LU solve; all twelve printed norms of the iterate after that step (the KSP line itself: tests/test_oracle_ellip.py)."""
    from dohp_b200.box import dense_pattern
    space = build_box_space((4, 4, 4), 3, morph=dict(morph=0.5, twist=0.1))
    o = ol.OracleEllip(space, p=1.9, exact=3, jac_qmethod=ol.FAST)
    x0 = np.zeros(space.n)
    F = o.function(x0)
    rp, ci = dense_pattern(space)
    y = lu_pc(o.jacobian_q1(x0, rp, ci), rp, ci, space.n)(F)
    check_error_lines("ellip-morph-qfast", o, x0 - y, algebraic=True)


def test_elast_newton_gmres_ilu0_history():
    """elast-e0-b3-g01: -pc_type ilu on the BAIJ(3) embedded-Q1 matrix in the natural ordering, GMRES(30) with restarts (up to 66
    iterations per solve), 10 Newton steps."""
    space = build_box_space((4, 4, 4), 3)
    o = ol.OracleElast(space, exact=0, a=0.98, b=1.01, c=1.01)
    rp, ci = q1_pattern(space, bs=3)                        # scalar CSR of the block matrix
    brp, bci = q1_pattern(space, bs=1)                      # block pattern (one block per coupled node pair)
    n = space.n

    def ilu(x):
        A = sp.csr_matrix((o.jacobian_q1(x, rp, ci), ci, rp), shape=(3 * n, 3 * n))
        B = sp.bsr_matrix(A, blocksize=(3, 3))
        B.sort_indices()
        assert np.array_equal(B.indptr, brp) and np.array_equal(B.indices, bci)
        return pl.block_ilu0(brp, bci, B.data, 3)

    x, fnorms, hists = pl.newton(o.function, o.mult, ilu, np.zeros(3 * n), linesearch="bt")
    check_history("elast-e0-b3-g01", fnorms, hists, last_digit=2)
    check_error_lines("elast-e0-b3-g01", o, x, bs=3)
