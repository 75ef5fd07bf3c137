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


def test_ellip_morph_cg_history():
    """ellip-morph: -morph 1 -twist 0.4 -snes_max_it 1 -ksp_type cg -ksp_rtol 1e-8 -pc_type lu (bdeg 3, ellip_p 2): 30 CG iterations."""
    space = build_box_space((4, 4, 4), 3, morph=dict(morph=1.0, twist=0.4))
    o = ol.OracleEllip(space, p=2.0, exact=0)
    rp, ci = q1_pattern(space)
    x0 = np.zeros(space.n)
    F = o.function(x0)
    _, hist = pl.cg(o.mult, lu_pc(o.jacobian_q1(x0, rp, ci), rp, ci, space.n), F, rtol=1e-8)
    check_history("ellip-morph", [], [hist])


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
