"""Known answers of the reference's own ellip goldens, reproduced by the oracle restatement driven over the
structured-box space builder.  These pin the full residual path (inhomogeneous Dirichlet data, geometry, FAST rule,
pointwise p-Laplacian, transposes, E scatter) and the embedded-Q1 assembly including reference quirks 6-i / 6-ii
(SURVEY.md section 0, Appendix C).  CPU only."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import oracle_lib as ol
from dohp_b200.box import build_box_space, q1_pattern

# src/fs/tests/CMakeLists.txt:53-54 : ellip -dmesh_gen block (default 4x4x4, mnp=5,5,5) -exact 0 -exact_a 1.6 -exact_b 1.7
#   -exact_c 1.8 -ellip_p 1.6 -const_bdeg {1,3}; goldens src/fs/tests/refout/ellip-e0-b{1,3}-p16.refout lines 1-2
CASES = [(1, 3.17818, 0.939628), (3, 4.93408, 158.134)]


@pytest.fixture(scope="module", autouse=True)
def _build():
    ol.build_oracle()


def _fmt6(x):
    return float("%.6g" % x)


@pytest.mark.parametrize("deg,fnorm,knorm", CASES)
def test_ellip_e0_golden_norms(deg, fnorm, knorm):
    space = build_box_space((4, 4, 4), deg)
    op = ol.OracleEllip(space, p=1.6, exact=0, a=1.6, b=1.7, c=1.8)
    x0 = np.zeros(space.n)
    F = op.function(x0)
    assert _fmt6(np.linalg.norm(F)) == fnorm                     # "0 SNES Function norm"
    rowptr, colidx = q1_pattern(space)
    vals = op.jacobian_q1(x0, rowptr, colidx)
    Jp = sp.csr_matrix((vals, colidx, rowptr), shape=(space.n, space.n))
    z = spla.splu(Jp.tocsc()).solve(F)
    assert _fmt6(np.linalg.norm(z)) == knorm                     # "0 KSP Residual norm" (left-preconditioned, LU)


def test_ellip_morph_golden_norm():
    """src/fs/tests/CMakeLists.txt:57 : -morph 1 -twist 0.4 -const_bdeg 3 (ellip_p=2, a=b=c=1); ellip-morph.refout:1
    '0 KSP Residual norm 111.034' pins quirk 6-ii (element-natural cjinv indexed patch-natural)."""
    space = build_box_space((4, 4, 4), 3, morph=dict(morph=1.0, twist=0.4))
    op = ol.OracleEllip(space, p=2.0, exact=0)
    x0 = np.zeros(space.n)
    F = op.function(x0)
    rowptr, colidx = q1_pattern(space)
    vals = op.jacobian_q1(x0, rowptr, colidx)
    Jp = sp.csr_matrix((vals, colidx, rowptr), shape=(space.n, space.n))
    z = spla.splu(Jp.tocsc()).solve(F)
    assert _fmt6(np.linalg.norm(z)) == 111.034


def test_matrix_free_jacobian_is_derivative_of_residual():
    """J(x) dx (EllipShellMatMult with the stash written by EllipFunction) == directional derivative of F."""
    space = build_box_space((2, 3, 2), 3, morph=dict(morph=0.3, twist=0.2))
    op = ol.OracleEllip(space, p=1.6, lam=0.7, exact=0, a=1.1, b=0.9, c=1.3)
    rng = np.random.default_rng(0)
    x, dx = 0.3 * rng.standard_normal(space.n), rng.standard_normal(space.n)
    h = 1e-6
    Fp, Fm = op.function(x + h * dx), op.function(x - h * dx)
    op.function(x)                                               # leaves the stash at x
    Jdx = op.mult(dx)
    fd = (Fp - Fm) / (2 * h)
    assert np.allclose(Jdx, fd, rtol=0, atol=1e-7 * np.abs(fd).max())


def test_fast_assembled_jacobian_equals_matrix_free():
    """ellip-morph-qfast (src/fs/tests/CMakeLists.txt:58): assembling on the FAST rule goes through
    dEFSGetExplicitSparse_Basic and yields the true Jacobian, so J_assembled*x == MatMult(x)."""
    space = build_box_space((2, 2, 2), 2, morph=dict(morph=0.5, twist=0.3))
    op = ol.OracleEllip(space, p=2.0, exact=0, jac_qmethod=ol.FAST)
    x0 = np.zeros(space.n)
    op.function(x0)
    # dense pattern within an element: every pair of nodes sharing an element
    d3 = (space.p + 1) ** 3
    g = space.l2g[space.e2l.reshape(space.nelem, d3)].astype(np.int64)
    rows, cols = np.repeat(g, d3, axis=1).ravel(), np.tile(g, (1, d3)).ravel()
    keep = (rows >= 0) & (cols >= 0)
    key = np.unique(rows[keep] * space.n + cols[keep])
    r, c = key // space.n, key % space.n
    rowptr = np.concatenate(([0], np.cumsum(np.bincount(r, minlength=space.n)))).astype(np.int32)
    colidx = c.astype(np.int32)
    vals = op.jacobian_q1(x0, rowptr, colidx)
    J = sp.csr_matrix((vals, colidx, rowptr), shape=(space.n, space.n))
    x = np.random.default_rng(1).standard_normal(space.n)
    y = op.mult(x)
    assert np.allclose(J @ x, y, rtol=0, atol=1e-12 * np.abs(y).max())


def test_qfast_golden_ksp_norm():
    """ellip-morph-qfast.refout:1 '0 KSP Residual norm 172.603' = ||LU(Jp_fast)^-1 F(0)|| on the morphed 4^3 box, degree 3,
    ellip_p 1.9, exact 3: pins the dEFSGetExplicitSparse_Basic restatement (dense fallback on the FAST rule)."""
    from dohp_b200.box import dense_pattern
    space = build_box_space((4, 4, 4), 3, morph=dict(morph=0.5, twist=0.1))
    op = ol.OracleEllip(space, p=1.9, exact=3, jac_qmethod=ol.FAST)
    x0 = np.zeros(space.n)
    F = op.function(x0)
    rp, ci = dense_pattern(space)
    Jp = sp.csr_matrix((op.jacobian_q1(x0, rp, ci), ci, rp), shape=(space.n, space.n))
    y = spla.splu(Jp.tocsc()).solve(F)
    assert float("%.6g" % np.linalg.norm(y)) == 172.603
    assert np.linalg.norm(spla.splu(Jp.tocsc()).solve(F - op.mult(y))) < 1e-11 * np.linalg.norm(y)
