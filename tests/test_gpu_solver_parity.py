"""BASELINE.json configs[0]: ellip on an 8x8x8 box, p=4, manufactured forcing, Newton-Krylov with the embedded-Q1 matrix
(LU) as preconditioner — the single-CPU-process reference case.  The same solver (tests/krylov.py) is driven by the CPU
oracle and by the GPU path through the C ABI: Newton and GMRES iteration counts must be IDENTICAL and the residual
histories must agree to the digits the reference's '%g' monitors print (north_star: 'SNES/KSP iteration counts must be
identical')."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import oracle_lib as ol
from dohp_b200 import lib
from dohp_b200.box import q1_pattern
from krylov import gmres, newton

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import torch
    ol.build_oracle(ref=False)
    ol.oracle().or_set_num_threads(8)
    c = lib.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    yield c
    c.close()


@pytest.mark.parametrize("ellip_p,lam,ne,deg", [(2.0, 0.0, (8, 8, 8), 4), (1.6, 0.5, (6, 6, 6), 3)])
def test_newton_krylov_iteration_counts_identical(ctx, ellip_p, lam, ne, deg):
    mesh = lib.BoxMesh(ne, deg)
    o = ol.OracleEllip(mesh, p=ellip_p, lam=lam, exact=0, a=1.0, b=1.0, c=1.0)
    rule = lib.Rule.preferred(ctx, deg)
    efs = lib.EFS(ctx, rule, deg)
    space = lib.Space.from_mesh(ctx, mesh)
    op = lib.Op(ctx, lib.OP_ELLIP, lib.EllipParam(ellip_p, 1.0, lam), rule, [space], [efs], mesh.coords)
    op.set_forcing(o.forcing(op.quadrature_points()))
    op.set_dirichlet(0, o.dirichlet_cache())
    rp, ci = q1_pattern(mesh)
    grp, gci = op.q1_pattern(mesh.l2g)
    assert np.array_equal(rp, grp) and np.array_equal(ci, gci)

    def lu_pc(vals):
        lu = spla.splu(sp.csr_matrix((vals, ci, rp), shape=(mesh.n, mesh.n)).tocsc())
        return lu.solve

    x0 = np.zeros(mesh.n)
    xo, fo, ko = newton(o.function, o.mult, lambda x: lu_pc(o.jacobian_q1(x, rp, ci)), x0)
    xg, fg, kg = newton(op.function, op.mult, lambda x: lu_pc(op.assemble_q1(x, mesh.l2g, None, None)), x0)
    assert len(fo) == len(fg) and ko == kg, (fo, fg, ko, kg)               # identical Newton and Krylov iteration counts
    assert ["%.6g" % v for v in fo[:2]] == ["%.6g" % v for v in fg[:2]]     # what -snes_monitor_short would print
    assert np.allclose(fo, fg, rtol=1e-6, atol=1e-9 * fo[0])
    assert np.abs(xo - xg).max() <= 1e-9 * np.abs(xo).max()
    # the discrete solution converges to the manufactured one (ellip.c:748-758 error norms)
    if lam == 0:          # the manufactured forcing is -div(eta grad u) only (ellip.c:42-49), so u is the solution only without the Bratu term
        u_exact = o.exact_solution(mesh.node_xyz[:mesh.n])[0]
        assert np.abs(xg - u_exact).max() <= 1e-4 * np.abs(u_exact).max()


@pytest.mark.parametrize("deg,ne,scale", [(3, (3, 3, 3), 0.05), (2, (4, 4, 4), 0.1)])
def test_elast_newton_krylov_iteration_counts_identical(ctx, deg, ne, scale):
    """BASELINE.json configs[2] in small: St-Venant-Kirchhoff elasticity (elast.c), matrix-free Jacobian, GMRES preconditioned
    with LU of the embedded-Q1 BAIJ-3 matrix (elast.c:343-407); oracle-driven and GPU-driven solves must take the same
    Newton and Krylov iterations."""
    mesh = lib.BoxMesh(ne, deg)
    o = ol.OracleElast(mesh, exact=0, a=0.6, b=0.5, c=0.4, scale=scale)
    rule = lib.Rule.preferred(ctx, deg)
    efs = lib.EFS(ctx, rule, deg)
    space = lib.Space.from_mesh(ctx, mesh, bs=3)
    op = lib.Op(ctx, lib.OP_ELAST, lib.ElastParam(o.lam, o.mu), rule, [space], [efs], mesh.coords)
    op.set_forcing(o.forcing(op.quadrature_points()))
    op.set_dirichlet(0, o.dirichlet_cache())
    rp, ci = q1_pattern(mesh, bs=3)
    brp, bci = op.q1_pattern(mesh.l2g)
    N = mesh.n * 3

    def lu_csr(vals):
        return spla.splu(sp.csr_matrix((vals, ci, rp), shape=(N, N)).tocsc()).solve

    def lu_bsr(vals):
        return spla.splu(sp.bsr_matrix((vals.reshape(-1, 3, 3), bci, brp), shape=(N, N)).tocsc()).solve

    x0 = np.zeros(N)
    xo, fo, ko = newton(o.function, o.mult, lambda x: lu_csr(o.jacobian_q1(x, rp, ci)), x0)
    xg, fg, kg = newton(op.function, op.mult, lambda x: lu_bsr(op.assemble_q1(x, mesh.l2g, None, None)), x0)
    assert len(fo) == len(fg) and ko == kg, (fo, fg, ko, kg)
    assert len(fo) >= 5 and fo[-1] <= 1e-10 * fo[0]
    assert ["%.6g" % v for v in fo[:2]] == ["%.6g" % v for v in fg[:2]]
    assert np.allclose(fo, fg, rtol=1e-6, atol=1e-9 * fo[0])
    assert np.abs(xo - xg).max() <= 1e-9 * np.abs(xo).max()


def test_stokes_gmres_history_identical(ctx):
    """configs[3] in small: the coupled matrix-free Stokes Jacobian (MatMult_Nest_StokesCoupled, stokes.c:541-577, velocity p=3,
    pressure p=2, power-law rheology) inside unpreconditioned GMRES(30): 40 iterations driven by the oracle and by the GPU
    path must give the same residual history (the saddle-point system is not solved to tolerance without the
    preconditioner; the histories are what -ksp_monitor would print)."""
    mu_ = lib.BoxMesh((3, 3, 3), 3)
    mp_ = lib.BoxMesh((3, 3, 3), 2, dirichlet=0)
    prm = dict(B=1.0, eps=0.1, p=1.2)
    o = ol.OracleStokes(mu_, mp_, exact=1, **prm)
    rule = lib.Rule.preferred(ctx, 3)
    efsu, efsp = lib.EFS(ctx, rule, 3), lib.EFS(ctx, rule, 2)
    su, sp_ = lib.Space.from_mesh(ctx, mu_, bs=3), lib.Space.from_mesh(ctx, mp_, bs=1)
    op = lib.Op(ctx, lib.OP_STOKES, lib.StokesParam(prm["B"], prm["eps"], prm["p"], 0.0), rule, [su, sp_], [efsu, efsp], mu_.coords)
    op.set_forcing(o.forcing(op.quadrature_points()))
    op.set_dirichlet(0, o.dirichlet_cache())
    nu, npr = mu_.n * 3, mp_.n
    x0u, x0p = np.zeros(nu), np.zeros(npr)
    Fo = np.concatenate(o.function(x0u, x0p))
    Fg = np.concatenate(op.function([x0u, x0p]))
    assert np.abs(Fo - Fg).max() <= 1e-12 * np.abs(Fo).max()
    Ao = lambda x: np.concatenate(o.mult(x[:nu], x[nu:]))
    Ag = lambda x: np.concatenate(op.mult([np.ascontiguousarray(x[:nu]), np.ascontiguousarray(x[nu:])]))
    ident = lambda r: r
    _, ho = gmres(Ao, ident, Fo, rtol=1e-8, maxit=40)
    _, hg = gmres(Ag, ident, Fo, rtol=1e-8, maxit=40)
    assert len(ho) == len(hg)
    rel = np.abs(np.array(ho) - np.array(hg)) / np.array(ho)
    print("stokes gmres history rel diff", ["%.1e" % v for v in rel], "history", ["%.3g" % v for v in ho])
    assert rel[:25].max() <= 1e-6              # rounding differences of the operator (1e-15) grow with the Krylov dimension
    assert ["%.5g" % v for v in ho[:10]] == ["%.5g" % v for v in hg[:10]]
    op.close()


# ---- the reference's golden solver histories, replayed with the GPU library as the operator -----------------------------
import petsc_like as pl
from test_refout_replay import ELLIP, check_error_lines, check_history, lu_pc


@pytest.mark.parametrize("name", sorted(ELLIP))
def test_refout_ellip_history_gpu(ctx, name):
    """Every '%d SNES Function norm' and '%d KSP Residual norm' line of src/fs/tests/refout/<name>.refout reproduced by PETSc-faithful
    Newton / GMRES(30) / LU (tests/petsc_like.py) around dB200OpFunction / dB200OpMult / dB200OpAssembleQ1."""
    deg, p, lam, (a, b, c), ls = ELLIP[name]
    mesh = lib.BoxMesh((4, 4, 4), deg)
    o = ol.OracleEllip(mesh, p=p, lam=lam, exact=0, a=a, b=b, c=c)           # forcing and Dirichlet data only (user callbacks in the reference)
    rule = lib.Rule.preferred(ctx, deg)
    efs = lib.EFS(ctx, rule, deg)
    space = lib.Space.from_mesh(ctx, mesh)
    op = lib.Op(ctx, lib.OP_ELLIP, lib.EllipParam(p, 1.0, lam), rule, [space], [efs], mesh.coords)
    op.set_forcing(o.forcing(op.quadrature_points()))
    op.set_dirichlet(0, o.dirichlet_cache())
    rp, ci = op.q1_pattern(mesh.l2g)
    x, fnorms, hists = pl.newton(op.function, op.mult, lambda x: lu_pc(op.assemble_q1(x, mesh.l2g, None, None), rp, ci, mesh.n), np.zeros(mesh.n), linesearch=ls)
    check_history(name, fnorms, hists)
    check_error_lines(name, o, x, function=op.function)          # the example's closing diagnostics at the GPU solution (the oracle evaluates the error integrals)
    op.close()


def test_refout_ellip_morph_cg_history_gpu(ctx):
    mesh = lib.BoxMesh((4, 4, 4), 3, morph=1.0, twist=0.4)
    o = ol.OracleEllip(mesh, p=2.0, exact=0)
    rule = lib.Rule.preferred(ctx, 3)
    efs = lib.EFS(ctx, rule, 3)
    space = lib.Space.from_mesh(ctx, mesh)
    op = lib.Op(ctx, lib.OP_ELLIP, lib.EllipParam(2.0, 1.0, 0.0), rule, [space], [efs], mesh.coords)
    op.set_forcing(o.forcing(op.quadrature_points()))
    op.set_dirichlet(0, o.dirichlet_cache())
    rp, ci = op.q1_pattern(mesh.l2g)
    x0 = np.zeros(mesh.n)
    F = op.function(x0)
    y, hist = pl.cg(op.mult, lu_pc(op.assemble_q1(x0, mesh.l2g, None, None), rp, ci, mesh.n), F, rtol=1e-8)
    check_history("ellip-morph", [], [hist])
    check_error_lines("ellip-morph", o, x0 - y, algebraic=True, function=op.function)
    op.close()


def test_refout_elast_ilu0_history_gpu(ctx):
    """elast-e0-b3-g01.refout: 10 Newton steps (cubic backtracking line search), GMRES(30) with restarts preconditioned by block ILU(0)
    of the BAIJ(3) embedded-Q1 matrix assembled on the GPU, natural ordering."""
    mesh = lib.BoxMesh((4, 4, 4), 3)
    o = ol.OracleElast(mesh, exact=0, a=0.98, b=1.01, c=1.01)
    rule = lib.Rule.preferred(ctx, 3)
    efs = lib.EFS(ctx, rule, 3)
    space = lib.Space.from_mesh(ctx, mesh, bs=3)
    op = lib.Op(ctx, lib.OP_ELAST, lib.ElastParam(o.lam, o.mu), rule, [space], [efs], mesh.coords)
    op.set_forcing(o.forcing(op.quadrature_points()))
    op.set_dirichlet(0, o.dirichlet_cache())
    brp, bci = op.q1_pattern(mesh.l2g)
    x, fnorms, hists = pl.newton(op.function, op.mult, lambda x: pl.block_ilu0_numpy(brp, bci, op.assemble_q1(x, mesh.l2g, None, None).reshape(-1, 3, 3), 3),
                                 np.zeros(mesh.n * 3), linesearch="bt")
    check_history("elast-e0-b3-g01", fnorms, hists, last_digit=2)
    check_error_lines("elast-e0-b3-g01", o, x, bs=3, function=op.function)
    op.close()
