"""Operator-level parity on the GPU for the elast and stokes examples (through the C ABI) against the oracle restatement
of elast.c:220-341 and stokes.c:447-673, plus the reference's golden known answers.  Bar 1e-12 relative (FP64)."""
import numpy as np
import pytest

import oracle_lib as ol
from dohp_b200 import lib

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def ctx():
    import torch
    ol.build_oracle(ref=False)
    c = lib.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    yield c
    c.close()


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def _fmt6(x):
    return float("%.6g" % x)


def make_elast(ctx, mesh, o):
    rule = lib.Rule.preferred(ctx, mesh.p)
    efs = lib.EFS(ctx, rule, mesh.p)
    space = lib.Space.from_mesh(ctx, mesh, bs=3)
    op = lib.Op(ctx, lib.OP_ELAST, lib.ElastParam(o.lam, o.mu), rule, [space], [efs], mesh.coords)
    op.set_forcing(o.forcing(op.quadrature_points()))
    op.set_dirichlet(0, o.dirichlet_cache())
    return op


def make_stokes(ctx, mu, mp, o):
    rule = lib.Rule.preferred(ctx, mu.p)
    efsu, efsp = lib.EFS(ctx, rule, mu.p), lib.EFS(ctx, rule, mp.p)
    su, sp = lib.Space.from_mesh(ctx, mu, bs=3), lib.Space.from_mesh(ctx, mp, bs=1)
    op = lib.Op(ctx, lib.OP_STOKES, lib.StokesParam(o.prm.B, o.prm.eps, o.prm.p, o.prm.nu), rule, [su, sp], [efsu, efsp], mu.coords)
    op.set_forcing(o.forcing(op.quadrature_points()))
    op.set_dirichlet(0, o.dirichlet_cache())
    return op


def test_elast_golden_function_norm(ctx):
    """elast-e0-b3-g01.refout:1 '0 SNES Function norm 14965.9' (src/fs/tests/CMakeLists.txt:60)."""
    mesh = lib.BoxMesh((4, 4, 4), 3)
    o = ol.OracleElast(mesh, exact=0, a=0.98, b=1.01, c=1.01)
    op = make_elast(ctx, mesh, o)
    F = op.function(np.zeros(mesh.n * 3))
    assert _fmt6(np.linalg.norm(F)) == 14965.9
    assert relerr(F, o.function(np.zeros(mesh.n * 3))) <= TOL


@pytest.mark.parametrize("deg", range(1, 11))
@pytest.mark.parametrize("morph", [None, dict(morph=0.25, twist=0.2)])
def test_elast_function_and_mult_match_oracle(ctx, deg, morph):
    rng = np.random.default_rng(10 + deg)
    mesh = lib.BoxMesh((2, 2, 3) if deg < 8 else (2, 2, 2), deg, **(morph or {}))
    o = ol.OracleElast(mesh, exact=0, a=0.6, b=0.5, c=0.4, scale=0.2)
    op = make_elast(ctx, mesh, o)
    x = 0.05 * rng.standard_normal(mesh.n * 3)
    assert relerr(op.function(x), o.function(x)) <= TOL
    assert relerr(op.stash(), o.stash) <= TOL                  # struct ElastStore {F[9]}
    dx = rng.standard_normal(mesh.n * 3)
    assert relerr(op.mult(dx), o.mult(dx)) <= TOL
    op.close()


def _stokes_meshes(ne, deg, **kw):
    return lib.BoxMesh(ne, deg, **kw), lib.BoxMesh(ne, deg - 1, dirichlet=0, **kw)


def test_stokes_golden_norms(ctx):
    """stokes-q3q2-scr.refout:1-3: 0.00343338 (residual at the exact solution), 0.00343338 (J*soln + F(0), the
    matrix-free MatMult_Nest_StokesCoupled), 1.75588 (||F(0)||)."""
    mu, mp = _stokes_meshes((4, 4, 4), 3)
    o = ol.OracleStokes(mu, mp, exact=1)
    op = make_stokes(ctx, mu, mp, o)
    z = [np.zeros(mu.n * 3), np.zeros(mp.n)]
    F0 = np.concatenate(op.function(z))
    assert _fmt6(np.linalg.norm(F0)) == 1.75588
    soln = [o.exact_solution(mu.node_xyz[:mu.n])[0].reshape(-1), o.exact_solution(mp.node_xyz[:mp.n])[1]]
    assert _fmt6(np.linalg.norm(np.concatenate(op.function(soln)))) == 0.00343338
    op.function(z)
    assert _fmt6(np.linalg.norm(np.concatenate(op.mult(soln)) + F0)) == 0.00343338


def test_stokes_fgmres_golden_norms(ctx):
    """stokes-q3q2-fgmres.refout:1-3 (power-law rheology p=1.2, eps=0.1): 0.00551936 (residual at the interpolated exact solution), 1.95026
    (J(soln) soln + F(0): the coupled matrix-free Jacobian with the stash left by F(soln), stokes.c:1125-1133), 1.81231 (||F(0)||)."""
    mu, mp = _stokes_meshes((4, 4, 4), 3)
    o = ol.OracleStokes(mu, mp, exact=1, p=1.2, eps=0.1)
    op = make_stokes(ctx, mu, mp, o)
    F0 = np.concatenate(op.function([np.zeros(mu.n * 3), np.zeros(mp.n)]))
    assert _fmt6(np.linalg.norm(F0)) == 1.81231
    soln = [o.exact_solution(mu.node_xyz[:mu.n])[0].reshape(-1), o.exact_solution(mp.node_xyz[:mp.n])[1]]
    assert _fmt6(np.linalg.norm(np.concatenate(op.function(soln)))) == 0.00551936
    assert _fmt6(np.linalg.norm(np.concatenate(op.mult(soln)) + F0)) == 1.95026
    op.close()


@pytest.mark.parametrize("deg", range(2, 11))
@pytest.mark.parametrize("morph", [None, dict(morph=0.2, twist=0.15)])
def test_stokes_function_mult_and_blocks_match_oracle(ctx, deg, morph):
    rng = np.random.default_rng(20 + deg)
    mu, mp = _stokes_meshes((2, 3, 2) if deg < 8 else (2, 2, 2), deg, **(morph or {}))
    o = ol.OracleStokes(mu, mp, exact=1, p=1.2, eps=0.1, nu=0.3)       # rheology of stokes-q3q2-fgmres, CMakeLists.txt:65-70
    op = make_stokes(ctx, mu, mp, o)
    xu, xp = 0.1 * rng.standard_normal(mu.n * 3), 0.1 * rng.standard_normal(mp.n)
    Fu, Fp = op.function([xu, xp])
    Ou, Op_ = o.function(xu, xp)
    assert relerr(Fu, Ou) <= TOL and relerr(Fp, Op_) <= TOL
    assert relerr(op.stash(), o.stash) <= TOL                          # struct StokesStore {eta, deta, Du[6]}
    du, dp = rng.standard_normal(mu.n * 3), rng.standard_normal(mp.n)
    Ju, Jp = op.mult([du, dp])
    Ru, Rp = o.mult(du, dp)
    assert relerr(Ju, Ru) <= TOL and relerr(Jp, Rp) <= TOL
    Au = op.mult([du, None], block=lib.STOKES_A)[0]
    assert relerr(Au, o.mult(du, None, block=1)[0]) <= TOL
    Bp = op.mult([du, None], block=lib.STOKES_B)[1]
    assert relerr(Bp, o.mult(du, None, block=2)[1]) <= TOL
    Btu = op.mult([None, dp], block=lib.STOKES_Bt)[0]
    assert relerr(Btu, o.mult(None, dp, block=3)[0]) <= TOL
    op.close()
