"""Known answers of the reference's elast and stokes goldens reproduced by the oracle restatement.
Forcing terms come from oracle/gen_exact.py (sympy regeneration of elastexact.py / stokesexact.py).  CPU only."""
import numpy as np
import pytest

import oracle_lib as ol
from dohp_b200 import lib


@pytest.fixture(scope="module", autouse=True)
def _build():
    ol.build_oracle()
    lib.build()


def _fmt6(x):
    return float("%.6g" % x)


def test_elast_e0_b3_golden_function_norm():
    """src/fs/tests/CMakeLists.txt:60 elast-e0-b3-g01: '0 SNES Function norm 14965.9' (refout line 1)."""
    mesh = lib.BoxMesh((4, 4, 4), 3)
    o = ol.OracleElast(mesh, exact=0, a=0.98, b=1.01, c=1.01)
    F = o.function(np.zeros(mesh.n * 3))
    assert _fmt6(np.linalg.norm(F)) == 14965.9


def test_elast_matrix_free_jacobian_is_derivative_of_residual():
    mesh = lib.BoxMesh((2, 2, 3), 2, morph=0.2, twist=0.2)
    o = ol.OracleElast(mesh, exact=0, a=0.5, b=0.4, c=0.3, scale=0.1)
    rng = np.random.default_rng(1)
    x, dx = 0.05 * rng.standard_normal(mesh.n * 3), rng.standard_normal(mesh.n * 3)
    h = 1e-6
    Fp, Fm = o.function(x + h * dx).copy(), o.function(x - h * dx).copy()
    o.function(x)
    Jdx = o.mult(dx)
    fd = (Fp - Fm) / (2 * h)
    assert np.allclose(Jdx, fd, rtol=0, atol=2e-7 * np.abs(fd).max())


def _stokes_meshes(ne=(4, 4, 4), deg=3, codim=1, **kw):
    # velocity Dirichlet on all walls, no pressure boundary condition (stokes.c:183-214)
    return lib.BoxMesh(ne, deg, **kw), lib.BoxMesh(ne, deg - codim, dirichlet=0, **kw)


def test_stokes_q3q2_golden_norms():
    """stokes-q3q2-scr (CMakeLists.txt:62-64): ||F(soln)|| = 0.00343338 (refout:1), ||J soln + F(0)|| = 0.00343338 (refout:2,
    pins the matrix-free MatMult_Nest_StokesCoupled), '0 SNES Function norm 1.75588' (refout:3)."""
    mu, mp = _stokes_meshes()
    o = ol.OracleStokes(mu, mp, exact=1)
    xu0, xp0 = np.zeros(mu.n * 3), np.zeros(mp.n)
    Fu, Fp = o.function(xu0, xp0)
    F0 = np.concatenate([Fu, Fp])
    assert _fmt6(np.linalg.norm(F0)) == 1.75588
    su, sp = o.exact_solution(mu.node_xyz[:mu.n])[0].reshape(-1), o.exact_solution(mp.node_xyz[:mp.n])[1]
    Fu, Fp = o.function(su, sp)
    assert _fmt6(np.linalg.norm(np.concatenate([Fu, Fp]))) == 0.00343338
    o.function(xu0, xp0)                               # linear problem: the stash does not matter, but follow stokes.c:1125-1133
    Ju, Jp = o.mult(su, sp)
    assert _fmt6(np.linalg.norm(np.concatenate([Ju, Jp]) + F0)) == 0.00343338


def test_stokes_q3q2_fgmres_golden_norms():
    """stokes-q3q2-fgmres (CMakeLists.txt:65-70: -rheo_p 1.2 -rheo_eps 0.1, the power-law rheology): refout:1-3 print ||F(soln)|| =
    0.00551936, ||J(soln) soln + F(0)|| = 1.95026 (the matrix-free coupled Jacobian linearised AT the interpolated exact solution:
    stokes.c:1125-1133 evaluates F(0), then F(soln) -- which leaves the stash there -- then the product) and '0 SNES Function norm 1.81231':
    pins the nonlinear viscosity, its derivative term and the stash."""
    mu, mp = _stokes_meshes()
    o = ol.OracleStokes(mu, mp, exact=1, p=1.2, eps=0.1)
    F0 = np.concatenate(o.function(np.zeros(mu.n * 3), np.zeros(mp.n)))
    assert _fmt6(np.linalg.norm(F0)) == 1.81231
    su, sp = o.exact_solution(mu.node_xyz[:mu.n])[0].reshape(-1), o.exact_solution(mp.node_xyz[:mp.n])[1]
    assert _fmt6(np.linalg.norm(np.concatenate(o.function(su, sp)))) == 0.00551936
    assert _fmt6(np.linalg.norm(np.concatenate(o.mult(su, sp)) + F0)) == 1.95026


def test_stokes_blocks_sum_to_coupled_operator():
    mu, mp = _stokes_meshes((2, 3, 2), 3, 1, morph=0.2, twist=0.1)
    o = ol.OracleStokes(mu, mp, exact=1, p=1.3, eps=0.2)
    rng = np.random.default_rng(2)
    xu, xp = rng.standard_normal(mu.n * 3), rng.standard_normal(mp.n)
    o.function(0.1 * xu, 0.1 * xp)
    Ju, Jp = o.mult(xu, xp)
    Au, _ = o.mult(xu, None, block=1)
    _, Bp = o.mult(xu, None, block=2)
    Btu, _ = o.mult(None, xp, block=3)
    assert np.allclose(Ju, Au + Btu, rtol=0, atol=1e-12 * np.abs(Ju).max())
    assert np.allclose(Jp, Bp, rtol=0, atol=1e-12 * np.abs(Jp).max())


def test_stokes_velocity_block_restatement_properties():
    """or_stokes_jacobian_q1_velocity (StokesJacobianAssemble_Velocity, stokes.c:675-737; Ap itself is pinned by the Stokes goldens replayed in
    tests/test_stokes_refout_replay.py, Mdiag by nothing upstream): the block
    must be symmetric, annihilate rigid translations away from the Dirichlet boundary, and reduce to eta-scaled
    sym-gradient stiffness; Mdiag must be positive."""
    import scipy.sparse as sp
    from dohp_b200.box import q1_pattern
    mu_ = lib.BoxMesh((3, 3, 3), 2, morph=0.2, twist=0.1)
    mp_ = lib.BoxMesh((3, 3, 3), 1, dirichlet=0, morph=0.2, twist=0.1)
    o = ol.OracleStokes(mu_, mp_, B=1.3, eps=0.1, p=1.2, nu=0.2, exact=1, scale=0.3)
    rp, ci = q1_pattern(mu_, bs=3)
    xu = 0.1 * np.random.default_rng(0).standard_normal(mu_.n * 3)
    vals, md = o.jacobian_q1_velocity(xu, rp, ci)
    J = sp.csr_matrix((vals, ci, rp), shape=(mu_.n * 3, mu_.n * 3))
    assert abs(J - J.T).max() <= 1e-12 * abs(J).max()
    brp, _ = q1_pattern(mu_, bs=1)
    interior = np.repeat(np.diff(brp) == 27, 3)
    assert interior.any()
    for f in range(3):
        t = np.zeros((mu_.n, 3)); t[:, f] = 1
        assert np.abs((J @ t.reshape(-1))[interior]).max() <= 1e-12 * abs(J).max()
    assert md.min() > 0 and np.allclose(md[0::3], md[1::3]) and np.allclose(md[0::3], md[2::3])


def test_stokes_pressure_blocks_restatement_properties():
    """or_stokes_jacobian_pressure (StokesJacobianAssemble_Pressure, stokes.c:739-794; D is pinned by the Stokes goldens replayed in
    tests/test_stokes_refout_replay.py, Daux -- used by PCLSC only -- by nothing upstream): D is the
    1/eta-weighted pressure mass matrix (symmetric positive definite, 1^T D 1 = int 1/eta), Daux the Neumann Laplacian (symmetric
    positive semi-definite with the constants in its null space -- exactly so on the affine mesh, where quirk (iii) of
    dEFSGetExplicitSparse_Basic is the identity)."""
    import scipy.sparse as sp
    from dohp_b200.box import dense_pattern
    mu_ = lib.BoxMesh((2, 2, 2), 3)
    mp_ = lib.BoxMesh((2, 2, 2), 2, dirichlet=0)
    o = ol.OracleStokes(mu_, mp_, B=2.0, eps=0.1, p=2.0, nu=0.0, exact=1, scale=0.3)      # p = 2: eta = B everywhere
    rp, ci = dense_pattern(mp_)
    D, Da = o.jacobian_pressure(np.zeros(mu_.n * 3), rp, ci, dcache=np.zeros((mu_.nc - mu_.n) * 3))
    D = sp.csr_matrix((D, ci, rp), shape=(mp_.n, mp_.n))
    Da = sp.csr_matrix((Da, ci, rp), shape=(mp_.n, mp_.n))
    one = np.ones(mp_.n)
    assert abs(D - D.T).max() <= 1e-13 * abs(D).max() and abs(Da - Da.T).max() <= 1e-12 * abs(Da).max()
    assert abs(one @ (D @ one) - 8.0 / 2.0) <= 1e-12                    # volume of [-1,1]^3 over eta = B = 2
    assert np.abs(Da @ one).max() <= 1e-11 * abs(Da).max()
    assert np.linalg.eigvalsh(D.toarray()).min() > 0 and np.linalg.eigvalsh(Da.toarray()).min() > -1e-10
