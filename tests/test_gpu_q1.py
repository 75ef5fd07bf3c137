"""Embedded-Q1 preconditioner on the GPU (dB200OpQ1Pattern / dB200OpAssembleQ1 through the C ABI): the sparsity pattern must
be bit-exact to what dFSCreateMatrix + the patch loop + MatAssembly produce (restated generically in numpy by
dohp_b200.box.q1_pattern), the values must match the oracle restatement of EllipJacobian / ElastJacobian incl. the
reference's quirks, and the LU-preconditioned residual norms must reproduce the reference's golden KSP lines."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import oracle_lib as ol
from dohp_b200 import lib
from dohp_b200.box import dense_pattern, q1_pattern

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


def make_ellip(ctx, mesh, o, **prm):
    rule = lib.Rule.preferred(ctx, mesh.p)
    efs = lib.EFS(ctx, rule, mesh.p)
    space = lib.Space.from_mesh(ctx, mesh)
    op = lib.Op(ctx, lib.OP_ELLIP, lib.EllipParam(prm.get("p", 2.0), prm.get("eps", 1.0), prm.get("lam", 0.0)), rule, [space], [efs], mesh.coords)
    op.set_forcing(o.forcing(op.quadrature_points()))
    op.set_dirichlet(0, o.dirichlet_cache())
    return op


@pytest.mark.parametrize("deg,fnorm,knorm", [(1, 3.17818, 0.939628), (3, 4.93408, 158.134)])
def test_golden_ksp_norms(ctx, deg, fnorm, knorm):
    """ellip-e0-b{1,3}-p16.refout:1-2: ||F(0)|| and ||LU(Jp)^-1 F(0)|| (left-preconditioned GMRES monitor), pins quirk (i)."""
    mesh = lib.BoxMesh((4, 4, 4), deg)
    o = ol.OracleEllip(mesh, p=1.6, exact=0, a=1.6, b=1.7, c=1.8)
    op = make_ellip(ctx, mesh, o, p=1.6)
    x0 = np.zeros(mesh.n)
    F = op.function(x0)
    assert _fmt6(np.linalg.norm(F)) == fnorm
    rowptr, colidx = op.q1_pattern(mesh.l2g)
    vals = op.assemble_q1(x0, mesh.l2g, rowptr, colidx)
    Jp = sp.csr_matrix((vals, colidx, rowptr), shape=(mesh.n, mesh.n))
    assert _fmt6(np.linalg.norm(spla.splu(Jp.tocsc()).solve(F))) == knorm


def test_golden_morph_ksp_norm(ctx):
    """ellip-morph.refout:1 '0 KSP Residual norm 111.034' pins quirk (ii) (non-affine mesh)."""
    mesh = lib.BoxMesh((4, 4, 4), 3, morph=1.0, twist=0.4)
    o = ol.OracleEllip(mesh, p=2.0, exact=0)
    op = make_ellip(ctx, mesh, o)
    x0 = np.zeros(mesh.n)
    F = op.function(x0)
    rowptr, colidx = op.q1_pattern(mesh.l2g)
    vals = op.assemble_q1(x0, mesh.l2g, rowptr, colidx)
    Jp = sp.csr_matrix((vals, colidx, rowptr), shape=(mesh.n, mesh.n))
    assert _fmt6(np.linalg.norm(spla.splu(Jp.tocsc()).solve(F))) == 111.034


@pytest.mark.parametrize("deg", [1, 2, 3, 4, 5, 7, 10])
@pytest.mark.parametrize("morph", [None, dict(morph=0.3, twist=0.25)])
def test_ellip_pattern_and_values_match_oracle(ctx, deg, morph):
    mesh = lib.BoxMesh((3, 2, 2) if deg < 7 else (2, 2, 1), deg, dirichlet=lib.WALLS | lib.BOTTOM, **(morph or {}))
    prm = dict(p=1.6, lam=0.7, eps=0.9)
    o = ol.OracleEllip(mesh, exact=0, a=1.1, b=0.9, c=1.3, **prm)
    op = make_ellip(ctx, mesh, o, **prm)
    rowptr, colidx = op.q1_pattern(mesh.l2g)
    rp, ci = q1_pattern(mesh)                                   # generic union-of-patches restatement (numpy)
    assert np.array_equal(rowptr, rp) and np.array_equal(colidx, ci)      # bit-exact pattern and column indices
    x = 0.3 * np.random.default_rng(deg).standard_normal(mesh.n)
    vals = op.assemble_q1(x, mesh.l2g, rowptr, colidx)
    ref = o.jacobian_q1(x, rp, ci)
    assert relerr(vals, ref) <= TOL
    # the cached pattern may be used implicitly
    vals2 = op.assemble_q1(x, mesh.l2g, None, None, vals=np.empty_like(vals))
    assert relerr(vals2, ref) <= TOL


@pytest.mark.parametrize("deg", [1, 2, 3, 4, 5, 6, 7, 9])
@pytest.mark.parametrize("morph", [None, dict(morph=0.25, twist=0.2)])
def test_elast_pattern_and_values_match_oracle(ctx, deg, morph):
    mesh = lib.BoxMesh((2, 2, 2) if deg < 6 else (2, 1, 1) if deg < 9 else (1, 1, 1), deg, **(morph or {}))
    o = ol.OracleElast(mesh, exact=0, a=0.6, b=0.5, c=0.4, scale=0.2)
    rule = lib.Rule.preferred(ctx, mesh.p)
    efs = lib.EFS(ctx, rule, mesh.p)
    space = lib.Space.from_mesh(ctx, mesh, bs=3)
    op = lib.Op(ctx, lib.OP_ELAST, lib.ElastParam(o.lam, o.mu), rule, [space], [efs], mesh.coords)
    op.set_dirichlet(0, o.dirichlet_cache())
    rowptr, colidx = op.q1_pattern(mesh.l2g)                    # block CSR = BAIJ(3) structure (elast.c:479 default -q1mat_type baij)
    brp, bci = q1_pattern(mesh, bs=1)
    assert np.array_equal(rowptr, brp) and np.array_equal(colidx, bci)
    x = 0.05 * np.random.default_rng(deg).standard_normal(mesh.n * 3)
    vals = op.assemble_q1(x, mesh.l2g, rowptr, colidx)
    J = sp.bsr_matrix((vals.reshape(-1, 3, 3), colidx, rowptr), shape=(mesh.n * 3, mesh.n * 3)).tocsr()
    J.sort_indices()
    rp, ci = q1_pattern(mesh, bs=3)                             # the same structure expanded to scalar AIJ rows
    ref = sp.csr_matrix((o.jacobian_q1(x, rp, ci), ci, rp), shape=J.shape)
    assert np.abs((J - ref).data).max() <= TOL * np.abs(ref.data).max() and J.nnz == ref.nnz
    # -dfs_assemble_reduced (fs.c:905-917): only the component-diagonal of every block
    op.q1_set_reduced(True)
    red = op.assemble_q1(x, mesh.l2g, rowptr, colidx).reshape(-1, 3, 3)
    full = vals.reshape(-1, 3, 3)
    eye = np.eye(3, dtype=bool)
    assert np.abs(red[:, eye] - full[:, eye]).max() <= TOL * np.abs(full).max() and not red[:, ~eye].any()


@pytest.mark.parametrize("deg,nu", [(2, 0.3), (3, 0.0), (4, 0.25), (6, 0.0)])
@pytest.mark.parametrize("morph", [None, dict(morph=0.25, twist=0.2)])
def test_stokes_velocity_block_and_lumped_mass_match_oracle(ctx, deg, nu, morph):
    """StokesJacobianAssemble_Velocity (stokes.c:675-737): the BAIJ-3 velocity block of the preconditioner (power-law rheology,
    StokesPointwiseJacobian_A with the nu penalty) and the Mdiag vector vs the oracle's literal restatement (the oracle's Ap is pinned
    by the reference's Stokes goldens, tests/test_stokes_refout_replay.py; Mdiag is used by no upstream test)."""
    mu_ = lib.BoxMesh((2, 2, 2) if deg < 4 else (2, 1, 1), deg, **(morph or {}))
    mp_ = lib.BoxMesh(mu_.ne, deg - 1, dirichlet=0, **(morph or {}))          # -pressure_codim 1
    prm = dict(B=1.3, eps=0.1, p=1.2, nu=nu)
    o = ol.OracleStokes(mu_, mp_, exact=1, a=0.9, b=1.1, c=0.8, scale=0.5, **prm)
    rule = lib.Rule.preferred(ctx, deg)
    efsu, efsp = lib.EFS(ctx, rule, deg), lib.EFS(ctx, rule, mp_.p)
    su, sp_ = lib.Space.from_mesh(ctx, mu_, bs=3), lib.Space.from_mesh(ctx, mp_, bs=1)
    op = lib.Op(ctx, lib.OP_STOKES, lib.StokesParam(prm["B"], prm["eps"], prm["p"], prm["nu"]), rule, [su, sp_], [efsu, efsp], mu_.coords)
    op.set_dirichlet(0, o.dirichlet_cache())
    rowptr, colidx = op.q1_pattern(mu_.l2g)
    brp, bci = q1_pattern(mu_, bs=1)
    assert np.array_equal(rowptr, brp) and np.array_equal(colidx, bci)
    xu = 0.2 * np.random.default_rng(deg).standard_normal(mu_.n * 3)
    vals = op.assemble_q1([xu, np.zeros(mp_.n)], mu_.l2g, rowptr, colidx)
    J = sp.bsr_matrix((vals.reshape(-1, 3, 3), colidx, rowptr), shape=(mu_.n * 3, mu_.n * 3)).tocsr()
    J.sort_indices()
    rp, ci = q1_pattern(mu_, bs=3)
    rvals, rmd = o.jacobian_q1_velocity(xu, rp, ci)
    ref = sp.csr_matrix((rvals, ci, rp), shape=J.shape)
    assert np.abs((J - ref).data).max() <= TOL * np.abs(ref.data).max() and J.nnz == ref.nnz
    md = op.stokes_lumped_mass()
    assert relerr(md, rmd) <= TOL
    op.close()


def test_q1_full_size_properties(ctx):
    """BASELINE.json size (ellip 64^3 elements, p=4: 4.4e8 nonzeros) is out of the oracle's reach; size-independent
    properties of the embedded-Q1 Laplacian instead: symmetry, zero row sums away from the Dirichlet boundary (the constants
    are in the null space of every patch matrix), positive diagonal, and idempotence of the assembly (two assemblies into
    the same array give the same values: the callee zeroes, ellip.c:489)."""
    import torch
    mesh = lib.BoxMesh((64, 64, 64), 4, node_xyz=False, ghost_l2g=False)
    rule = lib.Rule.preferred(ctx, 4)
    efs = lib.EFS(ctx, rule, 4)
    space = lib.Space.from_mesh(ctx, mesh)
    op = lib.Op(ctx, lib.OP_ELLIP, lib.EllipParam(2.0, 1.0, 0.0), rule, [space], [efs], mesh.coords)
    rowptr, colidx = op.q1_pattern(mesh.l2g)
    assert len(rowptr) == mesh.n + 1 and rowptr[-1] == len(colidx) == 444194947
    x = torch.zeros(mesh.n, dtype=torch.float64, device="cuda")
    vals = torch.empty(len(colidx), dtype=torch.float64, device="cuda")
    l2g = torch.from_numpy(np.ascontiguousarray(mesh.l2g)).cuda()
    op.assemble_q1(x, l2g, None, None, vals)
    ctx.synchronize()
    first = vals.clone()
    op.assemble_q1(x, l2g, None, None, vals)
    ctx.synchronize()
    assert (vals - first).abs().max().item() <= 1e-13 * first.abs().max().item()      # RED order may differ
    del first
    rp, ci = torch.from_numpy(rowptr.astype(np.int64)).cuda(), torch.from_numpy(colidx.astype(np.int64)).cuda()
    A = torch.sparse_csr_tensor(rp, ci, vals, size=(mesh.n, mesh.n))
    g = torch.Generator(device="cuda").manual_seed(7)
    u = torch.randn(mesh.n, dtype=torch.float64, device="cuda", generator=g)
    v = torch.randn(mesh.n, dtype=torch.float64, device="cuda", generator=g)
    Au, Av = A @ u, A @ v
    suv, svu = torch.dot(v, Au).item(), torch.dot(u, Av).item()
    assert abs(suv - svu) <= 1e-11 * torch.linalg.norm(v).item() * torch.linalg.norm(Au).item()
    assert torch.dot(u, Au).item() > 0
    rows = torch.repeat_interleave(torch.arange(mesh.n, device="cuda"), rp[1:] - rp[:-1])
    diag = vals[rows == ci]
    assert diag.numel() == mesh.n and diag.min().item() > 0
    rowsum = A @ torch.ones(mesh.n, dtype=torch.float64, device="cuda")
    interior = (rp[1:] - rp[:-1]) == 27                       # all 27 neighbours present: no Dirichlet column was dropped
    assert interior.sum().item() > 0.9 * mesh.n
    assert rowsum[interior].abs().max().item() <= 1e-12 * diag.max().item()
    op.close()


@pytest.mark.parametrize("ne", [24, 48])                             # 48^3 p=6 = BASELINE.json configs[2] (6.3e8 BAIJ-3 blocks, 46 GB of values)
def test_q1_elast_rigid_translations_are_in_the_null_space(ctx, ne):
    """elast ne^3 elements p=6 (BAIJ-3) linearised at u = 0: away from the Dirichlet boundary every block row annihilates the three
    rigid translations, sum_j K[(i,fi),(j,fj)] = 0, and the matrix is symmetric."""
    import torch
    mesh = lib.BoxMesh((ne, ne, ne), 6, node_xyz=False, ghost_l2g=False)
    rule = lib.Rule.preferred(ctx, 6)
    efs = lib.EFS(ctx, rule, 6)
    space = lib.Space.from_mesh(ctx, mesh, bs=3)
    op = lib.Op(ctx, lib.OP_ELAST, lib.ElastParam(0.5769, 0.3846), rule, [space], [efs], mesh.coords)
    rowptr, colidx = op.q1_pattern(mesh.l2g)
    x = torch.zeros(mesh.n * 3, dtype=torch.float64, device="cuda")
    vals = torch.empty(len(colidx) * 9, dtype=torch.float64, device="cuda")
    op.assemble_q1(x, torch.from_numpy(np.ascontiguousarray(mesh.l2g)).cuda(), None, None, vals)
    ctx.synchronize()
    rp, ci = torch.from_numpy(rowptr.astype(np.int64)).cuda(), torch.from_numpy(colidx.astype(np.int64)).cuda()
    K = vals.view(-1, 3, 3)
    rows = torch.repeat_interleave(torch.arange(mesh.n, device="cuda"), rp[1:] - rp[:-1])

    def matvec(v):          # block-CSR product with 64-bit indexing (6.3e8 blocks x 9 values exceed the 32-bit limits of the sparse libraries)
        y = torch.zeros(mesh.n, 3, dtype=torch.float64, device="cuda")
        step = 1 << 26
        for b0 in range(0, K.shape[0], step):
            sl = slice(b0, min(b0 + step, K.shape[0]))
            y.index_add_(0, rows[sl], torch.einsum("bij,bj->bi", K[sl], v.view(-1, 3)[ci[sl]]))
        return y.reshape(-1)

    interior = torch.repeat_interleave((rp[1:] - rp[:-1]) == 27, 3)
    scale = vals.abs().max().item()
    for f in range(3):
        t = torch.zeros(mesh.n, 3, dtype=torch.float64, device="cuda")
        t[:, f] = 1
        r = matvec(t.reshape(-1))
        assert r[interior].abs().max().item() <= 1e-12 * scale
    g = torch.Generator(device="cuda").manual_seed(3)
    u = torch.randn(mesh.n * 3, dtype=torch.float64, device="cuda", generator=g)
    v = torch.randn(mesh.n * 3, dtype=torch.float64, device="cuda", generator=g)
    suv, svu = (v * matvec(u)).sum().item(), (u * matvec(v)).sum().item()
    assert abs(suv - svu) <= 1e-11 * abs(suv)
    op.close()
    del K, vals, rp, ci, rows, u, v
    torch.cuda.empty_cache()


def test_q1_errors_are_loud(ctx):
    mesh = lib.BoxMesh((2, 2, 2), 2)
    op = make_ellip(ctx, mesh, ol.OracleEllip(mesh, exact=0))
    rowptr, colidx = op.q1_pattern(mesh.l2g)
    diag_only = np.arange(mesh.n + 1, dtype=np.int32), np.arange(mesh.n, dtype=np.int32)
    with pytest.raises(lib.DohpError, match="does not contain every coupling"):
        op.assemble_q1(np.zeros(mesh.n), mesh.l2g, *diag_only)
    op2 = make_ellip(ctx, mesh, ol.OracleEllip(mesh, exact=0))
    with pytest.raises(lib.DohpError, match="without a pattern"):
        op2.assemble_q1(np.zeros(mesh.n), mesh.l2g, None, None, vals=np.zeros(len(colidx)))


@pytest.mark.parametrize("deg", [1, 2, 3, 4, 5, 6, 8, 10])        # degree <= 4: element matrix in shared memory; above: column path
@pytest.mark.parametrize("morph", [None, dict(morph=0.5, twist=0.1)])
def test_ellip_fast_assembly_matches_oracle_and_is_the_exact_jacobian(ctx, deg, morph):
    """-ellip_jac_qmethod fast (dEFSGetExplicitSparse_Basic, jacobi.c:611-643): element-dense pattern bit-exact, values vs the
    oracle, and the property ellip-morph-qfast.refout pins upstream (Richardson + LU converges in one step): the assembled
    matrix IS the matrix-free Jacobian, J_assembled x == MatMult(x)."""
    mesh = lib.BoxMesh((2, 2, 2) if deg <= 6 else (1, 2, 1), deg, dirichlet=lib.WALLS | lib.BOTTOM, **(morph or {}))
    prm = dict(p=1.9, lam=0.4, eps=1.0)
    o = ol.OracleEllip(mesh, exact=0, jac_qmethod=ol.FAST, **prm)
    op = make_ellip(ctx, mesh, o, **prm)
    rowptr, colidx = op.dense_pattern(0, mesh.l2g)
    rp, ci = dense_pattern(mesh)
    assert np.array_equal(rowptr, rp) and np.array_equal(colidx, ci)
    x = 0.3 * np.random.default_rng(deg).standard_normal(mesh.n)
    vals = op.assemble_dense(lib.DENSE_ELLIP_FAST, x, mesh.l2g, rowptr, colidx)
    ref = o.jacobian_q1(x, rp, ci)
    assert relerr(vals, ref) <= TOL
    op.function(x)                                              # stash at x: mult is J(x)
    v = np.random.default_rng(7).standard_normal(mesh.n)
    Jv = op.mult(v)
    J = sp.csr_matrix((vals, colidx, rowptr), shape=(mesh.n, mesh.n))
    if deg > 1:           # degree 1 takes the Q1 path with the reference's deriv1 = +-1 quirk (efstopo.c:668): not the exact Jacobian
        assert np.abs(J @ v - Jv).max() <= 1e-11 * np.abs(Jv).max()
    op.close()


@pytest.mark.parametrize("deg", [1, 2, 3, 4, 5])                    # degree <= 4: element matrix in shared memory; 5: column path
@pytest.mark.parametrize("morph", [None, dict(morph=0.3, twist=0.1)])
def test_elast_fast_assembly_matches_oracle_and_is_the_exact_jacobian(ctx, deg, morph):
    """-elast_jac_qmethod fast (ElastJacobian on the FAST rule through dEFSGetExplicitSparse_Basic, elast.c:343-407, jacobi.c:611-643):
    BAIJ(3) blocks on the element-dense node pattern against the oracle, and J_assembled x == MatMult(x) (the FAST-rule matrix is
    the matrix-free Jacobian)."""
    mesh = lib.BoxMesh((2, 2, 2) if deg <= 4 else (1, 2, 1), deg, dirichlet=lib.WALLS | lib.BOTTOM, **(morph or {}))
    o = ol.OracleElast(mesh, E0=1.0, nu=0.3, exact=0, jac_qmethod=ol.FAST)
    rule = lib.Rule.preferred(ctx, deg)
    efs = lib.EFS(ctx, rule, deg)
    space = lib.Space.from_mesh(ctx, mesh, bs=3)
    lam, mu = 0.3 / (1.3 * 0.4), 1.0 / 2.6
    op = lib.Op(ctx, lib.OP_ELAST, lib.ElastParam(lam, mu), rule, [space], [efs], mesh.coords)
    op.set_forcing(o.forcing(op.quadrature_points()))
    op.set_dirichlet(0, o.dirichlet_cache())
    rowptr, colidx = op.dense_pattern(0, mesh.l2g)
    rp, ci = dense_pattern(mesh)
    assert np.array_equal(rowptr, rp) and np.array_equal(colidx, ci)
    x = 0.05 * np.random.default_rng(deg).standard_normal(mesh.n * 3)
    vals = op.assemble_dense(lib.DENSE_ELAST_FAST, x, mesh.l2g, rowptr, colidx).reshape(-1, 3, 3)
    J = sp.bsr_matrix((vals, colidx, rowptr), shape=(mesh.n * 3, mesh.n * 3)).tocsr()
    J.sort_indices()
    ref = sp.csr_matrix((o.jacobian_q1(x, J.indptr.astype(np.int32), J.indices.astype(np.int32)), J.indices, J.indptr), shape=J.shape)
    assert np.abs((J - ref).data).max() <= TOL * np.abs(ref.data).max()
    op.function(x)
    v = np.random.default_rng(7).standard_normal(mesh.n * 3)
    Jv = op.mult(v)
    if deg > 1:           # degree 1 takes the Q1 path with the reference's deriv1 = +-1 quirk (efstopo.c:668)
        assert np.abs(J @ v - Jv).max() <= 1e-11 * np.abs(Jv).max()
    op.close()


def test_dense_column_path_equals_shared_memory_path(ctx, monkeypatch):
    """The two implementations of the element-dense assembly (element matrix in shared memory / column by column) agree."""
    mesh = lib.BoxMesh((2, 2, 2), 3, dirichlet=lib.WALLS, morph=0.4, twist=0.2)
    prm = dict(p=1.7, lam=0.3, eps=1.0)
    o = ol.OracleEllip(mesh, exact=0, jac_qmethod=ol.FAST, **prm)
    op = make_ellip(ctx, mesh, o, **prm)
    rowptr, colidx = op.dense_pattern(0, mesh.l2g)
    x = 0.3 * np.random.default_rng(3).standard_normal(mesh.n)
    a = op.assemble_dense(lib.DENSE_ELLIP_FAST, x, mesh.l2g, rowptr, colidx)
    monkeypatch.setenv("DOHP_B200_DENSE_COLUMNS", "1")
    b = op.assemble_dense(lib.DENSE_ELLIP_FAST, x, mesh.l2g, rowptr, colidx)
    assert relerr(b, a) <= 1e-13
    op.close()


@pytest.mark.parametrize("deg", [2, 3, 4, 5, 6, 7])                # pressure degree deg - 1; velocity degree >= 6: column path
@pytest.mark.parametrize("morph", [None, dict(morph=0.3, twist=0.2)])
def test_stokes_pressure_blocks_match_oracle(ctx, deg, morph):
    """StokesJacobianAssemble_Pressure (stokes.c:739-794): D (pressure mass scaled by 1/eta) and Daux (Neumann Laplacian) on the
    element-dense pressure pattern, incl. quirk (iii) of dEFSGetExplicitSparse_Basic on the non-affine mesh."""
    ne = (2, 2, 2) if deg <= 5 else (1, 2, 1)
    mu_ = lib.BoxMesh(ne, deg, **(morph or {}))
    mp_ = lib.BoxMesh(ne, deg - 1, dirichlet=0, **(morph or {}))
    prm = dict(B=1.3, eps=0.1, p=1.2, nu=0.0)
    o = ol.OracleStokes(mu_, mp_, exact=1, a=0.9, b=1.1, c=0.8, scale=0.5, **prm)
    rule = lib.Rule.preferred(ctx, deg)
    efsu, efsp = lib.EFS(ctx, rule, deg), lib.EFS(ctx, rule, deg - 1)
    su, sp_ = lib.Space.from_mesh(ctx, mu_, bs=3), lib.Space.from_mesh(ctx, mp_, bs=1)
    op = lib.Op(ctx, lib.OP_STOKES, lib.StokesParam(prm["B"], prm["eps"], prm["p"], prm["nu"]), rule, [su, sp_], [efsu, efsp], mu_.coords)
    op.set_dirichlet(0, o.dirichlet_cache())
    rowptr, colidx = op.dense_pattern(1, mp_.l2g)
    rp, ci = dense_pattern(mp_)
    assert np.array_equal(rowptr, rp) and np.array_equal(colidx, ci)
    xu = 0.2 * np.random.default_rng(deg).standard_normal(mu_.n * 3)
    D = op.assemble_dense(lib.DENSE_STOKES_D, [xu, np.zeros(mp_.n)], mp_.l2g, rowptr, colidx)
    Da = op.assemble_dense(lib.DENSE_STOKES_DAUX, [xu, np.zeros(mp_.n)], mp_.l2g, rowptr, colidx)
    rD, rDa = o.jacobian_pressure(xu, rp, ci)
    assert relerr(D, rD) <= TOL
    assert relerr(Da, rDa) <= TOL
    op.close()


def test_golden_qfast_ksp_norms(ctx):
    """ellip-morph-qfast.refout (src/fs/tests/CMakeLists.txt:58; -morph 0.5 -twist 0.1 -ellip_p 1.9 -exact 3 -ksp_type richardson
    -pc_type lu -ellip_jac_qmethod fast):  '0 KSP Residual norm 172.603' and '1 KSP Residual norm < 1.e-11': the FAST-assembled
    matrix is the Jacobian itself, so one Richardson step with its LU solves the matrix-free system."""
    mesh = lib.BoxMesh((4, 4, 4), 3, morph=0.5, twist=0.1)
    o = ol.OracleEllip(mesh, p=1.9, exact=3, jac_qmethod=ol.FAST)
    op = make_ellip(ctx, mesh, o, p=1.9)
    x0 = np.zeros(mesh.n)
    F = op.function(x0)
    rowptr, colidx = op.dense_pattern(0, mesh.l2g)
    vals = op.assemble_dense(lib.DENSE_ELLIP_FAST, x0, mesh.l2g, rowptr, colidx)
    lu = spla.splu(sp.csr_matrix((vals, colidx, rowptr), shape=(mesh.n, mesh.n)).tocsc())
    y = lu.solve(F)
    assert _fmt6(np.linalg.norm(y)) == 172.603
    r1 = lu.solve(F - op.mult(y))                     # preconditioned residual after one Richardson step
    assert np.linalg.norm(r1) < 1e-11 * np.linalg.norm(y)
    op.close()
