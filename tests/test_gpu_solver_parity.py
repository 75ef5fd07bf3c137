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
from krylov import newton

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
