"""Value parity against the oracle at MID sizes (hundreds to thousands of elements, element counts that are not multiples of the
elements-per-block of any kernel, grids of more than one wave of blocks on 148 SMs): the small cases of test_gpu_ellip / _elast_stokes /
_q1 fit in a handful of blocks, the BASELINE-size cases are checked through properties only; this file closes the gap with cases the
oracle still finishes in about a second each.  Through the C ABI, bar 1e-12 relative (FP64)."""
import os

import numpy as np
import pytest

import oracle_lib as ol
from dohp_b200 import lib
from dohp_b200.box import q1_pattern
from test_gpu_elast_stokes import make_elast, make_stokes
from test_gpu_q1 import make_ellip

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def ctx():
    import torch
    ol.build_oracle(ref=False)
    ol.oracle().or_set_num_threads(os.cpu_count() or 1)
    c = lib.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    yield c
    c.close()


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("ne,deg,morph", [((13, 11, 9), 4, dict(morph=0.3, twist=0.25)), ((9, 8, 7), 7, None), ((7, 5, 3), 9, dict(morph=0.2, twist=0.1))])
def test_ellip_function_mult_q1_midsize(ctx, ne, deg, morph):
    mesh = lib.BoxMesh(ne, deg, **(morph or {}))
    prm = dict(p=1.6, lam=0.7, eps=0.9)
    o = ol.OracleEllip(mesh, exact=0, a=1.1, b=0.9, c=1.3, **prm)
    op = make_ellip(ctx, mesh, o, **prm)
    rng = np.random.default_rng(100 + deg)
    x = 0.3 * rng.standard_normal(mesh.n)
    assert relerr(op.function(x), o.function(x)) <= TOL
    dx = rng.standard_normal(mesh.n)
    assert relerr(op.mult(dx), o.mult(dx)) <= TOL                      # p-Laplace Jacobian with the stash F(x) left
    if deg <= 4:
        rowptr, colidx = op.q1_pattern(mesh.l2g)
        rp, ci = q1_pattern(mesh)
        assert np.array_equal(rowptr, rp) and np.array_equal(colidx, ci)
        assert relerr(op.assemble_q1(x, mesh.l2g, rowptr, colidx), o.jacobian_q1(x, rp, ci)) <= TOL
    op.close()


@pytest.mark.parametrize("deg", [2, 5, 8])
def test_linear_poisson_midsize(ctx, deg):
    """The headline kernel (ellip_p = 2: no stash) on an affine and on a morphed box of 11 x 9 x 7 = 693 elements."""
    for morph in (None, dict(morph=0.3, twist=0.12)):
        mesh = lib.BoxMesh((11, 9, 7), deg, **(morph or {}))
        o = ol.OracleEllip(mesh, exact=0)
        op = make_ellip(ctx, mesh, o)
        x = np.random.default_rng(deg).standard_normal(mesh.n)
        o.function(np.zeros(mesh.n))
        assert relerr(op.mult(x), o.mult(x)) <= TOL
        op.close()


def test_elast_midsize(ctx):
    mesh = lib.BoxMesh((9, 7, 5), 3, morph=0.25, twist=0.2)
    o = ol.OracleElast(mesh, exact=0, a=0.6, b=0.5, c=0.4, scale=0.2)
    op = make_elast(ctx, mesh, o)
    rng = np.random.default_rng(3)
    x = 0.05 * rng.standard_normal(mesh.n * 3)
    assert relerr(op.function(x), o.function(x)) <= TOL
    dx = rng.standard_normal(mesh.n * 3)
    assert relerr(op.mult(dx), o.mult(dx)) <= TOL
    import scipy.sparse as sp
    rowptr, colidx = op.q1_pattern(mesh.l2g)
    vals = op.assemble_q1(x, mesh.l2g, rowptr, colidx)
    J = sp.bsr_matrix((vals.reshape(-1, 3, 3), colidx, rowptr), shape=(mesh.n * 3, mesh.n * 3)).tocsr()
    rp, ci = q1_pattern(mesh, bs=3)
    ref = sp.csr_matrix((o.jacobian_q1(x, rp, ci), ci, rp), shape=J.shape)
    d = (J - ref).tocoo()
    assert (np.abs(d.data).max() if d.nnz else 0.0) <= TOL * np.abs(ref.data).max()
    op.close()


def test_stokes_midsize(ctx):
    ne, kw = (7, 6, 5), dict(morph=0.2, twist=0.15)
    mu, mp = lib.BoxMesh(ne, 3, **kw), lib.BoxMesh(ne, 2, dirichlet=0, **kw)
    o = ol.OracleStokes(mu, mp, exact=1, p=1.2, eps=0.1)
    op = make_stokes(ctx, mu, mp, o)
    rng = np.random.default_rng(4)
    xu, xp = 0.1 * rng.standard_normal(mu.n * 3), 0.1 * rng.standard_normal(mp.n)
    Fu, Fp = op.function([xu, xp])
    Ou, Op_ = o.function(xu, xp)
    assert relerr(Fu, Ou) <= TOL and relerr(Fp, Op_) <= TOL
    du, dp = rng.standard_normal(mu.n * 3), rng.standard_normal(mp.n)
    Ju, Jp = op.mult([du, dp])
    Ru, Rp = o.mult(du, dp)
    assert relerr(Ju, Ru) <= TOL and relerr(Jp, Rp) <= TOL
    op.close()
