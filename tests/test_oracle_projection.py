"""The reference's dFS integration test src/fs/tests/ex1.c (L2 projection through three iterator API styles) as known answers for the
oracle: on the two-hex mesh of ex1.c:53-102 the restated residual reproduces the printed `0 SNES Function norm` of five upstream
refouts, the sizes / sums of useFS (ex1.c:196-241), and -- after solving the (linear) projection -- every printed pointwise and
gradient residual norm (ex1.c:484-532).  These are numbering independent.  CPU only."""
import numpy as np
import pytest

import oracle_lib as ol
from dohp_b200 import mesh
from dohp_b200.box import q1_pattern
from projection_common import CASES, cg, parse, printed_equal


@pytest.fixture(scope="module", autouse=True)
def _build():
    ol.build_oracle(ref=False)


def ex1_space(deg, affine=False):
    return mesh.build_hex_space(ol.EX1_VERTS_AFFINE if affine else ol.EX1_VERTS_PARAMETRIC, ol.EX1_HEXES, deg, dirichlet_faces=[])


@pytest.mark.parametrize("name", sorted(CASES))
def test_projection_refout(name):
    deg, freq, sparse = CASES[name]
    gold = parse(name)
    s = ex1_space(deg)
    assert s.nelem * (deg + 1) ** 3 == gold["expanded"] and s.n == gold["global"]
    # useFS: expanded ones -> global (E^T, every node counted once per element) -> expanded
    x = np.ones(gold["expanded"])
    g = np.bincount(s.e2l, weights=x, minlength=s.n)
    y = g[s.e2l]
    assert (x.sum(), g.sum(), y.sum()) == gold["sums"]
    assert y.sum() - g.sum() == 2.0 * (deg + 1) ** 2                # the interface nodes are counted twice more (ex1.c:229-233)
    o = ol.OracleProj(s, frequency=freq, proj_qmethod=ol.SPARSE if sparse else ol.FAST)
    F0 = o.function(np.zeros(s.n))
    assert printed_equal(np.linalg.norm(F0), gold["fnorm0"], "%g")
    u = cg(o.mult, -F0)
    r = o.function(u)
    assert np.abs(r).max() < 1e-13                                  # "Algebraic projection residual"
    res, gres = o.residual_norms(u)
    for v, pr in zip(res, gold["pointwise"]):
        assert printed_equal(v, pr, "%8.2e"), (v, pr)
    for v, pr in zip(gres, gold["gradient"]):
        assert printed_equal(v, pr, "%8.2e"), (v, pr)


def test_projection_of_a_linear_field_is_exact():
    """-exact 1 (ex1.c:48-51): 10x + y + 0.1z is in the space on the trilinear elements, so the projection reproduces it to rounding"""
    s = ex1_space(3)
    o = ol.OracleProj(s, exact=1)
    u = cg(o.mult, -o.function(np.zeros(s.n)))
    assert np.abs(u - o.target(s.node_xyz[:s.n])).max() < 1e-12 * np.abs(u).max() * 10        # conditioning of the mass matrix
    res, gres = o.residual_norms(u)
    assert res.max() < 1e-10 and gres.max() < 1e-18


def test_q1_mass_matrix_properties():
    """ProjJacobian1 (ex1.c:432-482): the Q1 mass matrix on the SPARSE patches is symmetric and its entries sum to the volume"""
    import scipy.sparse as sp
    for affine, vol in ((True, 2.0), (False, None)):
        s = ex1_space(4, affine)
        rp, ci = q1_pattern(s)
        o = ol.OracleProj(s)
        M = sp.csr_matrix((o.jacobian_q1(rp, ci), ci, rp), shape=(s.n, s.n))
        assert abs(M - M.T).max() < 1e-15
        if vol is not None:
            assert abs(M.sum() - vol) < 1e-13
        else:
            assert abs(M.sum() - o.mult(np.ones(s.n)).sum()) < 1e-13        # both integrate 1 over the domain
