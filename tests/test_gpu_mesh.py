"""SURVEY.md 8(f1) on the GPU: the CUDA operators on a GENERIC hex space (dohp_b200/mesh.py: unstructured element list, every element
in a random one of the 24 cube orientations, natural or RCM entity order) against the oracle on the very same index arrays, and
against the structured-box space node by node.  Q1 pattern on the generic numbering: bit-exact against the numpy union of patches."""
import numpy as np
import pytest

import oracle_lib as ol
from dohp_b200 import lib, mesh
from dohp_b200.box import q1_pattern
from dist_common import field, key

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


def _space(ne, p, ordering, seed, morph=dict(morph=0.3, twist=0.2)):
    verts, hexes = mesh.box_hexes(ne, morph=morph, scramble=np.random.default_rng(seed))
    return mesh.build_hex_space(verts, hexes, p, ordering=ordering)


@pytest.mark.parametrize("p", [1, 2, 3, 5, 8])
@pytest.mark.parametrize("ordering", ["natural", "rcm"])
def test_ellip_on_scrambled_generic_space(ctx, p, ordering):
    s = _space((3, 2, 2), p, ordering, p)
    prm = dict(p=1.6, lam=0.7, eps=0.9)
    o = ol.OracleEllip(s, exact=0, a=1.1, b=0.9, c=1.3, **prm)
    rule = lib.Rule.preferred(ctx, p)
    efs = lib.EFS(ctx, rule, p)
    space = lib.Space.from_mesh(ctx, s)
    op = lib.Op(ctx, lib.OP_ELLIP, lib.EllipParam(prm["p"], prm["eps"], prm["lam"]), rule, [space], [efs], s.coords)
    op.set_forcing(o.forcing(op.quadrature_points()))
    op.set_dirichlet(0, o.dirichlet_cache())
    rng = np.random.default_rng(p)
    x = 0.3 * rng.standard_normal(s.n)
    assert relerr(op.function(x), o.function(x)) <= TOL
    dx = rng.standard_normal(s.n)
    assert relerr(op.mult(dx), o.mult(dx)) <= TOL
    # Q1 rows: pattern bit-exact, values to rounding
    rowptr, colidx = op.q1_pattern(s.l2g)
    rp, ci = q1_pattern(s)
    assert np.array_equal(rowptr, rp) and np.array_equal(colidx, ci)
    assert relerr(op.assemble_q1(x, s.l2g, rowptr, colidx), o.jacobian_q1(x, rp, ci)) <= TOL
    op.close()


@pytest.mark.parametrize("p", [2, 4])
def test_generic_space_equals_box_space_node_by_node(ctx, p):
    """The same physical operator, structured numbering vs scrambled + RCM numbering, matched by node coordinates."""
    ne, morph = (3, 3, 2), dict(morph=0.25, twist=0.15)
    out = []
    for sp_ in (lib.BoxMesh(ne, p, **morph), _space(ne, p, "rcm", 11, morph=morph)):
        rule = lib.Rule.preferred(ctx, p)
        efs = lib.EFS(ctx, rule, p)
        op = lib.Op(ctx, lib.OP_ELLIP, lib.EllipParam(2.0, 1.0, 0.0), rule, [lib.Space.from_mesh(ctx, sp_)], [efs], sp_.coords)
        xyz = np.asarray(sp_.node_xyz)[:sp_.n]
        out.append(dict(zip(key(xyz), op.mult(field(xyz)))))
        op.close()
    a, b = out
    assert set(a) == set(b)
    scale = max(abs(v) for v in a.values())
    assert max(abs(a[k] - b[k]) for k in a) <= TOL * scale


def test_elast_on_scrambled_generic_space(ctx):
    p = 3
    s = _space((2, 2, 2), p, "rcm", 5, morph=dict(morph=0.2, twist=0.2))
    o = ol.OracleElast(s, exact=0, a=0.6, b=0.5, c=0.4, scale=0.2)
    rule = lib.Rule.preferred(ctx, p)
    efs = lib.EFS(ctx, rule, p)
    op = lib.Op(ctx, lib.OP_ELAST, lib.ElastParam(o.lam, o.mu), rule, [lib.Space.from_mesh(ctx, s, bs=3)], [efs], s.coords)
    op.set_forcing(o.forcing(op.quadrature_points()))
    op.set_dirichlet(0, o.dirichlet_cache())
    rng = np.random.default_rng(0)
    x = 0.05 * rng.standard_normal(3 * s.n)
    assert relerr(op.function(x), o.function(x)) <= TOL
    dx = rng.standard_normal(3 * s.n)
    assert relerr(op.mult(dx), o.mult(dx)) <= TOL
    op.close()
