"""dB200_OP_PROJECTION on the GPU (through the C ABI): the L2 projection of src/fs/tests/ex1.c.  Parity 1e-12 against the oracle's
restatement on the reference's own two-hex mesh and on box meshes of every degree; the printed known answers of five upstream refouts
(function norm, pointwise / gradient residual norms after the projection solve driven by the GPU mass operator); the Q1 mass matrix of
ProjJacobian1 (ex1.c:432-482) against the oracle."""
import numpy as np
import pytest

import oracle_lib as ol
from dohp_b200 import lib, mesh
from dohp_b200.box import q1_pattern
from projection_common import CASES, cg, parse, printed_equal

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


def make_proj(ctx, s, o):
    rule = lib.Rule.preferred(ctx, s.p)
    efs = lib.EFS(ctx, rule, s.p)
    space = lib.Space.from_mesh(ctx, s)
    op = lib.Op(ctx, lib.OP_PROJECTION, None, rule, [space], [efs], s.coords)
    op.set_forcing(o.target(op.quadrature_points()))
    return op


@pytest.mark.parametrize("name", sorted(n for n in CASES if not CASES[n][2]))
def test_projection_refout_gpu(ctx, name):
    deg, freq, _ = CASES[name]
    gold = parse(name)
    s = mesh.build_hex_space(ol.EX1_VERTS_PARAMETRIC, ol.EX1_HEXES, deg, dirichlet_faces=[])
    o = ol.OracleProj(s, frequency=freq)
    op = make_proj(ctx, s, o)
    F0 = op.function(np.zeros(s.n))
    assert printed_equal(np.linalg.norm(F0), gold["fnorm0"], "%g")
    assert relerr(F0, o.function(np.zeros(s.n))) <= TOL
    u = cg(lambda v: op.mult(v).copy(), -F0)                        # the projection solve runs on the GPU mass operator
    assert np.abs(op.function(u)).max() < 1e-13
    res, gres = o.residual_norms(u)                                 # the test program's diagnostic (ex1.c:484-532), evaluated by the checker
    for v, pr in zip(res, gold["pointwise"]):
        assert printed_equal(v, pr, "%8.2e"), (v, pr)
    for v, pr in zip(gres, gold["gradient"]):
        assert printed_equal(v, pr, "%8.2e"), (v, pr)
    op.close()


@pytest.mark.parametrize("deg", range(1, 11))
@pytest.mark.parametrize("morph", [None, dict(morph=0.25, twist=0.2)])
def test_projection_function_and_mass_match_oracle(ctx, deg, morph):
    rng = np.random.default_rng(40 + deg)
    m = lib.BoxMesh((3, 2, 2) if deg < 8 else (2, 2, 2), deg, **(morph or {}))
    o = ol.OracleProj(m, frequency=(1.3, 0.7, 1.1))
    op = make_proj(ctx, m, o)
    x = rng.standard_normal(m.n)
    # the box has Dirichlet nodes: the residual sees the committed boundary values (zeros here, as the oracle wrapper passes), the mass product none
    assert relerr(op.function(x), o.function(x)) <= TOL
    assert relerr(op.mult(x), o.mult(x)) <= TOL
    op.set_forcing(None)                                            # no target: the residual is the mass operator
    assert relerr(op.function(x), o.mult(x)) <= TOL
    op.close()


@pytest.mark.parametrize("deg", [1, 2, 4, 7])
def test_projection_q1_mass_matrix_matches_oracle(ctx, deg):
    for s in (mesh.build_hex_space(ol.EX1_VERTS_PARAMETRIC, ol.EX1_HEXES, deg, dirichlet_faces=[]), lib.BoxMesh((2, 3, 2), deg, morph=0.2, twist=0.1)):
        o = ol.OracleProj(s)
        op = make_proj(ctx, s, o)
        rowptr, colidx = op.q1_pattern(s.l2g)
        rp, ci = q1_pattern(s)
        assert np.array_equal(rowptr, rp) and np.array_equal(colidx, ci)
        vals = op.assemble_q1(np.zeros(s.n), s.l2g, rowptr, colidx)
        assert relerr(vals, o.jacobian_q1(rp, ci)) <= TOL
        op.close()


def test_projection_errors_are_loud(ctx):
    s = lib.BoxMesh((2, 2, 2), 2)
    rule = lib.Rule.preferred(ctx, 2)
    efs = lib.EFS(ctx, rule, 2)
    space = lib.Space.from_mesh(ctx, s)
    with pytest.raises(lib.DohpError, match="needs its parameter struct"):
        lib.Op(ctx, lib.OP_ELLIP, None, rule, [space], [efs], s.coords)
    space3 = lib.Space.from_mesh(ctx, s, bs=3)
    with pytest.raises(lib.DohpError, match="one scalar space"):
        lib.Op(ctx, lib.OP_PROJECTION, None, rule, [space3], [efs], s.coords)
    op = lib.Op(ctx, lib.OP_PROJECTION, None, rule, [space], [efs], s.coords)
    with pytest.raises(lib.DohpError, match="does not apply"):           # the element-dense (FAST-rule) assembly is an ellip / elast / stokes path
        rp, ci = op.dense_pattern(0, s.l2g)
        op.assemble_dense(lib.DENSE_ELLIP_FAST, np.zeros(s.n), s.l2g, rp, ci)
    op.close()
