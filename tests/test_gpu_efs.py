"""dEFS-level parity on the GPU: dB200EFSApplyBatch / dB200RuleComputePhysicalBatch (through the C ABI) against the
oracle's restatement of dEFSApply_Tensor_Hex (efstopo.c:368-423) and dRuleComputePhysical (jacobi.c:383-414) on the
same seeded inputs.  Bar: max|gpu - oracle| <= 1e-12 * max|oracle| (FP64; only the summation order differs)."""
import numpy as np
import pytest

import oracle_lib as ol
from dohp_b200 import lib

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def ctx():
    ol.build_oracle(ref=False)
    import torch
    c = lib.Context(0, stream=torch.cuda.current_stream().cuda_stream)     # same stream as the torch tensors the tests pass in
    yield c
    c.close()


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def _geometry(rng, nelem, cefs_o):
    base = np.array([[i, j, k] for i in (-1., 1.) for j in (-1., 1.) for k in (-1., 1.)])
    coords = base[None] * (0.3 + 0.2 * rng.random((nelem, 1, 3))) + 0.04 * rng.standard_normal((nelem, 8, 3))
    jac = cefs_o.apply(None, 3, coords, ol.APPLY_GRAD)          # [nelem][Q][3][3] = d x_d / d xi_e
    return coords, jac.reshape(nelem, -1, 9)


# (basis degree, rule order): FAST rule of the space itself (P=Q), the Stokes pressure space on the velocity rule
# (P=Q-1) and the Q1 coordinate space on a high-order rule (P=2)
PAIRS = [(p, 2 * p) for p in range(1, 11)] + [(p - 1, 2 * p) for p in range(2, 11)] + [(1, 2 * p) for p in range(3, 11)]


@pytest.mark.parametrize("deg,rorder", PAIRS)
@pytest.mark.parametrize("D", [1, 3])
def test_apply_all_modes(ctx, deg, rorder, D):
    rng = np.random.default_rng(1000 * deg + rorder + D)
    nelem = 7
    # the checker is the reference's OWN dEFSApply / dRuleComputePhysical (oracle/_ref, compiled unmodified) wherever it was built;
    # the plain-C restatement (validated against it by tests/test_oracle_jacobi.py) otherwise
    EFS = ol.RefEFS if ol.have_ref() else ol.OracleEFS
    o = EFS(deg, rorder=rorder)
    oc = EFS(1, rorder=rorder)
    rule = lib.Rule(ctx, *lib.rule_1d(lib.FAST, lib.GAUSS, rorder))
    efs = lib.EFS(ctx, rule, deg)
    assert efs.P == o.P[0] and rule.Q == o.Q[0]
    _, jac = _geometry(rng, nelem, oc)
    jinv, jw = o.compute_physical(jac)
    gjinv, gjw = rule.compute_physical(jac)
    assert relerr(gjinv, jinv) <= TOL and relerr(gjw, jw) <= TOL
    x = rng.standard_normal((nelem, o.np, D))
    for amode, ji in ((lib.APPLY_INTERP, None), (lib.APPLY_GRAD, jinv), (lib.APPLY_GRAD, None)):
        ref = o.apply(ji, D, x, amode)
        got = efs.apply(ji, D, x, amode)
        assert relerr(got, ref) <= TOL, (amode, relerr(got, ref))
    v = rng.standard_normal((nelem, o.nq, D))
    dv = rng.standard_normal((nelem, o.nq, D, 3))
    ref = o.apply(jinv, D, v, ol.APPLY_INTERP_TRANSPOSE)
    got = efs.apply(jinv, D, v, lib.APPLY_INTERP_TRANSPOSE)
    assert relerr(got, ref) <= TOL
    # CommitPatchApplied: INTERP_TRANSPOSE INSERT then GRAD_TRANSPOSE ADD into the same ey (fsrulesetit.c:631-637)
    ref = o.apply(jinv, D, dv, ol.APPLY_GRAD_TRANSPOSE, ol.ADD_VALUES, out=ref)
    got = efs.apply(jinv, D, dv, lib.APPLY_GRAD_TRANSPOSE, lib.ADD_VALUES, out=got)
    assert relerr(got, ref) <= TOL
    ref = o.apply(jinv, D, dv, ol.APPLY_GRAD_TRANSPOSE)
    got = efs.apply(jinv, D, dv, lib.APPLY_GRAD_TRANSPOSE)
    assert relerr(got, ref) <= TOL


def test_apply_errors_like_the_reference(ctx):
    rule = lib.Rule(ctx, *lib.rule_1d(lib.FAST, lib.GAUSS, 4))
    efs = lib.EFS(ctx, rule, 2)
    x = np.zeros((1, 27, 1))
    with pytest.raises(lib.DohpError):          # dofs < 1, jacobi.c:495
        efs.apply(None, 0, x, lib.APPLY_INTERP, nelem=1, out=np.zeros(27))
    with pytest.raises(lib.DohpError):          # SYMGRAD is declared but unimplemented upstream, efstopo.c:419-420
        efs.apply(None, 1, x, 40005, nelem=1, out=np.zeros(27))
    with pytest.raises(lib.DohpError):          # GRAD_TRANSPOSE without jinv, efstopo.c:600
        efs.apply(None, 1, np.zeros((1, 27, 1, 3)), lib.APPLY_GRAD_TRANSPOSE, nelem=1)
    jac = np.tile(-np.eye(3).ravel(), (1, 27, 1))
    with pytest.raises(lib.DohpError, match="Negative Jacobian"):     # jacobi.c:397
        rule.compute_physical(jac)


def test_device_pointers_and_empty_batch(ctx):
    import torch
    rule = lib.Rule(ctx, *lib.rule_1d(lib.FAST, lib.GAUSS, 6))
    efs = lib.EFS(ctx, rule, 3)
    o = ol.OracleEFS(3)
    x = np.random.default_rng(5).standard_normal((11, 64, 2))
    ref = o.apply(None, 2, x, ol.APPLY_INTERP)
    xd = torch.from_numpy(x).cuda()
    od = torch.zeros(11, 64, 2, dtype=torch.float64, device="cuda")
    efs.apply(None, 2, xd, lib.APPLY_INTERP, out=od, nelem=11)
    ctx.synchronize()
    assert relerr(od.cpu().numpy(), ref) <= TOL
    out = efs.apply(None, 2, np.zeros((0, 64, 2)), lib.APPLY_INTERP, nelem=0, out=np.zeros((0, 64, 2)))
    assert out.size == 0


# ---- any dEFS dJacobiGetEFS_Tensor can produce (tensor.c:503-545), against the REFERENCE's own dEFSApply (oracle/_ref) -----------------
# (basis degrees, rule orders, method): anisotropic degree, the (bdeg j, rdeg 2i) pairs of src/jacobi/tests/ex1.c:285-286 incl. P > Q,
# SPARSE rules (Q = 2p, tensorquad.c:120-169), and a size the templated kernels cover (cross-check of the two GPU paths)
GENERIC = [((3, 2, 5), (6, 4, 10), lib.FAST), ((1, 4, 2), (8, 8, 8), lib.FAST), ((5, 5, 5), (4, 4, 4), lib.FAST), ((2, 2, 2), (10, 10, 10), lib.FAST),
           ((6, 3, 4), (2, 12, 6), lib.FAST), ((3, 3, 3), (6, 6, 6), lib.SPARSE), ((4, 2, 3), (8, 4, 6), lib.SPARSE), ((10, 10, 10), (20, 20, 20), lib.FAST),
           ((4, 4, 4), (8, 8, 8), lib.FAST)]


@pytest.mark.parametrize("bdeg,rorder,method", GENERIC)
@pytest.mark.parametrize("D", [1, 2, 3])
def test_apply_any_efs_matches_the_reference_itself(ctx, bdeg, rorder, method, D):
    if not ol.have_ref():
        pytest.skip("oracle/_ref (the reference's dJacobi layer compiled unmodified) is not built")
    rng = np.random.default_rng(sum(bdeg) * 100 + sum(rorder) + D)
    nelem = 3
    r = ol.RefEFS(bdeg, rorder=rorder, method=method)
    rc = ol.RefEFS(1, rorder=rorder, method=method)
    rule = lib.Rule(ctx, [r.rule(d)[0] for d in range(3)], [r.rule(d)[1] for d in range(3)])
    efs = lib.EFS(ctx, rule, bdeg, interp=[r.tables(d)[0] for d in range(3)], deriv=[r.tables(d)[1] for d in range(3)])
    assert efs.P3 == r.P and rule.Q3 == r.Q
    # the library's own table arithmetic (dB200Basis1D) reproduces the reference's tables in every direction
    own = lib.EFS(ctx, rule, bdeg)
    for d in range(3):
        assert np.abs(own.interps[d] - r.tables(d)[0]).max() <= 2e-14 and np.abs(own.derivs[d] - r.tables(d)[1]).max() <= 1e-12
    base = np.array([[i, j, k] for i in (-1., 1.) for j in (-1., 1.) for k in (-1., 1.)])
    coords = base[None] * (0.3 + 0.2 * rng.random((nelem, 1, 3))) + 0.04 * rng.standard_normal((nelem, 8, 3))
    jac = rc.apply(None, 3, coords, ol.APPLY_GRAD).reshape(nelem, -1, 9)
    jinv, jw = r.compute_physical(jac)
    gjinv, gjw = rule.compute_physical(jac)
    assert relerr(gjinv, jinv) <= TOL and relerr(gjw, jw) <= TOL
    x = rng.standard_normal((nelem, r.np, D))
    for amode, ji in ((lib.APPLY_INTERP, None), (lib.APPLY_GRAD, jinv), (lib.APPLY_GRAD, None)):
        assert relerr(efs.apply(ji, D, x, amode), r.apply(ji, D, x, amode)) <= TOL, amode
    # ADD_VALUES into a non-zero output
    o0 = rng.standard_normal((nelem, r.nq, D, 3))
    assert relerr(efs.apply(jinv, D, x, lib.APPLY_GRAD, lib.ADD_VALUES, out=o0.copy()), r.apply(jinv, D, x, ol.APPLY_GRAD, ol.ADD_VALUES, out=o0.copy())) <= TOL
    v = rng.standard_normal((nelem, r.nq, D))
    dv = rng.standard_normal((nelem, r.nq, D, 3))
    ref = r.apply(jinv, D, v, ol.APPLY_INTERP_TRANSPOSE)
    got = efs.apply(jinv, D, v, lib.APPLY_INTERP_TRANSPOSE)
    assert relerr(got, ref) <= TOL
    ref = r.apply(jinv, D, dv, ol.APPLY_GRAD_TRANSPOSE, ol.ADD_VALUES, out=ref)
    got = efs.apply(jinv, D, dv, lib.APPLY_GRAD_TRANSPOSE, lib.ADD_VALUES, out=got)
    assert relerr(got, ref) <= TOL
    assert relerr(efs.apply(jinv, D, dv, lib.APPLY_GRAD_TRANSPOSE), r.apply(jinv, D, dv, ol.APPLY_GRAD_TRANSPOSE)) <= TOL


def test_fused_operators_reject_an_anisotropic_efs_loudly(ctx):
    rule = lib.Rule(ctx, [lib.rule_1d(lib.FAST, lib.GAUSS, o)[0] for o in (4, 6, 4)], [lib.rule_1d(lib.FAST, lib.GAUSS, o)[1] for o in (4, 6, 4)])
    efs = lib.EFS(ctx, rule, (2, 3, 2))
    mesh = lib.BoxMesh((2, 2, 2), 2)
    space = lib.Space.from_mesh(ctx, mesh)
    with pytest.raises(lib.DohpError, match="isotropic"):
        lib.Op(ctx, lib.OP_ELLIP, lib.EllipParam(2.0, 1.0, 0.0), rule, [space], [efs], mesh.coords)
