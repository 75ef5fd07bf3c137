"""SURVEY.md 8f-3 on the GPU, through the C ABI: dEFSApply on dTOPO_LINE / dTOPO_QUAD dEFS (dB200RuleCreateDim / dB200EFSCreateDim +
dB200EFSApplyBatch) against the numpy restatement of dEFSApply_Tensor_Line / _Quad (pinned against the reference's own functions by
tests/test_oracle_lowdim.py), upstream's error cases, and hex traces on the facet rules (dB200FacetRuleNodes): values and gradients of
a polynomial field on each of the six facets against the analytic field.  Bar 1e-12 relative."""
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
    c = lib.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    yield c
    c.close()


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("degs", [(1,), (2,), (5,), (10,), (1, 1), (2, 3), (4, 4), (7, 2), (9, 9)])
@pytest.mark.parametrize("D", [1, 3])
def test_line_quad_apply(ctx, degs, D):
    rng = np.random.default_rng(17 * sum(degs) + D)
    rule = lib.RuleDim.preferred(ctx, degs)
    efs = lib.EFSDim(ctx, rule, degs)
    nelem, npts = 6, int(np.prod([d + 1 for d in degs]))
    x = rng.standard_normal((nelem, npts, D))
    for amode in (lib.APPLY_INTERP, lib.APPLY_GRAD):
        ref = ol.lowdim_apply(degs, x, amode, D)
        got = efs.apply(D, x, amode)
        assert got.shape == ref.shape and relerr(got, ref) <= TOL, (amode, relerr(got, ref))
        if amode == lib.APPLY_GRAD:
            assert np.all(got[..., len(degs):] == 0)        # the directions upstream reads out of bounds
    v = rng.standard_normal((nelem, rule.nq, D))
    ref = ol.lowdim_apply(degs, v, lib.APPLY_INTERP_TRANSPOSE, D)
    assert relerr(efs.apply(D, v, lib.APPLY_INTERP_TRANSPOSE), ref) <= TOL
    # ADD_VALUES accumulates
    out = np.ones((nelem, rule.nq, D))
    efs.apply(D, x, lib.APPLY_INTERP, lib.ADD_VALUES, out=out)
    assert relerr(out - 1.0, ol.lowdim_apply(degs, x, lib.APPLY_INTERP, D)) <= 1e-11
    # upstream's error cases (efstopo.c:321-325, 361-363, 638, 653)
    with pytest.raises(lib.DohpError):
        efs.apply(D, rng.standard_normal((nelem, rule.nq, D, 3)), lib.APPLY_GRAD_TRANSPOSE, jinv=np.zeros((nelem, rule.nq, 9)))
    with pytest.raises(lib.DohpError):
        efs.apply(D, x, lib.APPLY_GRAD, jinv=np.zeros((nelem, rule.nq, 9)))


def _field(X):
    x, y, z = X[..., 0], X[..., 1], X[..., 2]
    u = 1.0 + 0.5 * x - 0.25 * y * z + x * x * y + 0.3 * z ** 3 - 0.7 * x * y * z
    g = np.stack([0.5 + 2 * x * y - 0.7 * y * z, -0.25 * z + x * x - 0.7 * x * z, -0.25 * y + 0.9 * z * z - 0.7 * x * y], axis=-1)
    return u, g


@pytest.mark.parametrize("facet", range(6))
@pytest.mark.parametrize("p", [3, 6])
def test_hex_trace_on_facet_rule(ctx, facet, p):
    """A hex dEFS tabulated on the facet rule's points (one point in the normal direction) gives the trace and the full reference
    gradient of the element's field on the facet."""
    c2, w2 = lib.rule_1d(lib.FAST, lib.GAUSS, 2 * p)
    nodes, wts = lib.facet_rule_nodes(facet, [c2, c2], [w2, w2])
    jac = ol.FACET_JAC[facet]
    # hex direction d is driven by quad axis a with sign s (jac[d][a] = s), or fixed at the translation
    coords, weights, order = [], [], []
    for d in range(3):
        a = [k for k in range(2) if jac[d][k] != 0]
        if a:
            coords.append(jac[d][a[0]] * c2); weights.append(w2); order.append(a[0])
        else:
            coords.append(np.array([ol.FACET_TRANSLATION[facet][d]])); weights.append(np.array([1.0])); order.append(-1)
    rule = lib.Rule(ctx, coords, weights)
    efs = lib.EFS(ctx, rule, p)
    gll = efs.node
    Xn = np.array([[a, b, c] for a in gll for b in gll for c in gll])
    u, _ = _field(Xn)
    # hex tensor order of the facet points vs the quad's (i, j) order
    Xq = np.array([[a, b, c] for a in coords[0] for b in coords[1] for c in coords[2]])
    uq, gq = _field(Xq)
    val = efs.apply(None, 1, u.reshape(1, -1, 1), lib.APPLY_INTERP).reshape(-1)
    grad = efs.apply(None, 1, u.reshape(1, -1, 1), lib.APPLY_GRAD).reshape(-1, 3)
    assert relerr(val, uq) <= TOL and relerr(grad, gq) <= TOL
    # the same point set as dB200FacetRuleNodes hands out (possibly in another order), and the facet integral of u by its weights
    key = lambda X: sorted(map(tuple, np.round(X, 13)))
    assert key(Xq) == key(nodes)
    wq = np.array([a * b * c for a in weights[0] for b in weights[1] for c in weights[2]])
    uf, _ = _field(nodes)
    assert abs(wq @ val - wts @ uf) <= 1e-12 * abs(wts @ uf)
