"""SURVEY.md 8f-3, CPU: the numpy restatement of dEFSApply on dTOPO_LINE / dTOPO_QUAD (tests/oracle_lib.py:lowdim_apply) against the
reference's OWN dEFSApply_Tensor_Line / _Quad (oracle/_ref, compiled unmodified), and the facet-rule points of the C ABI (host
arithmetic) against the transform table of dQuadratureGetFacetRules_Tensor_Private.  Upstream implements neither mapped gradients nor
GRAD_TRANSPOSE below dimension 3: both must error."""
import numpy as np
import pytest

import oracle_lib as ol

CASES = [((1,), 1), ((3,), 2), ((6,), 1), ((9,), 3), ((2, 2), 1), ((3, 5), 2), ((4, 1), 3), ((7, 7), 1)]


@pytest.mark.parametrize("degs,D", CASES)
def test_restatement_matches_reference_line_quad(degs, D):
    ol.build_oracle()
    if not ol.have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    rng = np.random.default_rng(sum(degs) * 10 + D)
    npts = int(np.prod([d + 1 for d in degs]))
    x = rng.standard_normal((5, npts, D))
    for amode in (ol.APPLY_INTERP, ol.APPLY_GRAD):
        err, r, P, Q = ol.ref_lowdim_apply(degs, x, amode, D)
        assert err == 0 and P == [d + 1 for d in degs]
        o = ol.lowdim_apply(degs, x, amode, D)
        if amode == ol.APPLY_GRAD:                          # directions >= dim: out-of-bounds reads upstream (TensorRuleNoMapping), zero here
            assert np.all(o[..., len(degs):] == 0)
            o, r = o[..., :len(degs)], r[..., :len(degs)]
        assert o.shape == r.shape and np.abs(o - r).max() <= 1e-13 * np.abs(r).max()
    nq = int(np.prod(Q))
    v = rng.standard_normal((5, nq, D))
    err, r, _, _ = ol.ref_lowdim_apply(degs, v, ol.APPLY_INTERP_TRANSPOSE, D)
    o = ol.lowdim_apply(degs, v, ol.APPLY_INTERP_TRANSPOSE, D)
    assert err == 0 and np.abs(o - r).max() <= 1e-13 * np.abs(r).max()
    # dAPPLY_GRAD_TRANSPOSE is an invalid mode on Line / Quad upstream (efstopo.c:321-325, 361-363)
    err, _, _, _ = ol.ref_lowdim_apply(degs, rng.standard_normal((5, nq, D, 3)), ol.APPLY_GRAD_TRANSPOSE, D)
    assert err != 0
    with pytest.raises(ValueError):
        ol.lowdim_apply(degs, v, ol.APPLY_GRAD_TRANSPOSE, D)


@pytest.mark.parametrize("facet", range(6))
def test_facet_rule_nodes_host(facet):
    from dohp_b200 import lib
    lib.build() if not __import__("os").path.exists(lib.LIBPATH) else None
    ol.build_oracle(ref=False)
    c0, w0 = ol.rule_1d(ol.FAST, ol.GAUSS, 6)
    c1, w1 = ol.rule_1d(ol.FAST, ol.GAUSS, 4)
    nodes, wts = lib.facet_rule_nodes(facet, [c0, c1], [w0, w1])
    rn, rw = ol.facet_rule_nodes(facet, [c0, c1], [w0, w1])
    assert np.array_equal(nodes, rn) and np.allclose(wts, rw, rtol=1e-15, atol=0)
    # every point lies on the facet: one coordinate is +-1, the weights integrate the 2x2 reference square
    fixed = np.where(np.all(np.abs(np.abs(nodes) - 1.0) < 1e-15, axis=0))[0]
    assert len(fixed) == 1 and abs(wts.sum() - 4.0) < 1e-13
    assert np.allclose(nodes[:, fixed[0]], ol.FACET_TRANSLATION[facet][fixed[0]])
