"""The oracle restatement (oracle/oracle.c) against the reference's own dJacobi layer (oracle/_ref) and the
reference's golden refouts.  CPU only.  Tests that need /root/reference or oracle/_ref skip when it is absent
(the GPU box has neither the reference tree; _ref travels prebuilt)."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol

needs_reference = pytest.mark.skipif(not os.path.isdir(ol.REFERENCE), reason="/root/reference not mounted")


@pytest.fixture(scope="module", autouse=True)
def _build():
    ol.build_oracle()


GOLDEN_RUNS = [
    ("jacobi-1", "jacobi_ex1", "-min_rdeg 10 -max_rdeg 11 -max_bdeg 10"),
    ("jacobi-1-lobatto", "jacobi_ex1", "-min_rdeg 9 -max_rdeg 11 -max_bdeg 10 -dquad_tensor_gauss_family lobatto"),
    ("jacobi-3-tensor-gauss", "jacobi_ex3", "-djac_type tensor -rule_degree 8 -dquad_tensor_gauss_family gauss"),
    ("jacobi-3-tensor-lobatto", "jacobi_ex3", "-djac_type tensor -rule_degree 6 -dquad_tensor_gauss_family lobatto"),
    ("jacobi-3-tensor-sparse-gauss", "jacobi_ex3", "-djac_type tensor -rule_degree 4 -dquad_tensor_gauss_family gauss -qmethod sparse -just_view"),
    ("jacobi-3-tensor-sparsex-r4", "jacobi_ex3", "-djac_type tensor -rule_degree 4 -dquad_tensor_gauss_family gauss -qmethod sparse -sparse"),
    ("jacobi-3-tensor-sparsex-r6", "jacobi_ex3", "-djac_type tensor -rule_degree 6 -dquad_tensor_gauss_family gauss -qmethod sparse -sparse"),
    ("jacobi-4-tensor-sparse", "jacobi_ex4", "-quadrature_method sparse"),
    ("jacobi-4-tensor-fast", "jacobi_ex4", "-quadrature_method fast"),
]


@needs_reference
@pytest.mark.parametrize("name,exe,args", GOLDEN_RUNS, ids=[g[0] for g in GOLDEN_RUNS])
def test_ref_build_reproduces_upstream_golden_output(name, exe, args):
    """oracle/_ref (reference sources compiled unmodified) must reproduce src/jacobi/tests/refout/*.refout byte for
    byte: this is what pins the oracle (src/jacobi/tests/CMakeLists.txt:5-17)."""
    out = subprocess.run([os.path.join(ol.ORACLE_DIR, "_ref", exe)] + args.split(), capture_output=True, check=True).stdout
    with open(os.path.join(ol.REFERENCE, "src/jacobi/tests/refout", name + ".refout"), "rb") as f:
        golden = f.read()
    assert out == golden


needs_ref = pytest.mark.skipif(not ol.have_ref() and not os.path.isdir(ol.REFERENCE), reason="oracle/_ref not built")

CASES = [(p, ol.FAST, ol.GAUSS) for p in range(1, 11)] + [(p, ol.SPARSE, ol.GAUSS) for p in range(1, 8)] + \
        [(p, ol.SELF, ol.LOBATTO) for p in (1, 3, 6)] + [(p, ol.FAST, ol.LOBATTO) for p in (2, 5)]


@needs_ref
@pytest.mark.parametrize("p,method,family", CASES)
def test_tables_match_reference(p, method, family):
    o, r = ol.OracleEFS(p, method=method, family=family), ol.RefEFS(p, method=method, family=family)
    assert o.P == r.P and o.Q == r.Q
    for d in range(3):
        for a, b, name in zip(o.tables(d), r.tables(d), ("interp", "deriv", "node", "weight")):
            assert np.allclose(a, b, rtol=0, atol=2e-14 * max(1.0, np.abs(b).max())), (name, np.abs(a - b).max())
        for a, b in zip(o.rule(d), r.rule(d)):
            assert np.allclose(a, b, rtol=0, atol=1e-15)
    on, ops, oind, ow = o.patches()
    rn, rps, rind, rw = r.patches()
    assert (on, ops) == (rn, rps) and np.array_equal(oind, rind)
    assert np.allclose(ow, rw, rtol=1e-12, atol=0)      # includes quirk (iv) (tensorquad.c:161)


def _random_geometry(rng, nelem, efs_coord):
    """Mildly distorted unit-ish hexes -> jac (GRAD of coords with NULL jinv) at the rule's points."""
    base = np.array([[i, j, k] for i in (-1., 1.) for j in (-1., 1.) for k in (-1., 1.)])
    coords = base[None] * (0.3 + 0.2 * rng.random((nelem, 1, 3))) + 0.05 * rng.standard_normal((nelem, 8, 3))
    return coords


@needs_ref
@pytest.mark.parametrize("p,D", [(1, 1), (2, 1), (3, 1), (3, 3), (4, 3), (5, 1), (7, 2), (10, 1)])
def test_apply_and_physical_match_reference(p, D):
    rng = np.random.default_rng(100 + p)
    nelem = 3
    o, r = ol.OracleEFS(p), ol.RefEFS(p)
    oc, rc = ol.OracleEFS(1, rorder=2 * p), ol.RefEFS(1, rorder=2 * p)
    coords = _random_geometry(rng, nelem, oc)
    jac_o = oc.apply(None, 3, coords, ol.APPLY_GRAD)
    jac_r = rc.apply(None, 3, coords, ol.APPLY_GRAD)
    assert np.allclose(jac_o, jac_r, rtol=0, atol=1e-14)
    jinv_o, jw_o = o.compute_physical(jac_o.reshape(nelem, -1, 9))
    jinv_r, jw_r = r.compute_physical(jac_r.reshape(nelem, -1, 9))
    assert np.allclose(jinv_o, jinv_r, rtol=1e-12, atol=1e-13) and np.allclose(jw_o, jw_r, rtol=1e-12)
    u = rng.standard_normal((nelem, o.np, D))
    for amode in (ol.APPLY_INTERP, ol.APPLY_GRAD):
        a, b = o.apply(jinv_r, D, u, amode), r.apply(jinv_r, D, u, amode)
        assert np.allclose(a, b, rtol=0, atol=1e-12 * np.abs(b).max())
    v = rng.standard_normal((nelem, o.nq, D))
    dv = rng.standard_normal((nelem, o.nq, D, 3))
    for amode, src in ((ol.APPLY_INTERP_TRANSPOSE, v), (ol.APPLY_GRAD_TRANSPOSE, dv)):
        for imode in (ol.INSERT_VALUES, ol.ADD_VALUES):
            y0 = rng.standard_normal((nelem, o.np, D))
            a = o.apply(jinv_r, D, src, amode, imode, out=y0.copy())
            b = r.apply(jinv_r, D, src, amode, imode, out=y0.copy())
            assert np.allclose(a, b, rtol=0, atol=1e-12 * np.abs(b).max())


@needs_ref
def test_p3_sse_kernel_and_nounroll_agree():
    """P=Q=4, D=1 takes the reference's hand-written SSE3 kernel (inlinetmulthex.h:111-400); the restatement follows
    the generic loop nest.  They must agree to rounding."""
    rng = np.random.default_rng(7)
    o, r = ol.OracleEFS(3), ol.RefEFS(3)
    u = rng.standard_normal((5, 64, 1))
    for amode in (ol.APPLY_INTERP, ol.APPLY_GRAD):
        a, b = o.apply(None, 1, u, amode), r.apply(None, 1, u, amode)
        assert np.allclose(a, b, rtol=0, atol=1e-13 * np.abs(b).max())


@needs_ref
@pytest.mark.parametrize("p", [1, 2, 3, 5])
@pytest.mark.parametrize("with_jinv", [False, True])
def test_explicit_sparse_matches_reference(p, with_jinv):
    rng = np.random.default_rng(p)
    o, r = ol.OracleEFS(p, method=ol.SPARSE), ol.RefEFS(p, method=ol.SPARSE)
    cj = (np.eye(3).reshape(9)[None] + 0.1 * rng.standard_normal((o.nq, 9))).copy() if with_jinv else None
    Po, eo, io, do = o.explicit_sparse(cj, eoffset=17)
    Pr, er, ir, dr = r.explicit_sparse(cj, eoffset=17)
    assert Po == Pr == 8 and np.array_equal(eo, er)
    assert np.array_equal(io, ir) and np.allclose(do, dr, rtol=1e-15, atol=1e-16)


@needs_ref
def test_explicit_sparse_basic_fallback_matches_reference():
    """Non-Q1-aligned (FAST rule, one patch): dEFSGetExplicitSparse_Basic incl. its cjinv permutation (jacobi.c:619-625)."""
    rng = np.random.default_rng(3)
    o, r = ol.OracleEFS(2), ol.RefEFS(2)
    cj = (np.eye(3).reshape(9)[None] + 0.1 * rng.standard_normal((o.nq, 9))).copy()
    Po, eo, io, do = o.explicit_sparse(cj, eoffset=5)
    Pr, er, ir, dr = r.explicit_sparse(cj, eoffset=5)
    assert Po == Pr == 27 and np.array_equal(eo, er)
    assert np.allclose(io, ir, rtol=0, atol=1e-14) and np.allclose(do, dr, rtol=0, atol=1e-13)


def test_apply_rejects_bad_arguments():
    """dEFSApply errors: dofs<1 (jacobi.c:495), bad mode (efstopo.c:419-420), GRAD_TRANSPOSE without jinv (efstopo.c:600)."""
    o = ol.OracleEFS(2)
    lib = ol.oracle()
    x, y = np.zeros(27 * 3), np.zeros(27 * 9)
    assert lib.or_efs_apply(o.h, None, 0, ol.dptr(x), ol.dptr(y), ol.APPLY_INTERP, ol.INSERT_VALUES) != 0
    assert lib.or_efs_apply(o.h, None, 1, ol.dptr(x), ol.dptr(y), 40005, ol.INSERT_VALUES) != 0
    assert lib.or_efs_apply(o.h, None, 1, ol.dptr(y), ol.dptr(x), ol.APPLY_GRAD_TRANSPOSE, ol.INSERT_VALUES) != 0


def test_negative_jacobian_is_an_error():
    o = ol.OracleEFS(2)
    jac = np.tile(np.array([1., 0, 0, 0, 1, 0, 0, 0, -1]), (1, o.nq, 1))
    with pytest.raises(RuntimeError):
        o.compute_physical(jac)
