"""Structured-box space builders (dohp_b200/box.py in numpy, dohp_b200/csrc/boxmesh.cpp in C) against the reference's
golden numbering (tests/golden/bu0.json, extracted from src/fs/tests/refout/bu-0.refout by tests/golden/make_bu0.py)
and against each other, including the block partition and halo lists.  CPU only (no compute calls)."""
import json
import os

import numpy as np
import pytest

from dohp_b200 import box as pybox
from dohp_b200 import lib

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bu0.json")))
SETS = {100: lib.WALLS, 200: lib.TOP, 300: lib.BOTTOM}


@pytest.fixture(scope="module", autouse=True)
def _build():
    lib.build()


def _mask(sets):
    m = 0
    for s in sets:
        m |= SETS[s]
    return m


def _check_golden(e2l, n, nc, node_xyz):
    assert (n, nc, e2l.size) == (GOLD["n"], GOLD["nc"], GOLD["expanded"])
    # dFSGlobalToExpanded(HOMOGENEOUS) of the closure vector 1000+i (bu.c:122-131): explicit nodes show their closure index
    assert np.array_equal(np.where(e2l < n, 1000 + e2l, 0), np.array(GOLD["expanded_homogeneous"]))
    assert np.array_equal(np.where(e2l < n, 1000 + e2l, 10), np.array(GOLD["expanded_inhomogeneous"]))
    # dFSGetNodalCoordinatesGlobal (bu.c:137-139): pins the order of the Dirichlet part of the closure as well
    assert np.allclose(node_xyz[:nc], np.array(GOLD["closure_coordinates"]).reshape(-1, 3), rtol=0, atol=1e-14)


def test_numpy_builder_matches_reference_golden():
    s = pybox.build_box_space(GOLD["ne"], GOLD["p"], dirichlet=GOLD["dirichlet_sets"])
    _check_golden(s.e2l, s.n, s.nc, s.node_xyz)


def test_c_builder_matches_reference_golden():
    m = lib.BoxMesh(GOLD["ne"], GOLD["p"], dirichlet=_mask(GOLD["dirichlet_sets"]))
    _check_golden(m.e2l, m.n, m.nc, m.node_xyz)


CASES = [((3, 2, 4), 1, (100, 200, 300), None), ((2, 3, 2), 2, (100,), None), ((3, 3, 2), 3, (100, 200, 300), dict(morph=0.4, twist=0.3)),
         ((2, 2, 3), 4, (200, 300), None), ((2, 2, 2), 5, (), None)]


@pytest.mark.parametrize("ne,p,sets,morph", CASES)
def test_c_builder_equals_numpy_builder(ne, p, sets, morph):
    s = pybox.build_box_space(ne, p, dirichlet=sets, morph=morph)
    m = lib.BoxMesh(ne, p, dirichlet=_mask(sets), **(morph or {}))
    assert (m.nelem, m.n, m.nc, m.ngh, m.rstart, m.nglobal) == (s.nelem, s.n, s.nc, s.ngh, s.rstart, s.nglobal)
    assert np.array_equal(m.e2l, s.e2l) and np.array_equal(m.l2g, s.l2g)
    assert np.allclose(m.coords, s.coords, rtol=0, atol=1e-15)
    assert np.allclose(m.node_xyz, s.node_xyz, rtol=0, atol=1e-14)


@pytest.mark.parametrize("ne,p,procs", [((4, 4, 4), 2, (2, 2, 2)), ((5, 3, 4), 3, (2, 1, 2)), ((4, 2, 2), 1, (4, 1, 1)), ((3, 4, 5), 2, (1, 2, 3))])
def test_partitioned_c_builder_equals_numpy_builder(ne, p, procs):
    nranks = procs[0] * procs[1] * procs[2]
    owned = 0
    for r in range(nranks):
        s = pybox.build_box_space(ne, p, procs=procs, rank=r)
        m = lib.BoxMesh(ne, p, procs=procs, rank=r)
        assert (m.nelem, m.n, m.nc, m.ngh, m.rstart, m.nglobal) == (s.nelem, s.n, s.nc, s.ngh, s.rstart, s.nglobal)
        assert np.array_equal(m.e2l, s.e2l) and np.array_equal(m.l2g, s.l2g)
        assert [h[0] for h in m.halos] == [h.rank for h in s.halos]
        for (rk, snd, rcv), h in zip(m.halos, s.halos):
            assert np.array_equal(snd, h.send) and np.array_equal(rcv, h.recv)
        owned += m.n
    assert owned == m.nglobal == np.prod([n * p - 1 for n in ne])     # full-boundary Dirichlet: interior nodes


def _matsetvalues_pattern(space, eidx, bs=1):
    """Literal emulation of what PETSc is left with: for every element and every patch, MatSetValuesBlockedLocal(ADD_VALUES) of the
    P x P block matrix on the patch's nodes (dFSMatSetValuesBlockedExpanded, fs.c:826-930; negative = Dirichlet rows / columns are
    dropped, fsbuild.c:31-34), then assembly = per row the sorted set of the columns ever inserted, zeros kept.  Pure-Python sets;
    the patch nodes come from the ORACLE's restatement of dEFSGetExplicitSparse (efstopo.c:676-684 / jacobi.c:611-643), not from box.py."""
    np3 = (space.p + 1) ** 3
    rows = [set() for _ in range(space.n * bs)]
    for e in range(space.nelem):
        g = space.l2g[space.e2l[e * np3:(e + 1) * np3]]
        for patch in eidx:
            gl = [int(g[i]) for i in patch]
            for gi in gl:
                if gi < 0:
                    continue
                for gj in gl:
                    if gj < 0:
                        continue
                    for fi in range(bs):
                        rows[gi * bs + fi].update(gj * bs + fj for fj in range(bs))
    rowptr = np.concatenate(([0], np.cumsum([len(r) for r in rows]))).astype(np.int32)
    colidx = np.array([c for r in rows for c in sorted(r)], dtype=np.int32)
    return rowptr, colidx


@pytest.mark.parametrize("p,bs", [(1, 1), (2, 1), (3, 3), (4, 1)])
def test_q1_and_dense_patterns_equal_a_literal_matsetvalues_emulation(p, bs):
    """dohp_b200/box.py:q1_pattern / dense_pattern (what the GPU patterns are compared with bit for bit) against an independent
    insertion-by-insertion emulation driven by the oracle's patch index lists, on a box with a partial Dirichlet boundary."""
    import oracle_lib as ol
    ol.build_oracle(ref=False)
    s = pybox.build_box_space((3, 2, 2), p, dirichlet=(100, 300))
    _, eidx_sparse, _, _ = ol.OracleEFS(p, method=ol.SPARSE).explicit_sparse(np.tile(np.eye(3).ravel(), (2 * p) ** 3))
    assert eidx_sparse.shape == (p ** 3, 8)
    rp, ci = pybox.q1_pattern(s, bs=bs)
    erp, eci = _matsetvalues_pattern(s, eidx_sparse, bs)
    assert np.array_equal(rp, erp) and np.array_equal(ci, eci)
    if bs == 1:
        _, eidx_fast, _, _ = ol.OracleEFS(p, method=ol.FAST).explicit_sparse(np.tile(np.eye(3).ravel(), (p + 1) ** 3))
        assert eidx_fast.shape == (1, (p + 1) ** 3)            # quirk (iii): one patch holding every basis function of the element
        drp, dci = pybox.dense_pattern(s)
        erp, eci = _matsetvalues_pattern(s, eidx_fast, 1)
        assert np.array_equal(drp, erp) and np.array_equal(dci, eci)
