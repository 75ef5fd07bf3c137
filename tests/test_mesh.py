"""the generic (unstructured, arbitrarily oriented) hex space builder and the RCM entity ordering
(dohp_b200/mesh.py) -- CPU only.  The builder is checked through the oracle: an operator applied on the generic space must equal the
same operator on the structured-box space (dohp_b200/box.py, pinned by the reference's bu-0 golden), node by node, also when every
element lists its vertices in a random one of the 24 cube orientations (all edge directions and all 8 face permutations of
tensor.c:327-345 occur).  RCM: SPARSPAK's GENRCM restated, checked on a hand-computed graph and through the numbering-independent
solver golden of the reference."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import oracle_lib as ol
from dohp_b200 import box, mesh
from dist_common import field, key


@pytest.fixture(scope="module", autouse=True)
def _build():
    ol.build_oracle(ref=False)


def test_genrcm_hand_computed():
    """Path graph 0-1-2-3-4 with a pendant 5 on node 2:   0 - 1 - 2 - 3 - 4
                                                                  |
                                                                  5
    FNROOT from node 0: levels {0},{1},{2},{3,5},{4}: 5 levels < 6 nodes; last level {4}: root 4; levels from 4: {4},{3},{2},{1,5},{0}:
    not deeper -> root 4.  Cuthill-McKee from 4: 4, 3, 2, then the neighbours of 2 not yet numbered in adjacency order (1, 5) sorted by
    degree (deg 1 = 2, deg 5 = 1): 5, 1; then 0.  CM = 4 3 2 5 1 0, reversed: 0 1 5 2 3 4."""
    nbrs = {0: [1], 1: [0, 2], 2: [1, 3, 5], 3: [2, 4], 4: [3], 5: [2]}
    xadj, adj = [0], []
    for i in range(6):
        adj += sorted(nbrs[i] + [i])          # with the diagonal, as cont.c:150-153 inserts it
        xadj.append(len(adj))
    assert mesh.genrcm(np.array(xadj), np.array(adj)).tolist() == [0, 1, 5, 2, 3, 4]


def test_genrcm_two_components_and_bandwidth():
    rng = np.random.default_rng(3)
    n = 60
    A = sp.random(n, n, 0.05, random_state=4, format="csr")
    A = ((A + A.T) != 0).astype(int) + sp.eye(n, dtype=int)
    A = sp.csr_matrix(A)
    A.sort_indices()
    perm = mesh.genrcm(A.indptr, A.indices)
    assert sorted(perm.tolist()) == list(range(n))
    def bw(M):
        c = M.tocoo()
        return int(np.abs(c.row - c.col).max())
    B = A[perm][:, perm]
    assert bw(B) <= bw(A)


def _apply_ellip(space, x_of_xyz):
    o = ol.OracleEllip(space, p=1.7, lam=0.3, eps=0.8, exact=0)
    x = x_of_xyz(space.node_xyz[:space.n])
    o.function(0.3 * x, dcache=np.zeros(space.nc - space.n))
    return dict(zip(key(space.node_xyz[:space.n]), o.mult(x)))


@pytest.mark.parametrize("p", [1, 2, 3, 4])
@pytest.mark.parametrize("scramble,ordering", [(False, "natural"), (True, "natural"), (True, "rcm")])
def test_generic_builder_equals_box_builder_through_the_oracle(p, scramble, ordering):
    ne, morph = (3, 2, 2), dict(morph=0.3, twist=0.2)
    ref = _apply_ellip(box.build_box_space(ne, p, morph=morph), field)
    verts, hexes = mesh.box_hexes(ne, morph=morph, scramble=np.random.default_rng(p) if scramble else None)
    sp_ = mesh.build_hex_space(verts, hexes, p, ordering=ordering)
    got = _apply_ellip(sp_, field)
    assert set(got) == set(ref)
    scale = max(abs(v) for v in ref.values())
    assert max(abs(got[k] - ref[k]) for k in ref) <= 1e-12 * scale


def test_e2l_is_a_conforming_numbering():
    """Every local node index is hit, shared nodes coincide geometrically (all elements agree on the coordinates of a node)."""
    verts, hexes = mesh.box_hexes((2, 3, 2), morph=dict(morph=0.2, twist=0.3), scramble=np.random.default_rng(1))
    s = mesh.build_hex_space(verts, hexes, 3, ordering="rcm")
    assert sorted(set(s.e2l.tolist())) == list(range(s.nc))
    g = box._gll(4)
    W = np.stack([0.5 * (1 - g), 0.5 * (1 + g)], axis=0)
    xe = np.einsum("ia,jb,kc,eijkx->eabcx", W, W, W, s.coords.reshape(s.nelem, 2, 2, 2, 3)).reshape(-1, 3)
    assert np.abs(xe - s.node_xyz[s.e2l]).max() < 1e-13


def test_rcm_space_reproduces_the_reference_golden_norms():
    """ellip-e0-b3-p16.refout runs with the reference's DEFAULT ordering (RCM, fsreg.c:55-56): '0 SNES Function norm 4.93408' and
    '0 KSP Residual norm 158.134' are numbering independent and must come out of the RCM-ordered generic space as well."""
    verts, hexes = mesh.box_hexes((4, 4, 4))
    s = mesh.build_hex_space(verts, hexes, 3, ordering="rcm")
    o = ol.OracleEllip(s, p=1.6, exact=0, a=1.6, b=1.7, c=1.8)
    x0 = np.zeros(s.n)
    F = o.function(x0)
    assert float("%.6g" % np.linalg.norm(F)) == 4.93408
    # Q1 pattern on a generic space: union of the patch couplings
    d = 4
    e2l = s.e2l.reshape(s.nelem, d, d, d)
    rows, cols = [], []
    for i in range(3):
        for j in range(3):
            for k in range(3):
                c = e2l[:, i:i + 2, j:j + 2, k:k + 2].reshape(s.nelem, 8)
                rows.append(np.repeat(c, 8, axis=1).ravel())
                cols.append(np.tile(c, (1, 8)).ravel())
    rows, cols = np.concatenate(rows), np.concatenate(cols)
    keep = (rows < s.n) & (cols < s.n)
    pat = sp.csr_matrix((np.ones(keep.sum()), (rows[keep], cols[keep])), shape=(s.n, s.n))
    pat.sum_duplicates()
    pat.sort_indices()
    rp, ci = pat.indptr.astype(np.int32), pat.indices.astype(np.int32)
    vals = o.jacobian_q1(x0, rp, ci)
    z = spla.splu(sp.csr_matrix((vals, ci, rp), shape=(s.n, s.n)).tocsc()).solve(F)
    assert float("%.6g" % np.linalg.norm(z)) == 158.134
    # and the ordering does reduce the bandwidth of the matrix the reference hands to its factorisations
    sn = mesh.build_hex_space(verts, hexes, 3, ordering="natural")
    bwn = np.abs(np.diff(np.sort(sn.e2l.reshape(sn.nelem, -1), axis=1)[:, [0, -1]], axis=1)).max()
    bwr = np.abs(np.diff(np.sort(np.where(s.e2l.reshape(s.nelem, -1) < s.n, s.e2l.reshape(s.nelem, -1), 0), axis=1)[:, [0, -1]], axis=1)).max()
    assert bwr < bwn


def _cunit(name):
    """-> (closure index of every expanded node, expanded coordinate vector, nodal coordinates in expanded space) as printed by
    src/fs/tests/cunit.c (-expanded_view, -cexp_view, -nce_view), plus the header counts"""
    import os
    import re
    txt = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refout", name + ".refout")).read()

    def rows(after, before):
        sec = txt.split(after)[1]
        sec = sec.split(before)[0] if before else sec
        return np.array([[float(v) for v in ln.split(":")[1].split()] for ln in sec.split("\n") if ln.startswith("[0]")])
    idx = None
    if "Expanded vector (-expanded_view)" in txt:
        sec = txt.split("Expanded vector (-expanded_view)")[1].split("Expanded coordinate vector")[0]
        idx = np.array([int(ln) for ln in sec.split("\n") if re.fullmatch(r"\d+", ln.strip())]) - 1000
    cexp = rows("Expanded coordinate vector", "Nodal coordinates in expanded") if "Expanded coordinate vector" in txt else None
    nce = rows("Nodal coordinates in expanded space", "Comparison of re-expanded")
    m = re.search(r"number of vertices=(\d+) edges=(\d+) faces=(\d+) regions=(\d+)", txt)
    n = re.search(r"(\d+)/\d+ expanded nodes constrained against (\d+)\+0", txt)
    return idx, cexp, nce, tuple(int(v) for v in m.groups()), (int(n.group(1)), int(n.group(2))), "Node " in txt.split("Comparison of re-expanded")[1]


def test_cunit_goldens_on_the_hand_built_two_hex_mesh():
    """cunit-identity-0 / cunit-affine-0 (src/fs/tests/cunit.c, CMakeLists.txt:44-45; -dfs_ordering_type natural, degree 2): the second hex
    is rotated against the first, edges and faces are created explicitly in the test's own order.  The generic builder must give the printed
    closure index of all 54 expanded nodes (the boolean E), the expanded geometry vector in tensor order (fsgeom.c:150-152) and the
    coordinates evaluated at the nodes of the solution basis; the re-expansion comparison printed no mismatch upstream and finds none here."""
    for name, A, b in (("cunit-identity-0", np.eye(3), np.zeros(3)), ("cunit-affine-0", np.array([[1., 2, 3], [4, 5, 6], [7, 8, 9]]), np.array([10., 11, 12]))):
        idx, cexp, nce, counts, sizes, mismatch = _cunit(name)
        verts = ol.EX1_VERTS_AFFINE @ A.T + b                    # DeformApply_Affine, cunit.c:35-48
        s = mesh.build_hex_space(verts, ol.EX1_HEXES, 2, dirichlet_faces=[], edges=ol.EX1_EDGES, faces=ol.EX1_FACES)
        assert counts == (12, 20, 11, 2) and sizes == (s.nelem * 27, s.n) and not mismatch
        if idx is not None:
            assert np.array_equal(s.e2l, idx)
        if cexp is not None:
            assert np.array_equal(s.coords.reshape(-1, 3), cexp)
        assert np.abs(s.node_xyz[s.e2l] - nce).max() < 1e-12      # printed with %10G: exact for these half-integers
        # without the explicit lists the entities are discovered in another order: same space, another (valid) numbering
        s2 = mesh.build_hex_space(verts, ol.EX1_HEXES, 2, dirichlet_faces=[])
        assert s2.n == s.n and np.abs(s2.node_xyz[s2.e2l] - nce).max() < 1e-12


def test_projection_mesh_with_explicit_entities_gives_the_same_known_answers():
    """ex1.c creates the same edges / faces explicitly; the projection goldens are numbering independent: 0.353633 either way"""
    s = mesh.build_hex_space(ol.EX1_VERTS_PARAMETRIC, ol.EX1_HEXES, 5, dirichlet_faces=[], edges=ol.EX1_EDGES, faces=ol.EX1_FACES)
    o = ol.OracleProj(s)
    assert float("%.6g" % np.linalg.norm(o.function(np.zeros(s.n)))) == 0.353633


# the three hexes around the z axis of src/fs/tests/fsex4.c:13-25 (vertices, explicitly created edges / faces, regions)
FSEX4_VTX = np.array([0, 0, 0, 1, 0, 0, 0.5, 1.5, 0, -0.3, 0.8, 0, -1.5, 0, 0, -0.3, -0.8, 0, 0.5, -1.5, 0,
                      0, 0, 1, 1, 0, 1, 0.5, 1.5, 1, -0.3, 0.8, 1, -1.5, 0, 1, -0.3, -0.8, 1, 0.5, -1.5, 1]).reshape(14, 3)
FSEX4_EDGES = np.array([0, 7, 1, 8, 2, 9, 3, 10, 4, 11, 5, 12, 6, 13, 0, 1, 1, 2, 2, 3, 0, 3, 3, 4, 4, 5, 0, 5, 5, 6, 6, 1,
                        7, 8, 8, 9, 9, 10, 7, 10, 10, 11, 11, 12, 7, 12, 12, 13, 13, 8]).reshape(-1, 2)
FSEX4_FACES = np.array([0, 1, 2, 3, 0, 3, 4, 5, 0, 5, 6, 1, 7, 8, 9, 10, 7, 10, 11, 12, 7, 12, 13, 8, 0, 1, 8, 7, 1, 2, 9, 8, 2, 3, 10, 9,
                        0, 3, 10, 7, 3, 4, 11, 10, 4, 5, 12, 11, 0, 5, 12, 7, 5, 6, 13, 12, 6, 1, 8, 13]).reshape(-1, 4)
FSEX4_HEXES = np.array([0, 1, 2, 3, 7, 8, 9, 10, 0, 3, 4, 5, 7, 10, 11, 12, 0, 5, 6, 1, 7, 12, 13, 8]).reshape(3, 8)


@pytest.mark.parametrize("name,deg", [("fsex4-submesh-p1", 1), ("fsex4-submesh-p1-read", 1), ("fsex4-submesh-p2-read", 2)])
def test_fsex4_subelement_mesh_goldens(name, deg):
    """fsex4-submesh-p{1,2}* (src/fs/tests/fsex4.c:54-83, -dfs_ordering_type natural): dFSGetSubElementMesh on three hexes that meet in an
    edge, i.e. three different frames for the shared entities.  The printed sizes, offsets, connectivity (closure node numbers of every
    Q1 sub-hex in iTaps order) and the global nodal coordinates must come out of the generic builder exactly."""
    import os
    import re
    txt = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refout", name + ".refout")).read()
    s = mesh.build_hex_space(FSEX4_VTX, FSEX4_HEXES, deg, dirichlet_faces=[], edges=FSEX4_EDGES, faces=FSEX4_FACES)
    m = re.search(r"number of vertices=(\d+) edges=(\d+) faces=(\d+) regions=(\d+)", txt)
    assert tuple(int(v) for v in m.groups()) == (len(FSEX4_VTX), len(FSEX4_EDGES), len(FSEX4_FACES), len(FSEX4_HEXES))
    m = re.search(r"(\d+)/\d+ expanded nodes constrained against (\d+)\+0", txt)
    assert (int(m.group(1)), int(m.group(2))) == (s.nelem * (deg + 1) ** 3, s.n)
    P = deg + 1
    assert s.n == 3 * P ** 3 - 3 * P * P + P                        # fsex4.c:66
    def ints(after, before):
        sec = txt.split(after)[1].split(before)[0]
        return np.array([int(v) for ln in sec.strip().split("\n") for v in ln.split(":")[1].split()])
    off, conn = mesh.subelement_mesh(s)
    assert int(re.search(r"Offsets for (\d+) subelements", txt).group(1)) == len(off) - 1 == 3 * deg ** 3
    assert np.array_equal(off, ints("subelements\n", "Connectivity"))
    assert np.array_equal(conn, ints("Connectivity\n", "Coordinates"))
    vals = []
    for ln in txt.split("Process [0]\n")[1].split("\n"):
        try:
            vals.append(float(ln))
        except ValueError:
            break
    assert len(vals) == 3 * s.n and np.array_equal(s.node_xyz[:s.n], np.array(vals).reshape(-1, 3))
    # every sub-hex is a positively oriented cell: the iTaps order gives a positive Jacobian at its first corner
    X = s.node_xyz[conn.reshape(-1, 8)]
    det = np.einsum("ni,ni->n", np.cross(X[:, 1] - X[:, 0], X[:, 3] - X[:, 0]), X[:, 4] - X[:, 0])
    assert (det > 0).all()
