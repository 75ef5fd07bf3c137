"""Generic (unstructured, arbitrarily oriented) conforming hex meshes -> the index arrays the kernels consume: SURVEY.md 8(f1).

The reference builds a function space from a MOAB mesh through dFSBuildSpace_Cont (src/fs/impls/cont/cont.c:93-385) and the tensor
Jacobi's constraint assembly (dJacobiAddConstraints_Tensor, src/jacobi/impls/tensor/tensor.c:360-468: vertices, then the 12 edges with
forward / reverse traversal, then the 6 faces with the 8 quad permutations, then the region interior; one unit entry per expanded
node, so E is the index array ``e2l``).  MOAB does not exist here; the input format is the minimum a mesh file holds:

    verts[nv, 3]  vertex coordinates          hexes[ne, 8]  vertex ids in the MOAB / iMesh hex order
                                              (0-3 bottom face counter-clockwise, 4-7 above them; tensor.c:419-421)

Edges and faces are discovered from the connectivity (creation order = first appearance scanning elements and their local edges /
faces in MOAB's local order), each with the orientation of the element that created it; every other element maps its nodes onto the
entity's own frame, which is what the permutation tables of tensor.c:327-345,376-395 do.  Node numbering inside an entity follows
tensor.c: edges along their direction, faces lexicographic in the quad's own (v0->v1 slow, v0->v3 fast) axes, regions in tensor order.

Entity ORDER:  "natural" = type-major (vertices, edges, faces, regions) in creation order -- what every index-printing golden of the
reference uses -- or "rcm" = the reference's default (src/fs/interface/fsreg.c:55-56): dMeshPopulateOrderedSet_Private
(cont.c:93-178) builds the symmetric graph of the owned explicit entities (an entity is adjacent to everything in its closure and to
everything it bounds, plus the diagonal), asks PETSc for MATORDERINGRCM and permutes the ENTITIES; nodes follow their entity.
PETSc's RCM is SPARSPAK's GENRCM (pseudo-peripheral root by FNROOT, Cuthill-McKee with neighbours sorted by degree, reversed), on
the sorted adjacency of MatGetRowIJ: restated in `genrcm` below.  Upstream prints no index under RCM, but its Stokes
goldens were produced in this numbering and depend on it through ILU(0): `genblock_hex_space` + tests/test_stokes_refout_replay.py reproduce
stokes-q3q2-fgmres.refout (inner CG counts and 6-digit FGMRES histories) only with it, which pins the restatement; a hand-computed case and
numbering-independent solver goldens are checked as well.

Setup code (host, numpy): nothing here runs per apply.  Single rank (no ghosts); partitions come from dohp_b200/box.py."""
from dataclasses import dataclass
from typing import Optional

import numpy as np

# tensor coordinates (i, j, k in {0,1}) of the MOAB hex vertices 0..7 (tensor.c:419-421)
HEX_VERT = np.array([(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)])
HEX_EDGE = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]      # tensor.c:424-428
HEX_FACE = [(0, 1, 5, 4), (1, 2, 6, 5), (2, 3, 7, 6), (3, 0, 4, 7), (0, 3, 2, 1), (4, 5, 6, 7)]                    # tensor.c:440-444


@dataclass
class HexSpace:
    p: int
    bs: int
    nelem: int
    n: int                 # explicit nodes
    nc: int                # closure = explicit + Dirichlet
    ngh: int
    rstart: int
    nglobal: int
    e2l: np.ndarray        # [nelem * (p+1)^3] int32, element tensor order, k fastest
    l2g: np.ndarray        # [nc] int32, negative for Dirichlet
    coords: np.ndarray     # [nelem, 8, 3] vertices in tensor order (fsgeom.c:150-152)
    node_xyz: np.ndarray   # [nc, 3]
    entity_order: Optional[np.ndarray] = None      # explicit entities in their final order: (type, creation index) rows
    halos: tuple = ()

    @property
    def nloc(self):
        return self.nc + self.ngh


# ---- SPARSPAK GENRCM (George & Liu), as PETSc's MatGetOrdering(MATORDERINGRCM) runs it -----------------------------------------------------
def _rootls(root, xadj, adjncy, mask, ls):
    """ROOTLS: level structure rooted at `root` of the masked component; fills ls, returns (levels, level offsets, component size)."""
    mask[root] = 0
    ls[0] = root
    xls = []
    nlvl, lvlend, ccsize = 0, 0, 1
    while True:
        lbegin, lvlend = lvlend, ccsize
        nlvl += 1
        xls.append(lbegin)
        for i in range(lbegin, lvlend):
            node = ls[i]
            for j in range(xadj[node], xadj[node + 1]):
                nbr = adjncy[j]
                if mask[nbr]:
                    ls[ccsize] = nbr
                    ccsize += 1
                    mask[nbr] = 0
        if ccsize - lvlend <= 0:
            break
    xls.append(lvlend)
    for i in range(ccsize):
        mask[ls[i]] = 1
    return nlvl, xls, ccsize


def _fnroot(root, xadj, adjncy, mask, ls):
    """Pseudo-peripheral node of the masked component containing `root` (FNROOT)."""
    nlvl, xls, ccsize = _rootls(root, xadj, adjncy, mask, ls)
    if nlvl == 1 or nlvl == ccsize:
        return root, nlvl, xls, ccsize
    while True:
        jstrt = xls[nlvl - 1]
        mindeg = ccsize
        root = ls[jstrt]
        if ccsize != jstrt + 1:
            for j in range(jstrt, ccsize):
                node = ls[j]
                ndeg = sum(1 for k in range(xadj[node], xadj[node + 1]) if mask[adjncy[k]] > 0)
                if ndeg < mindeg:
                    root, mindeg = node, ndeg
        nunlvl, xls, ccsize = _rootls(root, xadj, adjncy, mask, ls)
        if nunlvl <= nlvl:
            return root, nunlvl, xls, ccsize
        nlvl = nunlvl
        if nlvl >= ccsize:
            return root, nlvl, xls, ccsize


def _rcm_component(root, xadj, adjncy, mask, ccsize_hint):
    """RCM of the masked component containing `root` (DEGREE + RCM of SPARSPAK); returns the ordering, unmasks the component."""
    # degrees within the component (DEGREE): breadth first from root over masked nodes
    comp = [root]
    seen = {root}
    i = 0
    while i < len(comp):
        node = comp[i]
        i += 1
        for j in range(xadj[node], xadj[node + 1]):
            nbr = adjncy[j]
            if mask[nbr] and nbr not in seen:
                seen.add(nbr)
                comp.append(nbr)
    deg = {node: sum(1 for j in range(xadj[node], xadj[node + 1]) if adjncy[j] in seen) for node in comp}
    perm = [root]
    mask[root] = 0
    lbegin, lvlend = 0, 1
    while lvlend > lbegin:
        lnbr = lvlend
        for i in range(lbegin, lvlend):
            node = perm[i]
            fnbr = lnbr
            for j in range(xadj[node], xadj[node + 1]):
                nbr = adjncy[j]
                if mask[nbr]:
                    mask[nbr] = 0
                    perm.append(nbr)
                    lnbr += 1
            # sort the neighbours just added by increasing degree: linear insertion, stable
            for k in range(fnbr + 1, lnbr):
                nbr = perm[k]
                l = k
                while l > fnbr and deg[perm[l - 1]] > deg[nbr]:
                    perm[l] = perm[l - 1]
                    l -= 1
                perm[l] = nbr
        lbegin, lvlend = lvlend, lnbr
    perm.reverse()
    return perm


def genrcm(xadj, adjncy):
    """perm[new] = old.  xadj/adjncy: symmetric structure, neighbours sorted (MatGetRowIJ); diagonal entries are ignored."""
    n = len(xadj) - 1
    xa = [0]
    adj = []
    for i in range(n):                       # drop the diagonal: it adds one to every degree and never changes an ordering decision
        adj.extend(int(j) for j in adjncy[xadj[i]:xadj[i + 1]] if j != i)
        xa.append(len(adj))
    mask = [1] * n
    ls = [0] * n
    perm = []
    for i in range(n):
        if not mask[i]:
            continue
        root, _, _, ccsize = _fnroot(i, xa, adj, mask, ls)
        perm.extend(_rcm_component(root, xa, adj, mask, ccsize))
    return np.array(perm, dtype=np.int64)


# ---- space builder ------------------------------------------------------------------------------------------------------------------------
def boundary_faces(hexes):
    """Faces (sorted vertex 4-tuples) that belong to exactly one element."""
    cnt = {}
    for h in hexes:
        for f in HEX_FACE:
            k = tuple(sorted(int(h[v]) for v in f))
            cnt[k] = cnt.get(k, 0) + 1
    return [k for k, c in cnt.items() if c == 1]


def build_hex_space(verts, hexes, p, bs=1, dirichlet_faces=None, ordering="natural", edges=None, faces=None):
    """Continuous degree-p space on an unstructured hex mesh.  dirichlet_faces: iterable of faces (any vertex order); their closure
    (vertices, edges, the faces themselves) carries Dirichlet nodes (dFSRegisterBoundary + closure, fs.c:182-212).  None = the
    whole boundary.  edges / faces: entities the mesh file created explicitly, in ITS creation order and with ITS vertex order (the
    entity's own frame), as the reference's hand-built test meshes do (iMesh_createEntArr in src/fs/tests/ex1.c:85-99, cunit.c:124-131);
    whatever the file does not list is discovered from the connectivity."""
    verts = np.asarray(verts, dtype=np.float64)
    hexes = np.asarray(hexes, dtype=np.int64)
    ne, d = len(hexes), p + 1
    if dirichlet_faces is None:
        dirichlet_faces = boundary_faces(hexes)
    # ---- entities in creation order, each with the frame of its creator ---------------------------------------------------------------
    edge_id, face_id = {}, {}
    given_edges, given_faces = ([] if edges is None else edges), ([] if faces is None else faces)
    edges, faces = [], []                      # (a, b) ; (v0, v1, v2, v3)
    for a, b in given_edges:
        edge_id[(min(int(a), int(b)), max(int(a), int(b)))] = len(edges)
        edges.append((int(a), int(b)))
    for f in given_faces:
        face_id[tuple(sorted(int(v) for v in f))] = len(faces)
        faces.append(tuple(int(v) for v in f))
    for h in hexes:
        for a, b in HEX_EDGE:
            k = (min(h[a], h[b]), max(h[a], h[b]))
            if k not in edge_id:
                edge_id[k] = len(edges)
                edges.append((int(h[a]), int(h[b])))
        for f in HEX_FACE:
            k = tuple(sorted(int(h[v]) for v in f))
            if k not in face_id:
                face_id[k] = len(faces)
                faces.append(tuple(int(h[v]) for v in f))
    nv, nedge, nface = len(verts), len(edges), len(faces)
    # Dirichlet closure
    dv, de, df = np.zeros(nv, bool), np.zeros(nedge, bool), np.zeros(nface, bool)
    for f in dirichlet_faces:
        k = tuple(sorted(int(v) for v in f))
        fid = face_id[k]
        df[fid] = True
        v0, v1, v2, v3 = faces[fid]
        for v in (v0, v1, v2, v3):
            dv[v] = True
        for a, b in ((v0, v1), (v1, v2), (v2, v3), (v3, v0)):
            de[edge_id[(min(a, b), max(a, b))]] = True
    # used vertices only (a mesh file may carry unused ones)
    used = np.zeros(nv, bool)
    used[hexes.ravel()] = True
    # ---- entity table: type-major creation order; size of each entity's node block -----------------------------------------------------
    pm = p - 1
    ent_type = np.concatenate([np.zeros(nv, int), np.ones(nedge, int), np.full(nface, 2), np.full(ne, 3)])
    ent_idx = np.concatenate([np.arange(nv), np.arange(nedge), np.arange(nface), np.arange(ne)])
    ent_size = np.concatenate([used.astype(int), np.full(nedge, pm), np.full(nface, pm * pm), np.full(ne, pm ** 3)])
    ent_diri = np.concatenate([dv, de, df, np.zeros(ne, bool)])
    off_type = np.array([0, nv, nv + nedge, nv + nedge + nface])
    nent = len(ent_type)
    explicit = np.nonzero(~ent_diri & (ent_size > 0))[0]
    if ordering == "rcm":
        # graph over the explicit entities (cont.c:131-150): closure relation both ways, indices = position in the natural explicit list
        pos = np.full(nent, -1)
        pos[explicit] = np.arange(len(explicit))
        nbrs = [set() for _ in range(len(explicit))]

        def link(a, b):
            if pos[a] >= 0 and pos[b] >= 0:
                nbrs[pos[a]].add(pos[b])
                nbrs[pos[b]].add(pos[a])
        for e, h in enumerate(hexes):
            r = off_type[3] + e
            evs = [int(v) for v in h]
            ees = [off_type[1] + edge_id[(min(h[a], h[b]), max(h[a], h[b]))] for a, b in HEX_EDGE]
            efs = [off_type[2] + face_id[tuple(sorted(int(h[v]) for v in f))] for f in HEX_FACE]
            for x in evs + ees + efs:
                link(r, x)
            for fi, f in enumerate(HEX_FACE):
                for v in f:
                    link(efs[fi], int(h[v]))
                for a, b in ((f[0], f[1]), (f[1], f[2]), (f[2], f[3]), (f[3], f[0])):
                    link(efs[fi], off_type[1] + edge_id[(min(h[a], h[b]), max(h[a], h[b]))])
            for ei, (a, b) in enumerate(HEX_EDGE):
                link(ees[ei], int(h[a]))
                link(ees[ei], int(h[b]))
        xadj, adjncy = [0], []
        for s in nbrs:
            adjncy.extend(sorted(s))
            xadj.append(len(adjncy))
        perm = genrcm(np.array(xadj), np.array(adjncy, dtype=np.int64))
        explicit = explicit[perm]
    elif ordering != "natural":
        raise ValueError("unknown ordering %r" % (ordering,))
    dirichlet = np.nonzero(ent_diri & (ent_size > 0))[0]
    order = np.concatenate([explicit, dirichlet])
    start = np.full(nent, -1, dtype=np.int64)
    start[order] = np.concatenate([[0], np.cumsum(ent_size[order])[:-1]])
    n = int(ent_size[explicit].sum())
    nc = n + int(ent_size[dirichlet].sum())
    # ---- e2l: every element maps its nodes onto the frames of its entities (tensor.c:360-468) ---------------------------------------------
    e2l = np.empty((ne, d, d, d), dtype=np.int64)
    corner_of = {tuple(c): v for v, c in enumerate(HEX_VERT)}
    ii, jj, kk = np.meshgrid(np.arange(d), np.arange(d), np.arange(d), indexing="ij")
    X = np.stack([ii, jj, kk], axis=-1).reshape(-1, 3)
    onb = (X == 0) | (X == p)
    nfix = onb.sum(axis=1)
    for e, h in enumerate(hexes):
        out = np.empty(d ** 3, dtype=np.int64)
        C = HEX_VERT * p                                        # tensor coordinates of this element's corners
        for m in range(d ** 3):
            x = X[m]
            if nfix[m] == 3:
                out[m] = start[int(h[corner_of[tuple(x // p)]])]
            elif nfix[m] == 2:
                free = int(np.nonzero(~onb[m])[0][0])
                ca, cb = x.copy(), x.copy()
                ca[free], cb[free] = 0, p
                va, vb = int(h[corner_of[tuple(ca // p)]]), int(h[corner_of[tuple(cb // p)]])
                eid = edge_id[(min(va, vb), max(va, vb))]
                s = x[free] if edges[eid][0] == va else p - x[free]
                out[m] = start[off_type[1] + eid] + (s - 1)
            elif nfix[m] == 1:
                fixed = int(np.nonzero(onb[m])[0][0])
                corners = [c for c in range(8) if C[c][fixed] == x[fixed]]
                fid = face_id[tuple(sorted(int(h[c]) for c in corners))]
                v0, v1, _, v3 = faces[fid]
                loc = {int(h[c]): C[c] for c in corners}
                c0, c1, c3 = loc[v0], loc[v1], loc[v3]
                s = int(np.dot(x - c0, c1 - c0)) // p
                t = int(np.dot(x - c0, c3 - c0)) // p
                out[m] = start[off_type[2] + fid] + (s - 1) * pm + (t - 1)
            else:
                out[m] = start[off_type[3] + e] + ((x[0] - 1) * pm + (x[1] - 1)) * pm + (x[2] - 1)
        e2l[e] = out.reshape(d, d, d)
    # ---- geometry and nodal coordinates ------------------------------------------------------------------------------------------------------
    coords = np.empty((ne, 8, 3))
    for t in range(8):
        c = ((t >> 2) & 1, (t >> 1) & 1, t & 1)
        coords[:, t, :] = verts[hexes[:, corner_of[c]]]
    from .box import _gll
    g1 = _gll(d)
    W = np.stack([0.5 * (1 - g1), 0.5 * (1 + g1)], axis=0)
    xe = np.einsum("ia,jb,kc,eijkx->eabcx", W, W, W, coords.reshape(ne, 2, 2, 2, 3))
    node_xyz = np.zeros((nc, 3))
    node_xyz[e2l.reshape(-1)] = xe.reshape(-1, 3)
    l2g = np.full(nc, -1, dtype=np.int64)
    l2g[:n] = np.arange(n)
    return HexSpace(p=p, bs=bs, nelem=ne, n=n, nc=nc, ngh=0, rstart=0, nglobal=n, e2l=e2l.reshape(-1).astype(np.int32), l2g=l2g.astype(np.int32),
                    coords=coords, node_xyz=node_xyz, entity_order=np.stack([ent_type[explicit], ent_idx[explicit]], axis=1))


def box_hexes(ne, box=((-1., 1.), (-1., 1.), (-1., 1.)), morph=None, scramble=None):
    """Vertex list + hex connectivity of an ne[0] x ne[1] x ne[2] box (vertex id (i*(ny+1)+j)*(nz+1)+k, elements k fastest as genblock
    writes them).  scramble: numpy Generator -- every element gets a random one of the 24 proper rotations of its vertex list, so the
    shared edges and faces are seen in all relative orientations."""
    from .box import morph_coords
    nx, ny, nz = ne
    g = [np.linspace(b[0], b[1], n + 1) for b, n in zip(box, ne)]
    V = np.stack(np.meshgrid(*g, indexing="ij"), axis=-1)
    if morph:
        V = morph_coords(V, **morph)
    verts = V.reshape(-1, 3)
    vid = lambda i, j, k: (i * (ny + 1) + j) * (nz + 1) + k
    rots = _hex_rotations()
    hexes = []
    for i in range(nx):
        for j in range(ny):
            for k in range(nz):
                h = [vid(i + c[0], j + c[1], k + c[2]) for c in HEX_VERT]
                if scramble is not None:
                    h = [h[q] for q in rots[scramble.integers(len(rots))]]
                hexes.append(h)
    return verts, np.array(hexes)


def genblock_entities(ne):
    """The edges and faces the reference's block generator creates EXPLICITLY, in its creation order and with its vertex order
    (src/fs/mesh/interface/genblock.c:309-347 faces: z-normal, y-normal, x-normal; :352-390 edges: along z, y, x), for the vertex ids of
    `box_hexes` (genblock.c:208-212: v[(i*n+j)*p+k]).  MOAB orders handles by entity type first, so the creation sequence across types
    (genblock creates the faces before the edges) does not matter."""
    m, n, q = ne[0] + 1, ne[1] + 1, ne[2] + 1
    v = lambda i, j, k: (i * n + j) * q + k
    edges = [(v(i, j, k), v(i, j, k + 1)) for i in range(m) for j in range(n) for k in range(q - 1)]
    edges += [(v(i, j, k), v(i, j + 1, k)) for i in range(m) for j in range(n - 1) for k in range(q)]
    edges += [(v(i, j, k), v(i + 1, j, k)) for i in range(m - 1) for j in range(n) for k in range(q)]
    faces = [(v(i, j, k), v(i + 1, j, k), v(i + 1, j + 1, k), v(i, j + 1, k)) for i in range(m - 1) for j in range(n - 1) for k in range(q)]
    faces += [(v(i, j, k), v(i + 1, j, k), v(i + 1, j, k + 1), v(i, j, k + 1)) for i in range(m - 1) for j in range(n) for k in range(q - 1)]
    faces += [(v(i, j, k), v(i, j + 1, k), v(i, j + 1, k + 1), v(i, j, k + 1)) for i in range(m) for j in range(n - 1) for k in range(q - 1)]
    return edges, faces


def genblock_hex_space(ne, p, bs=1, dirichlet=True, ordering="rcm", morph=None):
    """The function space `-dmesh_gen block` + dFSBuildSpace_Cont give on an ne[0] x ne[1] x ne[2] box, through the GENERIC builder fed
    with genblock's own entity creation order: ordering="natural" reproduces dohp_b200.box.build_box_space node for node (tests/test_mesh.py),
    ordering="rcm" is the reference's DEFAULT numbering (src/fs/interface/fsreg.c:55-56), the one its Stokes goldens were produced with.
    That numbering is pinned by stokes-q3q2-fgmres.refout: ILU(0) of the velocity and pressure preconditioning matrices depends on it,
    and the golden's inner CG iteration counts and 6-digit FGMRES residual histories come out only in this numbering
    (tests/test_stokes_refout_replay.py).  dirichlet: True = all six walls (velocity), False = none (the Stokes pressure space)."""
    verts, hexes = box_hexes(ne, morph=morph)
    edges, faces = genblock_entities(ne)
    kw = {} if dirichlet else dict(dirichlet_faces=[])
    return build_hex_space(verts, hexes, p, bs=bs, ordering=ordering, edges=edges, faces=faces, **kw)


def _hex_rotations():
    """The 24 orientation-preserving symmetries of the cube as permutations of the MOAB vertex numbering."""
    import itertools
    out = []
    pts = 2 * HEX_VERT - 1
    for perm in itertools.permutations(range(3)):
        for signs in itertools.product((1, -1), repeat=3):
            M = np.zeros((3, 3), int)
            for r in range(3):
                M[r, perm[r]] = signs[r]
            if round(np.linalg.det(M)) != 1:
                continue
            img = pts @ M.T
            out.append([int(np.nonzero((pts == q).all(axis=1))[0][0]) for q in img])
    return out


def subelement_mesh(space):
    """dFSGetSubElementMesh (src/fs/interface/fsgeom.c:273, dFSGetSubElementMesh_Cont src/fs/impls/cont/cont.c:417-468): the
    visualisation mesh of a space -- every element split into its p^3 Q1 sub-hexes (the patches of the SPARSE rule, in patch order,
    tensorquad.c:144-158), each as 8 closure node indices in the iTaps hex vertex order (cont.c:449: itaps_to_tensor) looked up through
    the boolean E.  Returns (offsets[nsub+1], connectivity[8*nsub]); the vertex coordinates are space.node_xyz."""
    p, d = space.p, space.p + 1
    e2l = np.asarray(space.e2l).reshape(space.nelem, d, d, d)
    itaps_to_tensor = (0, 4, 6, 2, 1, 5, 7, 3)
    sub = np.empty((space.nelem, p, p, p, 8), dtype=np.int64)
    for i, t in enumerate(itaps_to_tensor):
        a, b, c = (t >> 2) & 1, (t >> 1) & 1, t & 1             # tensor corner t of the patch: element node (pi+a, pj+b, pk+c)
        sub[..., i] = e2l[:, a:a + p, b:b + p, c:c + p]
    conn = sub.reshape(-1)
    return np.arange(0, len(conn) + 1, 8), conn
