"""Structured hex-box function spaces: the input side of the hot path.

The reference builds its function space from a MOAB mesh (``genblock`` structured block generator,
src/fs/mesh/interface/genblock.c:208-420) through ``dFSBuildSpace_Cont`` (src/fs/impls/cont/cont.c:258-385).  Neither
MOAB nor PETSc exists here, and the space builder is setup code, not the hot path, so this module restates *what those
two produce* for conforming isotropic-degree hex boxes -- the index arrays the kernels consume:

* ``e2l``  expanded node -> local node, i.e. the boolean element-assembly matrix E (tensor.c:360-468)
* local layout ``[explicit n | dirichlet nc-n | ghost ngh]`` (src/vec/vecd.c:123-151, fsbuild.c:107-146)
* ``l2g``  local node -> global row, negative for Dirichlet (fsbuild.c:27-37)
* the expanded Q1 geometry vector in tensor order (fsgeom.c:111-170)
* halo descriptors for the block partition genblock writes (genblock.c:257-304)

Numbering (``ordering="natural"``) is MOAB handle order = type-major (vertex, edge, face, region), creation order
within a type, interior nodes of an entity contiguous -- this is what every index-printing golden of the reference
uses (``-dfs_ordering_type natural``) and it is checked against src/fs/tests/refout/bu-0.refout in
tests/test_box.py.  ``ordering="element"`` is a locality-friendly alternative (entities ordered by the first
element that touches them); the reference's default RCM numbering is dohp_b200.mesh.genblock_hex_space(ordering="rcm").

Everything is vectorised numpy; nothing here runs per apply.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

WALLS, TOP, BOTTOM = 100, 200, 300     # genblock boundary set ids, genblock.c:407-420


def _gll(n):
    """Gauss-Lobatto-Legendre nodes (host-side helper for nodal coordinates only)."""
    if n == 2:
        return np.array([-1.0, 1.0])
    from numpy.polynomial import legendre as L
    c = np.zeros(n)
    c[-1] = 1
    r = L.legroots(L.legder(c))
    for _ in range(3):          # Newton polish on P'_{n-1}
        d1 = L.legval(r, L.legder(c))
        d2 = L.legval(r, L.legder(c, 2))
        r = r - d1 / d2
    r = np.sort(r)
    r = 0.5 * (r - r[::-1])
    return np.concatenate(([-1.0], r, [1.0]))


def morph_coords(xyz, morph=0.0, twist=0.0, stretch=0.0):
    """Mesh deformation used by the reference's non-affine tests (ellip.c:173-181)."""
    x, y, z = xyz[..., 0].copy(), xyz[..., 1].copy(), xyz[..., 2].copy()
    out = np.empty_like(xyz)
    out[..., 0] = np.cos(twist * z) * x + np.sin(twist * z) * y
    out[..., 1] = -np.sin(twist * z) * x + np.cos(twist * z) * y
    out[..., 2] = -1 + (z + 1) * (1 + morph * ((x + 1) ** 2 + (y + 1) ** 2)) + stretch * z
    return out


def part_range(nel, nparts, i):
    """Element range of part i in one direction: genblock.c:286-303 (ii from i*(m-1)/M to (i+1)*(m-1)/M)."""
    return (i * nel) // nparts, ((i + 1) * nel) // nparts


def part_of_element(e, nel, nparts):
    """Inverse of part_range, i.e. membership in the PARALLEL_PARTITION sets (genblock.c:286-303), which is what the
    parallel mesh load partitions by.  (The per-region tag formula I = i*M/(m-1), genblock.c:260-266, disagrees with
    the sets when the element count is not divisible by the part count; the sets win.)"""
    return ((e + 1) * nparts - 1) // nel


@dataclass
class Halo:
    """One neighbour: ``send`` are local indices of nodes this rank owns and the neighbour ghosts (forward INSERT,
    fs.c:578-579); ``recv`` are local ghost indices owned by the neighbour.  The reverse ADD (fs.c:627-628) uses the
    same lists with roles swapped.  Both ranks order the lists by global lexicographic node id."""
    rank: int
    send: np.ndarray
    recv: np.ndarray


@dataclass
class BoxSpace:
    ne: Tuple[int, int, int]          # global element grid
    p: int                            # polynomial degree (isotropic)
    bs: int
    nelem: int                        # local elements
    n: int                            # owned explicit nodes
    nc: int                           # closure = n + owned Dirichlet nodes
    ngh: int                          # ghost nodes
    rstart: int                       # first global row owned (MPI_Scan, cont.c:326-336)
    nglobal: int                      # global number of explicit nodes over all ranks
    e2l: np.ndarray                   # [nelem * (p+1)^3] int32
    l2g: np.ndarray                   # [nc+ngh] int32, negative for Dirichlet
    coords: np.ndarray                # [nelem, 8, 3] expanded geometry vector, tensor order
    node_xyz: np.ndarray              # [nc+ngh, 3] nodal coordinates (fsgeom.c:20-106)
    halos: List[Halo] = field(default_factory=list)
    elem_ijk: Optional[np.ndarray] = None    # [nelem,3] global element index of each local element
    rank: int = 0
    nranks: int = 1

    @property
    def nloc(self):
        return self.nc + self.ngh

    @property
    def np_elem(self):
        return (self.p + 1) ** 3

    @property
    def ndof_global(self):
        return self.nglobal * self.bs


def _entity_tables(NE, p, diri_sets):
    """Per entity type: creation-order index formulas and Dirichlet membership, over the GLOBAL mesh.

    Returns a function mapping global node coordinates (gx,gy,gz arrays) to
    (etype, entity creation index, index within entity, is_dirichlet)."""
    Nx, Ny, Nz = NE
    m, n, q = Nx + 1, Ny + 1, Nz + 1          # vertex counts per direction ("m,n,p" of genblock.c)

    def classify(gx, gy, gz):
        rx, ry, rz = gx % p, gy % p, gz % p
        I, J, K = gx // p, gy // p, gz // p
        ox, oy, oz = rx == 0, ry == 0, rz == 0
        non = ox.astype(np.int8) + oy.astype(np.int8) + oz.astype(np.int8)
        etype = (3 - non).astype(np.int8)     # 0 vertex, 1 edge, 2 face, 3 region
        ent = np.zeros(gx.shape, dtype=np.int64)
        within = np.zeros(gx.shape, dtype=np.int64)
        pm = p - 1
        # vertices: (i*n+j)*p+k, genblock.c:208-212
        s = etype == 0
        ent[s] = (I[s] * n + J[s]) * q + K[s]
        # edges, creation order z-, y-, x-direction: genblock.c:352-390
        s = (etype == 1) & ~oz                # runs along z
        ent[s] = (I[s] * n + J[s]) * (q - 1) + K[s]
        within[s] = rz[s] - 1
        off = m * n * (q - 1)
        s = (etype == 1) & ~oy
        ent[s] = off + (I[s] * (n - 1) + J[s]) * q + K[s]
        within[s] = ry[s] - 1
        off += m * (n - 1) * q
        s = (etype == 1) & ~ox
        ent[s] = off + (I[s] * n + J[s]) * q + K[s]
        within[s] = rx[s] - 1
        # faces, creation order z-normal, y-normal, x-normal: genblock.c:309-347; interior index is lexicographic in
        # the quad's own (v0->v1 slow, v0->v3 fast) axes = global lexicographic restricted to the face (tensor.c:328-345)
        s = (etype == 2) & oz
        ent[s] = (I[s] * (n - 1) + J[s]) * q + K[s]
        within[s] = (rx[s] - 1) * pm + (ry[s] - 1)
        off = (m - 1) * (n - 1) * q
        s = (etype == 2) & oy
        ent[s] = off + (I[s] * n + J[s]) * (q - 1) + K[s]
        within[s] = (rx[s] - 1) * pm + (rz[s] - 1)
        off += (m - 1) * n * (q - 1)
        s = (etype == 2) & ox
        ent[s] = off + (I[s] * (n - 1) + J[s]) * (q - 1) + K[s]
        within[s] = (ry[s] - 1) * pm + (rz[s] - 1)
        # regions: (i*(n-1)+j)*(p-1)+k, genblock.c:235-250; interior tensor.c:453-461
        s = etype == 3
        ent[s] = (I[s] * (n - 1) + J[s]) * (q - 1) + K[s]
        within[s] = ((rx[s] - 1) * pm + (ry[s] - 1)) * pm + (rz[s] - 1)
        # Dirichlet: closure of the registered face sets (fs.c:182-212): walls = x-,x+,y-,y+; 200 = z+; 300 = z-
        diri = np.zeros(gx.shape, dtype=bool)
        if WALLS in diri_sets:
            diri |= (gx == 0) | (gx == Nx * p) | (gy == 0) | (gy == Ny * p)
        if TOP in diri_sets:
            diri |= gz == Nz * p
        if BOTTOM in diri_sets:
            diri |= gz == 0
        return etype, ent, within, diri

    nent = [m * n * q,
            m * n * (q - 1) + m * (n - 1) * q + (m - 1) * n * q,
            (m - 1) * (n - 1) * q + (m - 1) * n * (q - 1) + m * (n - 1) * (q - 1),
            (m - 1) * (n - 1) * (q - 1)]
    return classify, nent


def build_box_space(ne, p, bs=1, dirichlet=(WALLS, TOP, BOTTOM), box=((-1., 1.), (-1., 1.), (-1., 1.)),
                    morph=None, procs=(1, 1, 1), rank=0, ordering="natural"):
    """Continuous-Galerkin degree-p space on an ne[0] x ne[1] x ne[2] hex box, as seen by ``rank`` of the
    procs[0]*procs[1]*procs[2] block partition (part id (I*N+J)*P+K, genblock.c:263-264).

    Ownership rule (the reference takes it from MOAB's parallel status, mesh.c:1299-1319, and does not choose it
    itself): a node is owned by the LOWEST part id among the parts whose elements touch it.  Dirichlet nodes are never
    ghosted: their values are data every rank can evaluate, so each rank keeps them in its own closure."""
    ne = tuple(int(v) for v in ne)
    Nx, Ny, Nz = ne
    M, N, P = procs
    nranks = M * N * P
    rI, rJ, rK = rank // (N * P), (rank // P) % N, rank % P
    ex0, ex1 = part_range(Nx, M, rI)
    ey0, ey1 = part_range(Ny, N, rJ)
    ez0, ez1 = part_range(Nz, P, rK)
    lne = (ex1 - ex0, ey1 - ey0, ez1 - ez0)
    nelem = lne[0] * lne[1] * lne[2]
    d = p + 1
    classify, nent = _entity_tables(ne, p, set(dirichlet))

    # ---- all nodes of the local node box, lexicographic -------------------------------------------------------------
    gx0, gy0, gz0 = ex0 * p, ey0 * p, ez0 * p
    nx, ny, nz = lne[0] * p + 1, lne[1] * p + 1, lne[2] * p + 1
    gx = (gx0 + np.arange(nx, dtype=np.int64))[:, None, None] + np.zeros((1, ny, nz), dtype=np.int64)
    gy = (gy0 + np.arange(ny, dtype=np.int64))[None, :, None] + np.zeros((nx, 1, nz), dtype=np.int64)
    gz = (gz0 + np.arange(nz, dtype=np.int64))[None, None, :] + np.zeros((nx, ny, 1), dtype=np.int64)
    gx, gy, gz = gx.ravel(), gy.ravel(), gz.ravel()
    etype, ent, within, diri = classify(gx, gy, gz)

    # ---- owner part of every node: lowest part id among touching elements ---------------------------------------------
    def lowest_part(g, nel, nparts):
        # elements touching node coordinate g in this direction: floor((g-1)/p) and floor(g/p), clipped
        elo = np.clip((g - 1) // p, 0, nel - 1)
        return part_of_element(elo, nel, nparts)
    owner = (lowest_part(gx, Nx, M) * N + lowest_part(gy, Ny, N)) * P + lowest_part(gz, Nz, P)
    is_ghost = (owner != rank) & ~diri
    cat = np.where(diri, 1, np.where(is_ghost, 2, 0)).astype(np.int8)      # 0 explicit-owned, 1 dirichlet, 2 ghost

    # ---- local numbering: sort by (category, [type, entity, within]) -------------------------------------------------
    pm = max(p - 1, 1)
    if ordering == "natural":
        ent_key = ent
        type_key = etype.astype(np.int64)
    elif ordering == "element":
        # entities ordered by the lexicographically first element touching them; type is a tie-break inside that element
        I = np.clip((gx - (gx % p == 0)) // p, 0, Nx - 1)
        J = np.clip((gy - (gy % p == 0)) // p, 0, Ny - 1)
        K = np.clip((gz - (gz % p == 0)) // p, 0, Nz - 1)
        first_elem = (I * Ny + J) * Nz + K
        ent_key = first_elem * (8 * max(nent)) + ent        # unique per (element, entity)
        type_key = np.zeros_like(ent)
        ent_key = ent_key * 4 + etype
    else:
        raise ValueError("unknown ordering %r" % (ordering,))
    nnodes = gx.size
    order = np.lexsort((within, ent_key, type_key, cat))
    local_of_lex = np.empty(nnodes, dtype=np.int64)
    local_of_lex[order] = np.arange(nnodes)
    n_owned = int(np.count_nonzero(cat == 0))
    n_diri = int(np.count_nonzero(cat == 1))
    n_ghost = nnodes - n_owned - n_diri
    nc = n_owned + n_diri

    # ---- global numbering: rstart = sum of owned counts on lower ranks (cont.c:326-336) --------------------------------
    counts = np.zeros(nranks, dtype=np.int64)
    if nranks == 1:
        counts[0] = n_owned
    else:
        for r in range(nranks):
            counts[r] = _count_owned(ne, p, set(dirichlet), procs, r)
    rstart = int(counts[:rank].sum())
    nglobal = int(counts.sum())

    # ---- e2l -------------------------------------------------------------------------------------------------------------
    ei, ej, ek = np.meshgrid(np.arange(lne[0]), np.arange(lne[1]), np.arange(lne[2]), indexing="ij")
    a = np.arange(d)
    lx = (ei.ravel()[:, None, None, None] * p + a[None, :, None, None])
    ly = (ej.ravel()[:, None, None, None] * p + a[None, None, :, None])
    lz = (ek.ravel()[:, None, None, None] * p + a[None, None, None, :])
    lex = (lx * ny + ly) * nz + lz
    e2l = local_of_lex[lex.reshape(-1)].astype(np.int32)
    elem_ijk = np.stack([ei.ravel() + ex0, ej.ravel() + ey0, ek.ravel() + ez0], axis=1).astype(np.int32)

    # ---- geometry: vertex grid -> expanded Q1 coordinates in tensor order (i,j,k), k fastest (fsgeom.c:150-152) ----------
    (x0, x1), (y0, y1), (z0, z1) = box
    vx = x0 + (x1 - x0) * (np.arange(ex0, ex1 + 1) / Nx)      # genblock.c:219-221: 1.*i/(m-1)
    vy = y0 + (y1 - y0) * (np.arange(ey0, ey1 + 1) / Ny)
    vz = z0 + (z1 - z0) * (np.arange(ez0, ez1 + 1) / Nz)
    V = np.stack(np.meshgrid(vx, vy, vz, indexing="ij"), axis=-1)
    if morph:
        V = morph_coords(V, **morph)
    coords = np.empty((nelem, 8, 3))
    for ci, (di, dj, dk) in enumerate([(i, j, k) for i in (0, 1) for j in (0, 1) for k in (0, 1)]):
        coords[:, ci, :] = V[ei.ravel() + di, ej.ravel() + dj, ek.ravel() + dk]

    # ---- nodal coordinates: Q1 geometry interpolated to the GLL nodes (fsgeom.c:20-106) ------------------------------------
    g1 = _gll(d)
    N0, N1 = 0.5 * (1 - g1), 0.5 * (1 + g1)
    W = np.stack([N0, N1], axis=0)                             # [corner, node]
    xe = np.einsum("ia,jb,kc,eijkx->eabcx", W, W, W, coords.reshape(nelem, 2, 2, 2, 3))
    node_xyz = np.zeros((nnodes, 3))
    node_xyz[e2l] = xe.reshape(-1, 3)                          # all elements agree on shared nodes to rounding

    # ---- l2g ---------------------------------------------------------------------------------------------------------------
    l2g = np.full(nnodes, -1, dtype=np.int64)
    l2g[:n_owned] = rstart + np.arange(n_owned)
    halos = []
    if n_ghost:
        glob_lex = (gx * (Ny * p + 1) + gy) * (Nz * p + 1) + gz   # global lexicographic id
        ghost_lex_idx = np.nonzero(cat == 2)[0]
        for r in np.unique(owner[ghost_lex_idx]):
            sel = ghost_lex_idx[owner[ghost_lex_idx] == r]
            sel = sel[np.argsort(glob_lex[sel])]
            remote_local, remote_rstart = _owned_local_index(ne, p, set(dirichlet), procs, int(r), gx[sel], gy[sel], gz[sel], ordering, counts)
            l2g[local_of_lex[sel]] = remote_rstart + remote_local
            halos.append(Halo(int(r), np.zeros(0, dtype=np.int32), local_of_lex[sel].astype(np.int32)))
    if nranks > 1:
        # nodes I own that other ranks ghost: by symmetry of the rule, computed from the neighbours' node boxes
        halos = _complete_halos(ne, p, set(dirichlet), procs, rank, halos, gx, gy, gz, cat, local_of_lex, ordering)

    return BoxSpace(ne=ne, p=p, bs=bs, nelem=nelem, n=n_owned, nc=nc, ngh=n_ghost, rstart=rstart, nglobal=nglobal,
                    e2l=e2l, l2g=l2g.astype(np.int32), coords=coords, node_xyz=node_xyz,
                    halos=halos, elem_ijk=elem_ijk, rank=rank, nranks=nranks)


def _inverse(perm):
    inv = np.empty_like(perm)
    inv[perm] = np.arange(perm.size)
    return inv


def _rank_box(ne, procs, r):
    M, N, P = procs
    rI, rJ, rK = r // (N * P), (r // P) % N, r % P
    return part_range(ne[0], M, rI), part_range(ne[1], N, rJ), part_range(ne[2], P, rK)


def _owner_of(ne, p, procs, gx, gy, gz):
    M, N, P = procs

    def low(g, nel, nparts):
        return part_of_element(np.clip((g - 1) // p, 0, nel - 1), nel, nparts)
    return (low(gx, ne[0], M) * N + low(gy, ne[1], N)) * P + low(gz, ne[2], P)


def _rank_nodes(ne, p, diri_sets, procs, r):
    (ex0, ex1), (ey0, ey1), (ez0, ez1) = _rank_box(ne, procs, r)
    nx, ny, nz = (ex1 - ex0) * p + 1, (ey1 - ey0) * p + 1, (ez1 - ez0) * p + 1
    gx = (ex0 * p + np.arange(nx, dtype=np.int64))[:, None, None] + np.zeros((1, ny, nz), dtype=np.int64)
    gy = (ey0 * p + np.arange(ny, dtype=np.int64))[None, :, None] + np.zeros((nx, 1, nz), dtype=np.int64)
    gz = (ez0 * p + np.arange(nz, dtype=np.int64))[None, None, :] + np.zeros((nx, ny, 1), dtype=np.int64)
    return gx.ravel(), gy.ravel(), gz.ravel()


def _count_owned(ne, p, diri_sets, procs, r):
    gx, gy, gz = _rank_nodes(ne, p, diri_sets, procs, r)
    classify, _ = _entity_tables(ne, p, diri_sets)
    _, _, _, diri = classify(gx, gy, gz)
    owner = _owner_of(ne, p, procs, gx, gy, gz)
    return int(np.count_nonzero((owner == r) & ~diri))


def _local_numbering(ne, p, diri_sets, procs, r, ordering):
    """(gx,gy,gz,cat,local_of_lex) of rank r -- the same computation build_box_space does for its own rank."""
    gx, gy, gz = _rank_nodes(ne, p, diri_sets, procs, r)
    classify, nent = _entity_tables(ne, p, diri_sets)
    etype, ent, within, diri = classify(gx, gy, gz)
    owner = _owner_of(ne, p, procs, gx, gy, gz)
    cat = np.where(diri, 1, np.where(owner != r, 2, 0)).astype(np.int8)
    if ordering == "natural":
        ent_key, type_key = ent, etype.astype(np.int64)
    else:
        Nx, Ny, Nz = ne
        I = np.clip((gx - (gx % p == 0)) // p, 0, Nx - 1)
        J = np.clip((gy - (gy % p == 0)) // p, 0, Ny - 1)
        K = np.clip((gz - (gz % p == 0)) // p, 0, Nz - 1)
        ent_key = (((I * Ny + J) * Nz + K) * (8 * max(nent)) + ent) * 4 + etype
        type_key = np.zeros_like(ent)
    order = np.lexsort((within, ent_key, type_key, cat))
    local_of_lex = np.empty(gx.size, dtype=np.int64)
    local_of_lex[order] = np.arange(gx.size)
    return gx, gy, gz, cat, local_of_lex


def _owned_local_index(ne, p, diri_sets, procs, r, qx, qy, qz, ordering, counts):
    """Local (= global - rstart) index on rank r of the nodes (qx,qy,qz) that rank r owns."""
    gx, gy, gz, cat, local_of_lex = _local_numbering(ne, p, diri_sets, procs, r, ordering)
    (ex0, ex1), (ey0, ey1), (ez0, ez1) = _rank_box(ne, procs, r)
    ny, nz = (ey1 - ey0) * p + 1, (ez1 - ez0) * p + 1
    lex = ((qx - ex0 * p) * ny + (qy - ey0 * p)) * nz + (qz - ez0 * p)
    assert np.all(cat[lex] == 0)
    return local_of_lex[lex], int(counts[:r].sum())


def _complete_halos(ne, p, diri_sets, procs, rank, halos, gx, gy, gz, cat, local_of_lex, ordering):
    """Fill the ``send`` side: for every other rank r whose node box overlaps mine, the nodes I own that r ghosts."""
    nranks = procs[0] * procs[1] * procs[2]
    Ny, Nz = ne[1], ne[2]
    by_rank: Dict[int, Halo] = {h.rank: h for h in halos}
    (ax0, ax1), (ay0, ay1), (az0, az1) = _rank_box(ne, procs, rank)
    glob_lex = (gx * (Ny * p + 1) + gy) * (Nz * p + 1) + gz
    for r in range(nranks):
        if r == rank:
            continue
        (bx0, bx1), (by0, by1), (bz0, bz1) = _rank_box(ne, procs, r)
        # r ghosts my owned nodes lying in its node box
        inbox = (gx >= bx0 * p) & (gx <= bx1 * p) & (gy >= by0 * p) & (gy <= by1 * p) & (gz >= bz0 * p) & (gz <= bz1 * p)
        sel = np.nonzero(inbox & (cat == 0))[0]
        if sel.size == 0:
            continue
        sel = sel[np.argsort(glob_lex[sel])]
        h = by_rank.get(r)
        if h is None:
            h = Halo(r, np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int32))
            by_rank[r] = h
        h.send = local_of_lex[sel].astype(np.int32)
    return [by_rank[r] for r in sorted(by_rank)]


# ---- Q1 preconditioning-matrix pattern --------------------------------------------------------------------------------------

def q1_pattern(space: BoxSpace, bs: Optional[int] = None):
    """CSR pattern of the embedded-Q1 matrix exactly as PETSc AIJ holds it after
    ``MatSetValuesBlockedLocal(ADD_VALUES)`` of every patch matrix and final assembly (dFSMatSetValuesBlockedExpanded,
    fs.c:826-930; ellip.c:521,528-529): per (scalar) row the sorted union of the columns of all patches containing the
    row's node -- zeros are kept -- with Dirichlet rows and columns dropped (negative global index, fsbuild.c:31-34).

    Generic union-of-patches construction (works for any e2l); rows are the rows owned by this rank, columns global."""
    bs = space.bs if bs is None else bs
    p, d = space.p, space.p + 1
    # the 8 element-local corner indices of every Q1 patch (efstopo.c:676-684)
    i, j, k = np.meshgrid(np.arange(p), np.arange(p), np.arange(p), indexing="ij")
    corners = [((i + a) * d + (j + b)) * d + (k + c) for a in (0, 1) for b in (0, 1) for c in (0, 1)]
    corners = np.stack([c.ravel() for c in corners], axis=1)                    # [p^3, 8]
    e2l = space.e2l.reshape(space.nelem, d ** 3)
    pn = space.l2g[e2l[:, corners]]                                             # [nelem, p^3, 8] global node or <0
    pn = pn.reshape(-1, 8).astype(np.int64)
    rows = np.repeat(pn, 8, axis=1).ravel()
    cols = np.tile(pn, (1, 8)).ravel()
    keep = (rows >= 0) & (cols >= 0)
    rows, cols = rows[keep], cols[keep]
    own = (rows >= space.rstart) & (rows < space.rstart + space.n)
    rows, cols = rows[own] - space.rstart, cols[own]
    key = np.unique(rows * (space.nglobal + 1) + cols)
    brow, bcol = key // (space.nglobal + 1), key % (space.nglobal + 1)
    bcount = np.bincount(brow, minlength=space.n)
    if bs == 1:
        rowptr = np.concatenate(([0], np.cumsum(bcount))).astype(np.int32)
        return rowptr, bcol.astype(np.int32)
    # expand blocks: scalar row r*bs+fi has columns c*bs+fj for every block column c, fj
    counts = np.repeat(bcount * bs, bs)
    rowptr = np.concatenate(([0], np.cumsum(counts))).astype(np.int64)
    browptr = np.concatenate(([0], np.cumsum(bcount)))
    colidx = np.empty(int(rowptr[-1]), dtype=np.int32)
    sc = (bcol[:, None] * bs + np.arange(bs)[None, :])                          # [nblocks, bs]
    for fi in range(bs):
        # rows r*bs+fi: concatenation of sc rows of block-row r
        dst0 = rowptr[fi::bs][:-1] if False else rowptr[np.arange(space.n) * bs + fi]
        lens = bcount * bs
        idx = np.repeat(dst0, lens) + (np.arange(int(lens.sum())) - np.repeat(np.cumsum(lens) - lens, lens))
        colidx[idx] = sc.ravel()
    assert rowptr[-1] < 2 ** 31
    return rowptr.astype(np.int32), colidx


def dense_pattern(space):
    """CSR pattern of the element-dense coupling (dEFSGetExplicitSparse_Basic path, jacobi.c:611-643): every pair of nodes
    sharing an element; rows owned by this rank, columns global, Dirichlet rows/columns dropped, sorted."""
    d3 = (space.p + 1) ** 3
    g = space.l2g[space.e2l.reshape(space.nelem, d3)].astype(np.int64)
    rows, cols = np.repeat(g, d3, axis=1).ravel(), np.tile(g, (1, d3)).ravel()
    keep = (rows >= space.rstart) & (rows < space.rstart + space.n) & (cols >= 0)
    key = np.unique((rows[keep] - space.rstart) * (space.nglobal + 1) + cols[keep])
    r, c = key // (space.nglobal + 1), key % (space.nglobal + 1)
    rowptr = np.concatenate(([0], np.cumsum(np.bincount(r, minlength=space.n)))).astype(np.int32)
    return rowptr, c.astype(np.int32)
