"""Host mirror of the reference's VecDohp (src/vec/vecd.c:123-151): the vector layout every operator call of this library assumes.

    VecCreateDohp(comm, bs, n, nc, nghosts, ghosts[])

creates ONE array of (nc + nghosts) * bs doubles per rank seen through three windows:

    global form   a[0 : n*bs]            the owned explicit dofs: what Krylov methods see, what dB200OpMult takes and returns
    closure       a[0 : nc*bs]           + the owned Dirichlet dofs (VecDohpGetClosure)
    local form    a[0 : (nc+ngh)*bs]     + the ghosts (VecGhostGetLocalForm of the closure)

`ghosts` are BLOCK indices in the global numbering of the CLOSURE vector (rank r owns the blocks [cstart_r, cstart_r + nc_r), cstart = the
prefix sum of nc: VecCreateGhostBlock(comm, bs, nc*bs, PETSC_DECIDE, nghosts, ghosts), vecd.c:132).  VecGhostUpdate(INSERT, FORWARD)
copies owner -> ghost, (ADD, REVERSE) adds ghost -> owner (src/fs/interface/fs.c:578-579, 627-628).

`halos_from_ghosts` turns the ghost list into the neighbour lists dB200SpaceSetHalo takes (include/dohp_b200.h): per neighbour the local
nodes to send, in the order the neighbour lists them as ghosts, and the local ghost slots to receive into.  Setup code (numpy); the
exchange itself runs on the GPUs (csrc/p2p.cu, csrc/comm.cu).  `VecDohp` below carries the same windows on the host over any
torch.distributed backend: it is what tests/test_vecd_gloo.py drives through the reference's own closure test
(src/vec/tests/vec-ex1.c, golden vec-ex1.refout)."""
import numpy as np

INSERT_VALUES, ADD_VALUES = 1, 2
SCATTER_FORWARD, SCATTER_REVERSE = 0, 1


def closure_starts(nc_all):
    """cstart[r] = first global closure block of rank r (PetscLayout of the closure vector); cstart[world] = global size."""
    return np.concatenate(([0], np.cumsum(np.asarray(nc_all, dtype=np.int64))))


def halos_from_ghosts(nc_all, ghosts_all, rank):
    """[(neighbour, sendidx, recvidx)] of `rank`, neighbours in increasing rank order.

    nc_all[r]      closure blocks owned by rank r
    ghosts_all[r]  rank r's ghost list (global closure block indices), any order, duplicates allowed as in VecCreateGhostBlock
    recvidx        local block indices nc + j of the ghosts j owned by the neighbour, in ghost order
    sendidx        local block indices (< nc) the neighbour ghosts, in THE NEIGHBOUR'S ghost order: both sides list a shared set alike"""
    cstart = closure_starts(nc_all)
    world = len(nc_all)
    nc = int(nc_all[rank])

    def owner_of(g):
        g = np.asarray(g, dtype=np.int64)
        if g.size and (g.min() < 0 or g.max() >= cstart[-1]):
            raise ValueError("ghost index outside the closure layout [0, %d)" % cstart[-1])
        return np.searchsorted(cstart, g, side="right") - 1

    mine = np.asarray(ghosts_all[rank], dtype=np.int64)
    own = owner_of(mine)
    if np.any(own == rank):
        raise ValueError("rank %d lists one of its own blocks as a ghost" % rank)
    halos = []
    for r in range(world):
        if r == rank:
            continue
        recv = nc + np.nonzero(own == r)[0]
        theirs = np.asarray(ghosts_all[r], dtype=np.int64)
        send = theirs[owner_of(theirs) == rank] - cstart[rank]
        if len(recv) or len(send):
            halos.append((r, send.astype(np.int32), recv.astype(np.int32)))
    return halos


class VecDohp:
    """The three windows over one array, and the ghost update over torch.distributed point-to-point (any backend; `dist` = the
    torch.distributed module of an initialised process group, or None for a single rank)."""

    def __init__(self, bs, n, nc, ghosts, dist=None):
        if not 0 <= n <= nc:
            raise ValueError("need 0 <= n <= nc")
        self.bs, self.n, self.nc, self.ghosts, self.dist = bs, n, nc, np.asarray(ghosts, dtype=np.int64), dist
        self.rank, self.world = (dist.get_rank(), dist.get_world_size()) if dist is not None else (0, 1)
        if dist is not None:
            everyone = [None] * self.world
            dist.all_gather_object(everyone, (nc, self.ghosts.tolist()))
        else:
            everyone = [(nc, self.ghosts.tolist())]
        self.nc_all = [e[0] for e in everyone]
        self.halos = halos_from_ghosts(self.nc_all, [e[1] for e in everyone], self.rank)
        self.a = np.zeros((nc + len(self.ghosts)) * bs)

    # windows (views: writing through one is seen through the others, as in the reference where they share the array)
    @property
    def global_form(self):
        return self.a[:self.n * self.bs]

    @property
    def closure(self):
        return self.a[:self.nc * self.bs]

    @property
    def local_form(self):
        return self.a

    def duplicate(self):
        """VecDuplicate_Dohp (vecd.c:84-121): same layout and ghosts, new storage (values are not copied)."""
        y = object.__new__(VecDohp)
        y.__dict__.update(self.__dict__)
        y.a = np.zeros_like(self.a)
        return y

    def ghost_update(self, imode, smode):
        """VecGhostUpdateBegin/End on the closure: (INSERT_VALUES, SCATTER_FORWARD) or (ADD_VALUES, SCATTER_REVERSE)."""
        import torch
        if (imode, smode) not in ((INSERT_VALUES, SCATTER_FORWARD), (ADD_VALUES, SCATTER_REVERSE)):
            raise ValueError("the closure is updated with INSERT/FORWARD or ADD/REVERSE (fs.c:578-579, 627-628)")
        blocks = self.a.reshape(-1, self.bs)
        fwd = smode == SCATTER_FORWARD
        reqs, landing = [], []
        for r, send, recv in self.halos:
            out, into = (send, recv) if fwd else (recv, send)
            if len(out):
                reqs.append(self.dist.isend(torch.from_numpy(np.ascontiguousarray(blocks[out])), r))
            if len(into):
                buf = torch.empty((len(into), self.bs), dtype=torch.float64)
                landing.append((into, buf))
                reqs.append(self.dist.irecv(buf, r))
        for q in reqs:
            q.wait()
        for into, buf in landing:                     # neighbours in rank order: deterministic sums
            if fwd:
                blocks[into] = buf.numpy()
            else:
                np.add.at(blocks, into, buf.numpy())
