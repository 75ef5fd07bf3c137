"""The reference's closure test (src/vec/tests/vec-ex1.c, 2 ranks, golden src/vec/tests/refout/vec-ex1.refout) replayed on dohp_b200.vecd
over gloo: VecDohp's three windows (global form | closure | local form, src/vec/vecd.c:123-151) and the forward ghost update, printed
as PETSc's VecView prints them, must equal the golden file byte for byte.  Pins the layout SURVEY.md 8a A11 names with the
reference's own known answer.  CPU only."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from dohp_b200 import vecd

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refout", "vec-ex1.refout")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _view(kind, parts):
    """VecView of a parallel vector on the ASCII stdout viewer: header, then every rank's values with %g."""
    out = ["Vector Object: %d MPI processes" % len(parts), "  type: " + kind]
    for r, vals in enumerate(parts):
        out.append("Process [%d]" % r)
        out += ["%g" % v for v in vals]
    return out


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lines = []

        def show(kind, mine):                   # collective: every rank contributes its window
            parts = [None] * world
            dist.all_gather_object(parts, np.array(mine).tolist())
            lines.extend(_view(kind, parts))

        n, bs = 3, 2                             # vec-ex1.c:16
        ghosts = [n * ((world + rank - 1) % world) + 1, n * ((world + rank + 1) % world)]       # :26-27
        said = [None] * world
        dist.all_gather_object(said, "[%d] ghosts %d %d" % (rank, ghosts[0], ghosts[1]))
        lines += said
        x = vecd.VecDohp(bs, n - 1, n, ghosts, dist)                                             # VecCreateDohp(comm,bs,n-1,n,nghost,ghosts,&x), :31
        assert x.global_form.size == (n - 1) * bs
        lines.append("Empty vector")
        show("dohp", x.global_form)
        x.global_form[:] = 1000 + rank * 100 + np.arange(x.global_form.size)                    # :41-43
        lines.append("Global unclosed form")
        show("dohp", x.global_form)
        lines.append("Global closed form")
        show("mpi", x.closure)
        lines.append("Before VecGhostUpdateBegin/End")
        show("mpi", x.local_form)
        x.ghost_update(vecd.INSERT_VALUES, vecd.SCATTER_FORWARD)                                 # :64-65
        lines.append("After VecGhostUpdateBegin/End")
        show("mpi", x.local_form)
        y = x.duplicate()                                                                        # :75-76
        y.global_form[:] = x.global_form
        lines.append("Global unclosed form of y")
        show("dohp", y.global_form)
        lines.append("Global closed form of y")
        show("mpi", y.closure)
        y.closure[:] = 0                                                                         # VecZeroEntries(yc); VecShift(xc,1000); VecCopy(xc,yc), :82-84
        x.closure[:] += 1000
        y.closure[:] = x.closure
        lines.append("Global closed form of y, after copy of shifted closure")
        show("mpi", y.closure)
        assert np.all(x.local_form[n * bs:] < 2000)          # the shift touched the closure, not the ghosts
        if rank == 0:
            ret["text"] = "\n".join(lines) + "\n"
            ret["halos"] = [(r, s.tolist(), v.tolist()) for r, s, v in x.halos]
    finally:
        dist.destroy_process_group()


def test_vec_ex1_golden():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert ret["text"] == open(GOLDEN).read()
    assert ret["halos"] == [(1, [1, 0], [3, 4])]            # rank 1 ghosts my blocks 1 then 0; my ghosts land in the slots behind the closure


def test_halos_from_ghosts_lists_agree_pairwise():
    """Both sides of every pair list the shared blocks in the same order (what dB200SpaceSetHalo requires), on an irregular layout."""
    rng = np.random.default_rng(7)
    nc_all = [5, 0, 7, 3]
    cstart = vecd.closure_starts(nc_all)
    ghosts_all = []
    for r in range(4):
        others = np.array([g for g in range(cstart[-1]) if not cstart[r] <= g < cstart[r + 1]])
        ghosts_all.append(rng.permutation(others)[:6].tolist())
    halos = [vecd.halos_from_ghosts(nc_all, ghosts_all, r) for r in range(4)]
    for r in range(4):
        assert [h[0] for h in halos[r]] == sorted(h[0] for h in halos[r])
        got = np.concatenate([h[2] for h in halos[r]] or [np.zeros(0, int)])
        assert sorted(got.tolist()) == list(range(nc_all[r], nc_all[r] + 6))                   # every ghost slot is filled exactly once
        for q, send, recv in halos[r]:
            back = [h for h in halos[q] if h[0] == r][0]
            assert len(send) == len(back[2]) and len(recv) == len(back[1])
            # what q receives from r, slot by slot, is the block q asked for
            for s, slot in zip(send, back[2]):
                assert cstart[r] + s == ghosts_all[q][slot - nc_all[q]]
    with pytest.raises(ValueError):
        vecd.halos_from_ghosts(nc_all, [[0], [], [], []], 0)                                    # own block as a ghost
    with pytest.raises(ValueError):
        vecd.halos_from_ghosts(nc_all, [[99], [], [], []], 0)                                   # outside the layout


@pytest.mark.parametrize("procs,ne,p", [((2, 1, 1), (4, 2, 2), 3), ((2, 2, 1), (4, 4, 2), 2), ((2, 2, 2), (4, 4, 4), 2), ((1, 3, 2), (2, 5, 4), 4)])
def test_box_partition_halo_lists_are_what_veccreatedohp_would_build(procs, ne, p):
    """Cross-check of two independent statements of the ghost scatter: the neighbour lists dB200BoxMeshCreate (csrc/boxmesh.cpp) hands to
    dB200SpaceSetHalo, and vecd.halos_from_ghosts applied to the ghost list a VecCreateDohp caller would pass (the ghosts' block indices in
    the closure numbering: owner's closure start + the owner's local index).  Same neighbours, same node sets, and the order both sides of a
    pair use is the same."""
    from dohp_b200 import lib
    lib.build()
    world = procs[0] * procs[1] * procs[2]
    meshes = [lib.BoxMesh(ne, p, procs=procs, rank=r, morph=0.2, twist=0.1) for r in range(world)]
    nc_all = [m.nc for m in meshes]
    cstart = vecd.closure_starts(nc_all)
    rstart = np.array([m.rstart for m in meshes] + [meshes[0].nglobal])
    ghosts_all = []
    for m in meshes:
        g = m.l2g[m.nc:].astype(np.int64)                               # global EXPLICIT index of every ghost node
        assert np.all(g >= 0)
        owner = np.searchsorted(rstart, g, side="right") - 1
        ghosts_all.append((cstart[owner] + (g - rstart[owner])).tolist())   # explicit nodes come first in the owner's closure
    for r, m in enumerate(meshes):
        mine = {q: (s, v) for q, s, v in vecd.halos_from_ghosts(nc_all, ghosts_all, r)}
        theirs = {q: (np.asarray(s), np.asarray(v)) for q, s, v in m.halos}
        assert sorted(mine) == sorted(theirs)
        for q in mine:
            assert sorted(mine[q][0].tolist()) == sorted(theirs[q][0].tolist())          # nodes r sends to q
            assert sorted(mine[q][1].tolist()) == sorted(theirs[q][1].tolist())          # ghost slots r fills from q
            # pairing: the k-th entry r sends lands in the k-th slot q receives, and it is the same physical node
            back = {qq: (np.asarray(s), np.asarray(v)) for qq, s, v in meshes[q].halos}[r]
            assert len(theirs[q][0]) == len(back[1])
            assert np.allclose(m.node_xyz[theirs[q][0]], meshes[q].node_xyz[back[1]], rtol=0, atol=1e-12)
