"""N > 1 host logic on CPU (gloo, world_size 2 and 4): the block partition, ownership rule and halo lists produced by
dB200BoxMeshCreate drive a forward INSERT / reverse ADD ghost exchange (fs.c:578-579, 627-628) around the ORACLE's
local apply; the assembled result must equal the single-domain oracle apply.  This is the same exchange the library
performs with NCCL on the GPUs (tests/mgpu_check.py covers that path on real devices)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as ol
from dist_common import field, key, single_domain_q1, single_domain_reference
from dohp_b200 import lib


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ne, p, procs, morph, prm, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mesh = lib.BoxMesh(ne, p, procs=procs, rank=rank, **morph)
        o = ol.OracleEllip(mesh, **prm)
        nloc = mesh.nloc

        def forward(xg):                       # VecGhostUpdate INSERT FORWARD
            loc = np.zeros(nloc)
            loc[:mesh.n] = xg
            reqs, bufs = [], []
            for r, snd, rcv in mesh.halos:
                if len(snd):
                    reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(loc[snd])), r))
                if len(rcv):
                    b = torch.empty(len(rcv), dtype=torch.float64)
                    bufs.append((rcv, b))
                    reqs.append(dist.irecv(b, r))
            for q in reqs:
                q.wait()
            for rcv, b in bufs:
                loc[rcv] = b.numpy()
            return loc

        def reverse(yloc):                     # VecGhostUpdate ADD REVERSE
            reqs, bufs = [], []
            for r, snd, rcv in mesh.halos:
                if len(rcv):
                    reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(yloc[rcv])), r))
                if len(snd):
                    b = torch.empty(len(snd), dtype=torch.float64)
                    bufs.append((snd, b))
                    reqs.append(dist.irecv(b, r))
            for q in reqs:
                q.wait()
            for snd, b in bufs:
                np.add.at(yloc, snd, b.numpy())
            return yloc[:mesh.n]

        x = field(mesh.node_xyz[:mesh.n])
        xl = forward(0.3 * x)
        o.function_local(xl)                   # stash at the same state as the single-domain run
        y = reverse(o.mult_local(forward(x)))
        out = [None] * world
        dist.all_gather_object(out, (key(mesh.node_xyz[:mesh.n]), y, mesh.n, mesh.rstart, mesh.nglobal))
        if rank == 0:
            ret["out"] = out
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("procs,ne,p", [((2, 1, 1), (4, 2, 2), 3), ((1, 2, 2), (2, 4, 3), 2), ((2, 1, 1), (3, 2, 2), 4)])
def test_partitioned_apply_equals_single_domain(procs, ne, p):
    ol.build_oracle(ref=False)
    lib.build()
    morph = dict(morph=0.2, twist=0.15)
    prm = dict(p=1.7, lam=0.3, eps=0.8)
    world = procs[0] * procs[1] * procs[2]
    ref, nglobal = single_domain_reference(ne, p, morph, prm)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ne, p, procs, morph, prm, ret), nprocs=world, join=True)
    out = ret["out"]
    assert sum(o[2] for o in out) == nglobal == out[0][4]
    assert [o[3] for o in out] == list(np.cumsum([0] + [o[2] for o in out[:-1]]))      # rstart = MPI_Scan of owned counts (cont.c:326-336)
    scale = max(abs(v) for v in ref.values())
    seen = set()
    for keys, y, *_ in out:
        for k, v in zip(keys, y):
            assert k in ref and k not in seen
            seen.add(k)
            assert abs(v - ref[k]) <= 1e-12 * scale
    assert len(seen) == nglobal


def _q1_worker(rank, world, port, ne, p, procs, morph, prm, ret):
    """Each rank assembles the Q1 patches of ITS elements with the oracle (owned rows and rows of its ghost nodes: the entries
    MatSetValuesBlockedLocal would stash, fs.c:919), as triplets keyed by node coordinates."""
    import ctypes as C
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mesh = lib.BoxMesh(ne, p, procs=procs, rank=rank, **morph)
        o = ol.OracleEllip(mesh, **prm)
        # a CSR over ALL global rows that holds this rank's couplings (the oracle indexes rows by global number)
        d = p + 1
        i, j, k = np.meshgrid(np.arange(p), np.arange(p), np.arange(p), indexing="ij")
        corners = np.stack([(((i + a) * d + (j + b)) * d + (k + c)).ravel() for a in (0, 1) for b in (0, 1) for c in (0, 1)], axis=1)
        pn = mesh.l2g[mesh.e2l.reshape(mesh.nelem, d ** 3)[:, corners]].reshape(-1, 8).astype(np.int64)
        rows, cols = np.repeat(pn, 8, axis=1).ravel(), np.tile(pn, (1, 8)).ravel()
        keep = (rows >= 0) & (cols >= 0)
        kk = np.unique(rows[keep] * (mesh.nglobal + 1) + cols[keep])
        r, c = kk // (mesh.nglobal + 1), kk % (mesh.nglobal + 1)
        rowptr = np.concatenate(([0], np.cumsum(np.bincount(r, minlength=mesh.nglobal)))).astype(np.int32)
        colidx = c.astype(np.int32)
        # state on every local node incl. ghosts (analytic field: no exchange needed); homogeneous-zero Dirichlet data
        xloc = 0.3 * field(mesh.node_xyz)
        xloc[mesh.n:mesh.nc] = 0.0
        vals = np.zeros(len(colidx))
        err = ol.oracle().or_ellip_jacobian_q1(o.cefs_s.h, o.efs_s.h, C.byref(o.s), ol.dptr(o.coords), C.byref(o.prm), ol.dptr(xloc),
                                               ol.iptr(mesh.l2g), ol.iptr(rowptr), ol.iptr(colidx), ol.dptr(vals))
        assert err == 0
        g2k = {int(g): kx for kx, g in zip(key(mesh.node_xyz), mesh.l2g) if g >= 0}
        rr = np.repeat(np.arange(mesh.nglobal), np.diff(rowptr))
        trip = {(g2k[int(a)], g2k[int(b)]): v for a, b, v in zip(rr, colidx, vals)}
        owned = set(range(mesh.rstart, mesh.rstart + mesh.n))
        noff = sum(1 for a in rr if int(a) not in owned)
        out = [None] * world
        dist.all_gather_object(out, (trip, noff))
        if rank == 0:
            ret["out"] = out
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("procs,ne,p", [((2, 1, 1), (4, 2, 2), 2), ((1, 2, 2), (2, 4, 2), 3)])
def test_partitioned_q1_assembly_sums_to_single_domain(procs, ne, p):
    """Host logic of the partitioned Q1 assembly (what dB200OpAssembleQ1 does per rank on the GPUs, tests/mgpu_check.py): every
    rank adds the patches of its own elements into its owned rows and into rows of its ghost nodes; adding the ranks' entries by
    global (row, column) -- MatAssemblyEnd -- gives the single-domain matrix."""
    ol.build_oracle(ref=False)
    lib.build()
    morph = dict(morph=0.2, twist=0.15)
    prm = dict(p=1.7, lam=0.3, eps=0.8)
    world = procs[0] * procs[1] * procs[2]
    ref = single_domain_q1(ne, p, morph, prm)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_q1_worker, args=(world, _free_port(), ne, p, procs, morph, prm, ret), nprocs=world, join=True)
    tot, noff = {}, 0
    for trip, n_off in ret["out"]:
        noff += n_off
        for kx, v in trip.items():
            tot[kx] = tot.get(kx, 0.0) + v
    assert noff > 0                                   # some entries really are off-process contributions
    scale = max(abs(v) for v in ref.values())
    assert set(kx for kx, v in ref.items() if v != 0.0) <= set(tot)
    assert all(abs(v) <= 1e-12 * scale for kx, v in tot.items() if kx not in ref)
    assert max(abs(tot[kx] - v) for kx, v in ref.items() if kx in tot) <= 1e-12 * scale
