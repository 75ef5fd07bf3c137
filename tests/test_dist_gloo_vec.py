"""N > 1 host logic on CPU for the vector-valued operators (gloo, world_size 2 and 4): elast (one space, bs = 3) and stokes (TWO spaces,
two ghost exchanges each way per product: MatMult_Nest_StokesCoupled, src/fs/tests/stokes.c:541-577 with fs.c:578-579, 627-628) on the
block partition and halo lists of dB200BoxMeshCreate, the ORACLE doing the local apply, against the single-domain oracle.  Companion of
tests/test_dist_gloo.py (ellip MatMult and Q1 rows); the same exchanges run on the GPUs in tests/partition_parity.py / mgpu_check.py."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as ol
from dist_common import field, key
from dohp_b200 import lib

MORPH = dict(morph=0.2, twist=0.15)
COMP = np.array([1.0, -0.7, 0.4])
STOKES = dict(exact=1, p=1.2, eps=0.1)                 # rheology of stokes-q3q2-fgmres (CMakeLists.txt:65-70): the stash matters


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def vfield(xyz, amp=1.0):
    return amp * (field(xyz)[:, None] * COMP[None, :]).reshape(-1)


def pfield(xyz):
    return np.cos(0.9 * xyz[:, 0] - 0.3) * xyz[:, 1] + 0.2 * xyz[:, 2]


def forward(mesh, bs, xg):
    """global (owned explicit) -> local with the ghosts filled: VecGhostUpdate(INSERT, FORWARD); Dirichlet part left zero"""
    loc = np.zeros((mesh.nloc, bs))
    loc[:mesh.n] = xg.reshape(-1, bs)
    reqs, land = [], []
    for r, snd, rcv in mesh.halos:
        if len(snd):
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(loc[snd])), r))
        if len(rcv):
            b = torch.empty((len(rcv), bs), dtype=torch.float64)
            land.append((rcv, b))
            reqs.append(dist.irecv(b, r))
    for q in reqs:
        q.wait()
    for rcv, b in land:
        loc[rcv] = b.numpy()
    return loc.reshape(-1)


def reverse(mesh, bs, yloc):
    """local -> global: VecGhostUpdate(ADD, REVERSE), owned explicit part returned"""
    y = yloc.reshape(-1, bs).copy()
    reqs, land = [], []
    for r, snd, rcv in mesh.halos:
        if len(rcv):
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(y[rcv])), r))
        if len(snd):
            b = torch.empty((len(snd), bs), dtype=torch.float64)
            land.append((snd, b))
            reqs.append(dist.irecv(b, r))
    for q in reqs:
        q.wait()
    for snd, b in land:
        np.add.at(y, snd, b.numpy())
    return y[:mesh.n].reshape(-1)


def _elast_worker(rank, world, port, ne, p, procs, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mesh = lib.BoxMesh(ne, p, procs=procs, rank=rank, **MORPH)
        o = ol.OracleElast(mesh)
        orc = ol.oracle()
        xyz = mesh.node_xyz[:mesh.n]
        xl = forward(mesh, 3, vfield(xyz, 0.05))                                         # state (homogeneous Dirichlet data), ghosts exchanged
        yl = np.zeros(mesh.nloc * 3)
        assert orc.or_elast_function(o.cefs.h, o.efs.h, C.byref(o.s), ol.dptr(o.coords), C.byref(o.prm), ol.dptr(xl), ol.dptr(yl), ol.dptr(o.stash)) == 0
        F = reverse(mesh, 3, yl)
        xl = forward(mesh, 3, vfield(xyz))
        yl = np.zeros(mesh.nloc * 3)
        assert orc.or_elast_mult(o.cefs.h, o.efs.h, C.byref(o.s), ol.dptr(o.coords), C.byref(o.prm), ol.dptr(o.stash), ol.dptr(xl), ol.dptr(yl)) == 0
        y = reverse(mesh, 3, yl)
        out = [None] * world
        dist.all_gather_object(out, (key(xyz), F.reshape(-1, 3), y.reshape(-1, 3), len(mesh.halos)))
        if rank == 0:
            ret["out"] = out
    finally:
        dist.destroy_process_group()


def _stokes_worker(rank, world, port, ne, p, procs, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mu = lib.BoxMesh(ne, p, procs=procs, rank=rank, **MORPH)
        mpr = lib.BoxMesh(ne, p - 1, dirichlet=0, procs=procs, rank=rank, **MORPH)      # -pressure_codim 1, no Dirichlet pressure nodes
        o = ol.OracleStokes(mu, mpr, **STOKES)
        orc = ol.oracle()
        xu, xp = mu.node_xyz[:mu.n], mpr.node_xyz[:mpr.n]

        def apply(fn, u, pr, *extra):                                                   # two forward and two reverse exchanges per product
            ul, pl = forward(mu, 3, u), forward(mpr, 1, pr)
            yu, yp = np.zeros(mu.nloc * 3), np.zeros(mpr.nloc)
            head = (o.cefs.h, o.efsu.h, o.efsp.h, C.byref(o.osu), C.byref(o.osp), ol.dptr(o.coords), C.byref(o.prm))
            if fn == "function":
                err = orc.or_stokes_function(*head, ol.dptr(ul), ol.dptr(pl), ol.dptr(yu), ol.dptr(yp), ol.dptr(o.stash))
            else:
                err = orc.or_stokes_mult(*head, ol.dptr(o.stash), extra[0], ol.dptr(ul), ol.dptr(pl), ol.dptr(yu), ol.dptr(yp))
            assert err == 0
            return reverse(mu, 3, yu), reverse(mpr, 1, yp)

        Fu, Fp = apply("function", vfield(xu, 0.1), 0.1 * pfield(xp))
        Ju, Jp = apply("mult", vfield(xu), pfield(xp), 0)                               # coupled Jacobian with the stash F left
        out = [None] * world
        dist.all_gather_object(out, (key(xu), key(xp), Fu.reshape(-1, 3), Fp, Ju.reshape(-1, 3), Jp, (len(mu.halos), len(mpr.halos))))
        if rank == 0:
            ret["out"] = out
    finally:
        dist.destroy_process_group()


def _spawn(worker, world, *args):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(worker, args=(world, _free_port()) + args + (ret,), nprocs=world, join=True)
    return ret["out"]


def _merge(pairs, ref_keys, ref_vals, what):
    """every global node appears on exactly one rank and carries the single-domain value"""
    ref = dict(zip(ref_keys, ref_vals))
    scale = np.abs(np.asarray(ref_vals)).max()
    seen = set()
    for keys, vals in pairs:
        for k, v in zip(keys, vals):
            assert k in ref and k not in seen, what
            seen.add(k)
            assert np.abs(v - ref[k]).max() <= 1e-12 * scale, (what, k)
    assert len(seen) == len(ref), what


@pytest.mark.parametrize("procs,ne,p", [((2, 1, 1), (4, 2, 2), 2), ((1, 2, 2), (2, 4, 3), 3)])
def test_partitioned_elast_equals_single_domain(procs, ne, p):
    ol.build_oracle(ref=False)
    lib.build()
    mesh = lib.BoxMesh(ne, p, **MORPH)
    o = ol.OracleElast(mesh)
    xyz = mesh.node_xyz[:mesh.n]
    F = o.function(vfield(xyz, 0.05), dcache=np.zeros((mesh.nc - mesh.n) * 3))
    y = o.mult(vfield(xyz))
    out = _spawn(_elast_worker, procs[0] * procs[1] * procs[2], ne, p, procs)
    assert all(r[3] > 0 for r in out)                                                   # every rank really has neighbours
    _merge([(r[0], r[1]) for r in out], key(xyz), F.reshape(-1, 3), "elast F")
    _merge([(r[0], r[2]) for r in out], key(xyz), y.reshape(-1, 3), "elast J x")


@pytest.mark.parametrize("procs,ne,p", [((2, 1, 1), (4, 2, 2), 3), ((2, 2, 1), (4, 4, 2), 2)])
def test_partitioned_stokes_two_spaces_equals_single_domain(procs, ne, p):
    ol.build_oracle(ref=False)
    lib.build()
    mu, mpr = lib.BoxMesh(ne, p, **MORPH), lib.BoxMesh(ne, p - 1, dirichlet=0, **MORPH)
    o = ol.OracleStokes(mu, mpr, **STOKES)
    xu, xp = mu.node_xyz[:mu.n], mpr.node_xyz[:mpr.n]
    Fu, Fp = o.function(vfield(xu, 0.1), 0.1 * pfield(xp), dcache=np.zeros((mu.nc - mu.n) * 3))
    Ju, Jp = o.mult(vfield(xu), pfield(xp))
    out = _spawn(_stokes_worker, procs[0] * procs[1] * procs[2], ne, p, procs)
    assert all(r[6][0] > 0 and r[6][1] > 0 for r in out)                                # both spaces exchange on every rank
    _merge([(r[0], r[2]) for r in out], key(xu), Fu.reshape(-1, 3), "stokes F_u")
    _merge([(r[1], r[3]) for r in out], key(xp), Fp, "stokes F_p")
    _merge([(r[0], r[4]) for r in out], key(xu), Ju.reshape(-1, 3), "stokes (J x)_u")
    _merge([(r[1], r[5]) for r in out], key(xp), Jp, "stokes (J x)_p")
