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
from dist_common import field, key, single_domain_reference
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
