"""Multi-GPU parity check, run under torchrun on N GPUs of one box (not a pytest file: needs one process per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py
Block-partitioned ellip MatMult through the C ABI with the NCCL ghost exchange vs the single-domain oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch
import torch.distributed as dist

import oracle_lib as ol
from dist_common import field, key, single_domain_q1, single_domain_reference
from dohp_b200 import lib

rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lrank)
dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    idt = torch.frombuffer(bytearray(lib.Context.nccl_unique_id()), dtype=torch.uint8).cuda()
dist.broadcast(idt, 0)
ctx = lib.Context(lrank, rank, world, bytes(idt.cpu().numpy().tobytes()), stream=torch.cuda.current_stream().cuda_stream)
procs = {2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[world]
ol.build_oracle(ref=False)
worst = 0.0
for p, ne in ((2, (4, 4, 4)), (3, (4, 2, 4)), (5, (2, 4, 2))):
    morph, prm = dict(morph=0.2, twist=0.15), dict(p=1.7, lam=0.3, eps=0.8)
    ref, nglobal = single_domain_reference(ne, p, morph, prm)
    mesh = lib.BoxMesh(ne, p, procs=procs, rank=rank, **morph)
    rule = lib.Rule.preferred(ctx, p)
    efs = lib.EFS(ctx, rule, p)
    space = lib.Space.from_mesh(ctx, mesh)
    op = lib.Op(ctx, lib.OP_ELLIP, lib.EllipParam(prm["p"], prm["eps"], prm["lam"]), rule, [space], [efs], mesh.coords)
    x = field(mesh.node_xyz[:mesh.n])
    op.function(0.3 * x)                       # stash (homogeneous-zero Dirichlet data, no forcing)
    xd = torch.from_numpy(x).cuda()
    yd = torch.empty_like(xd)
    op.mult(xd, yd)
    ctx.synchronize()
    y = yd.cpu().numpy()
    scale = max(abs(v) for v in ref.values())
    err = max(abs(v - ref[k]) for k, v in zip(key(mesh.node_xyz[:mesh.n]), y)) / scale
    tot = torch.tensor([mesh.n], device="cuda")
    dist.all_reduce(tot)
    assert tot.item() == nglobal, (tot.item(), nglobal)
    worst = max(worst, err)
    print("rank %d/%d p=%d ne=%s owned=%d ghosts=%d neighbours=%d  max rel err vs single-domain oracle %.2e" % (rank, world, p, ne, mesh.n, mesh.ngh, len(mesh.halos), err), flush=True)
assert worst <= 1e-12, worst

# ---- embedded-Q1 assembly on the partitioned space: every rank assembles the patches of its own elements into its owned rows
# and into the rows of its ghost nodes (what MatSetValues stashes for the owner, fs.c:919); summing the ranks' triplets by
# global (row, column) -- MatAssemblyEnd's job -- must give the single-domain oracle matrix.
worstq = 0.0
for kind, p, ne in (("ellip", 2, (4, 4, 4)), ("ellip", 3, (4, 2, 4)), ("elast", 2, (2, 4, 4))):
    morph = dict(morph=0.2, twist=0.15)
    elast = kind == "elast"
    prm = dict(E0=1.0, nu=0.3) if elast else dict(p=1.7, lam=0.3, eps=0.8)
    mesh = lib.BoxMesh(ne, p, procs=procs, rank=rank, **morph)
    rule = lib.Rule.preferred(ctx, p)
    efs = lib.EFS(ctx, rule, p)
    bs = 3 if elast else 1
    space = lib.Space.from_mesh(ctx, mesh, bs=bs)
    if elast:
        o = ol.OracleElast(mesh, **prm)
        op = lib.Op(ctx, lib.OP_ELAST, lib.ElastParam(o.lam, o.mu), rule, [space], [efs], mesh.coords)
        x = 0.05 * np.repeat(field(mesh.node_xyz[:mesh.n]), 3) * np.tile([1.0, -0.7, 0.4], mesh.n)
    else:
        op = lib.Op(ctx, lib.OP_ELLIP, lib.EllipParam(prm["p"], prm["eps"], prm["lam"]), rule, [space], [efs], mesh.coords)
        x = 0.3 * field(mesh.node_xyz[:mesh.n])
    rowptr, colidx = op.q1_pattern(mesh.l2g)
    assert len(rowptr) == mesh.n + mesh.ngh + 1
    vals = op.assemble_q1(x, mesh.l2g, None, None).reshape(len(colidx), bs, bs)
    ctx.synchronize()
    kloc = key(mesh.node_xyz)
    g2k = {int(g): kloc[i] for i, g in enumerate(mesh.l2g) if g >= 0}
    rownode = np.concatenate([np.arange(mesh.n), mesh.nc + np.arange(mesh.ngh)])
    rows = np.repeat(rownode, np.diff(rowptr))
    trip = {}
    for r, c, blk in zip(rows, colidx, vals):
        for fi in range(bs):
            for fj in range(bs):
                kk = (kloc[r], g2k[int(c)], fi, fj) if elast else (kloc[r], g2k[int(c)])
                trip[kk] = trip.get(kk, 0.0) + blk[fi, fj]
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(trip, gathered, 0)
    if rank == 0:
        ref = single_domain_q1(ne, p, morph, prm, elast)
        tot = {}
        for t in gathered:
            for kk, v in t.items():
                tot[kk] = tot.get(kk, 0.0) + v
        scale = max(abs(v) for v in ref.values())
        missing = [kk for kk in ref if kk not in tot and ref[kk] != 0.0]
        extra = [kk for kk in tot if kk not in ref and tot[kk] != 0.0]
        assert not missing and not extra, (len(missing), len(extra))
        err = max(abs(tot[kk] - v) for kk, v in ref.items() if kk in tot) / scale
        worstq = max(worstq, err)
        print("Q1 %s p=%d ne=%s: %d entries summed over %d ranks, max rel err vs single-domain oracle %.2e" % (kind, p, ne, len(tot), world, err), flush=True)
    op.close()
assert worstq <= 1e-12, worstq
dist.barrier()
if rank == 0:
    print("MGPU PARITY OK (world %d): MatMult worst %.2e, Q1 assembly worst %.2e" % (world, worst, worstq))
dist.destroy_process_group()
