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
from dist_common import field, key, single_domain_reference
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
dist.barrier()
if rank == 0:
    print("MGPU PARITY OK (world %d): worst %.2e" % (world, worst))
dist.destroy_process_group()
