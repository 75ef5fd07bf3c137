"""One launch of each non-headline kernel at its BASELINE.json size, for an `ncu --set full` capture:
   ncu --set full --clock-control none --import-source on -k regex:'q1_assemble|vec_kernel|ellip_kernel' -c 8 -o gpurun_out/prof python tools/prof_kernels.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from dohp_b200 import lib

which = sys.argv[1:] or ["q1", "plap", "elast", "stokes"]
stream = torch.cuda.current_stream()
ctx = lib.Context(0, stream=stream.cuda_stream)
for w in which:
    if w.startswith("lin"):                    # headline kernel at degree p: "lin4", "lin8", ...
        p = int(w[3:])
        mesh = lib.BoxMesh((64, 64, 64), p, node_xyz=False, ghost_l2g=False, morph=float(os.environ.get("MORPH", 0)), twist=0.4 * float(os.environ.get("MORPH", 0)))
        rule = lib.Rule.preferred(ctx, p); efs = lib.EFS(ctx, rule, p); space = lib.Space.from_mesh(ctx, mesh)
        op = lib.Op(ctx, lib.OP_ELLIP, lib.EllipParam(2.0, 1.0, 0.0), rule, [space], [efs], mesh.coords)
        x = torch.sin(0.37 * torch.arange(mesh.n, dtype=torch.float64, device="cuda")); y = torch.empty_like(x)
        op.mult(x, y)
        torch.cuda.synchronize()
        op.close(); del mesh, x, y, space
if "q1" in which or "plap" in which:
    p = 4
    mesh = lib.BoxMesh((64, 64, 64), p, node_xyz=False, ghost_l2g=False)
    rule = lib.Rule.preferred(ctx, p); efs = lib.EFS(ctx, rule, p); space = lib.Space.from_mesh(ctx, mesh)
    op = lib.Op(ctx, lib.OP_ELLIP, lib.EllipParam(1.6, 1.0, 0.0), rule, [space], [efs], mesh.coords)
    x = torch.sin(0.37 * torch.arange(mesh.n, dtype=torch.float64, device="cuda")); y = torch.empty_like(x)
    op.function(x, y)
    op.mult(x, y)
    if "q1" in which:
        rowptr, colidx = op.q1_pattern(mesh.l2g)
        vals = torch.empty(len(colidx), dtype=torch.float64, device="cuda")
        rp, ci = torch.from_numpy(rowptr).cuda(), torch.from_numpy(colidx).cuda()
        l2g = torch.from_numpy(np.ascontiguousarray(mesh.l2g)).cuda()
        op.assemble_q1(x, l2g, None, None, vals)
    torch.cuda.synchronize()
if "elast" in which:
    p = 6
    mesh = lib.BoxMesh((32, 32, 32), p, node_xyz=False, ghost_l2g=False)
    rule = lib.Rule.preferred(ctx, p); efs = lib.EFS(ctx, rule, p); space = lib.Space.from_mesh(ctx, mesh, bs=3)
    op = lib.Op(ctx, lib.OP_ELAST, lib.ElastParam(0.5769, 0.3846), rule, [space], [efs], mesh.coords)
    x = 0.01 * torch.sin(0.37 * torch.arange(mesh.n * 3, dtype=torch.float64, device="cuda")); y = torch.empty_like(x)
    op.function(x, y)
    op.mult(x, y)
    torch.cuda.synchronize()
if "q1elast" in which:
    p = 6
    mesh = lib.BoxMesh((32, 32, 32), p, node_xyz=False, ghost_l2g=False)
    rule = lib.Rule.preferred(ctx, p); efs = lib.EFS(ctx, rule, p); space = lib.Space.from_mesh(ctx, mesh, bs=3)
    op = lib.Op(ctx, lib.OP_ELAST, lib.ElastParam(0.5769, 0.3846), rule, [space], [efs], mesh.coords)
    x = 0.01 * torch.sin(0.37 * torch.arange(mesh.n * 3, dtype=torch.float64, device="cuda"))
    rowptr, colidx = op.q1_pattern(mesh.l2g)
    vals = torch.empty(len(colidx) * 9, dtype=torch.float64, device="cuda")
    l2g = torch.from_numpy(np.ascontiguousarray(mesh.l2g)).cuda()
    op.assemble_q1(x, l2g, None, None, vals)
    torch.cuda.synchronize()
    op.close(); del vals, mesh, x, space
    torch.cuda.empty_cache()
if "stokes" in which:
    mu, mp = lib.BoxMesh((32, 32, 32), 4, node_xyz=False, ghost_l2g=False), lib.BoxMesh((32, 32, 32), 3, dirichlet=0, node_xyz=False, ghost_l2g=False)
    rule = lib.Rule.preferred(ctx, 4); efsu, efsp = lib.EFS(ctx, rule, 4), lib.EFS(ctx, rule, 3)
    su, sp = lib.Space.from_mesh(ctx, mu, bs=3), lib.Space.from_mesh(ctx, mp, bs=1)
    op = lib.Op(ctx, lib.OP_STOKES, lib.StokesParam(1.0, 0.1, 1.2, 0.0), rule, [su, sp], [efsu, efsp], mu.coords)
    xu = 0.01 * torch.sin(0.37 * torch.arange(mu.n * 3, dtype=torch.float64, device="cuda")); xp = torch.cos(0.11 * torch.arange(mp.n, dtype=torch.float64, device="cuda"))
    yu, yp = torch.empty_like(xu), torch.empty_like(xp)
    op.function([xu, xp], [yu, yp])
    op.mult([xu, xp], [yu, yp])
    torch.cuda.synchronize()
print("done")
