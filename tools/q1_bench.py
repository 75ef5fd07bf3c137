"""Times the embedded-Q1 assembly (dB200OpAssembleQ1 through the C ABI, cached pattern, device-resident state and values)
for a few degrees:   python tools/q1_bench.py [ellip|elast] [p:ne ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from dohp_b200 import lib

kind = sys.argv[1] if len(sys.argv) > 1 else "ellip"
cases = [tuple(int(v) for v in a.split(":")) for a in sys.argv[2:]] or ([(2, 64), (4, 64), (6, 40), (8, 32)] if kind == "ellip" else [(2, 48), (4, 48), (6, 48)])
stream = torch.cuda.current_stream()
ctx = lib.Context(0, stream=stream.cuda_stream)
for p, ne in cases:
    bs = 3 if kind == "elast" else 1
    mesh = lib.BoxMesh((ne, ne, ne), p, node_xyz=False, ghost_l2g=False, morph=float(os.environ.get("MORPH", 0)), twist=0.4 * float(os.environ.get("MORPH", 0)))
    rule = lib.Rule.preferred(ctx, p); efs = lib.EFS(ctx, rule, p); space = lib.Space.from_mesh(ctx, mesh, bs=bs)
    if kind == "elast":
        op = lib.Op(ctx, lib.OP_ELAST, lib.ElastParam(0.5769, 0.3846), rule, [space], [efs], mesh.coords)
    else:
        op = lib.Op(ctx, lib.OP_ELLIP, lib.EllipParam(float(os.environ.get("ELLIP_P", 1.6)), 1.0, float(os.environ.get("ELLIP_LAMBDA", 0.0))), rule, [space], [efs], mesh.coords)
    x = 0.01 * torch.sin(0.37 * torch.arange(mesh.n * bs, dtype=torch.float64, device="cuda"))
    rowptr, colidx = op.q1_pattern(mesh.l2g)
    nnzb = len(colidx)
    vals = torch.empty(nnzb * bs * bs, dtype=torch.float64, device="cuda")
    l2g = torch.from_numpy(np.ascontiguousarray(mesh.l2g)).cuda()
    call = lambda: lib.chk(lib.load().dB200OpAssembleQ1(op.h, (lib.C.c_void_p * 1)(x.data_ptr()), lib.ptr(l2g), None, None, lib.ptr(vals)))
    call(); call()
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(reps):
        call()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    by = nnzb * (8 * bs * bs + 4)
    print("%s p=%d ne=%d^3: %.3f ms  nnz blocks %d  %.1f GB/s algorithmic (12 B per AIJ entry / 76 B per BAIJ-3 block)  checksum %.12e"
          % (kind, p, ne, ms, nnzb, by / ms / 1e6, float(vals.sum())), flush=True)
    op.close()
    del vals, x, l2g, mesh, op, space
    torch.cuda.empty_cache()
