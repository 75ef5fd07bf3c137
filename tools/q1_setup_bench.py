"""Times the one-off Q1 pattern build (dB200OpQ1Pattern) and the first assembly (row records):  python tools/q1_setup_bench.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from dohp_b200 import lib
ctx = lib.Context(0, stream=torch.cuda.current_stream().cuda_stream)
for kind, p, ne in (("ellip", 4, 64), ("elast", 6, 48)):
    bs = 3 if kind == "elast" else 1
    mesh = lib.BoxMesh((ne, ne, ne), p, node_xyz=False, ghost_l2g=False)
    rule = lib.Rule.preferred(ctx, p); efs = lib.EFS(ctx, rule, p); space = lib.Space.from_mesh(ctx, mesh, bs=bs)
    op = lib.Op(ctx, lib.OP_ELAST if bs == 3 else lib.OP_ELLIP, lib.ElastParam(0.5769, 0.3846) if bs == 3 else lib.EllipParam(1.6, 1.0, 0.0), rule, [space], [efs], mesh.coords)
    l2g = torch.from_numpy(np.ascontiguousarray(mesh.l2g)).cuda()
    nrows, nnz = lib.C.c_longlong(), lib.C.c_longlong()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    lib.chk(lib.load().dB200OpQ1Pattern(op.h, lib.ptr(l2g), lib.C.byref(nrows), lib.C.byref(nnz), None, None))
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print("%s p=%d %d^3: pattern build %.1f ms (%d rows, %d blocks)" % (kind, p, ne, (t1 - t0) * 1e3, nrows.value, nnz.value), flush=True)
    op.close()
