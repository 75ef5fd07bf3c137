"""Times the fused operator kernels of one build of the library (DOHP_B200_LIB selects a variant): ellip linear Poisson sweep and
p-Laplace Jacobian on 64^3, elast 48^3 p=6, stokes 32^3 p4/p3.  One line of JSON per run; a development tool, not the bench."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dohp_b200 import lib

what = sys.argv[1] if len(sys.argv) > 1 else "lin,jac,elast,stokes"
degs = [int(d) for d in (sys.argv[2] if len(sys.argv) > 2 else "2,3,4,5,6,7,8,9,10").split(",")]
morph = float(os.environ.get("MORPH", 0))
stream = torch.cuda.current_stream()
ctx = lib.Context(0, stream=stream.cuda_stream)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = {"lib": os.path.basename(lib.LIBPATH), "morph": morph}
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float64, device="cuda")       # 512 MB > L2
for mode in what.split(","):
    if mode in ("lin", "jac"):
        res = {}
        for p in degs:
            mesh = lib.BoxMesh((64, 64, 64), p, node_xyz=False, ghost_l2g=False, morph=morph, twist=0.4 * morph)
            rule = lib.Rule.preferred(ctx, p); efs = lib.EFS(ctx, rule, p); space = lib.Space.from_mesh(ctx, mesh)
            op = lib.Op(ctx, lib.OP_ELLIP, lib.EllipParam(2.0 if mode == "lin" else 1.6, 1.0, 0.0), rule, [space], [efs], mesh.coords)
            x = torch.sin(0.37 * torch.arange(mesh.n, dtype=torch.float64, device="cuda")); y = torch.empty_like(x)
            if mode == "jac":
                op.function(x, y)
            def f():
                if p <= 4:
                    flush.zero_()
                op.mult(x, y)
            ms = timed(f)
            if p <= 4:
                ms -= timed(lambda: flush.zero_())
            res[p] = round(ms, 4)
            op.close(); del mesh, x, y, space
        out[mode + "_ms"] = res
        n64 = {p: (64 * p - 1) ** 3 for p in degs}
        out[mode + "_gdofs"] = {p: round(n64[p] / res[p] / 1e6, 2) for p in degs}
        out[mode + "_sweep_gdofs"] = round(sum(n64.values()) / sum(res.values()) / 1e6, 2)
    if mode == "elast":
        p = 6
        mesh = lib.BoxMesh((48, 48, 48), p, node_xyz=False, ghost_l2g=False, morph=morph, twist=0.4 * morph)
        rule = lib.Rule.preferred(ctx, p); efs = lib.EFS(ctx, rule, p); space = lib.Space.from_mesh(ctx, mesh, bs=3)
        op = lib.Op(ctx, lib.OP_ELAST, lib.ElastParam(0.5769, 0.3846), rule, [space], [efs], mesh.coords)
        x = 0.01 * torch.sin(0.37 * torch.arange(mesh.n * 3, dtype=torch.float64, device="cuda")); y = torch.empty_like(x)
        op.function(x, y)
        ms = timed(lambda: op.mult(x, y), 5)
        out["elast_p6"] = {"ms": round(ms, 4), "gdofs": round(mesh.n * 3 / ms / 1e6, 2), "function_ms": round(timed(lambda: op.function(x, y), 3), 4)}
        op.close(); del mesh, x, y, space
    if mode == "stokes":
        mu, mp = lib.BoxMesh((32, 32, 32), 4, node_xyz=False, ghost_l2g=False, morph=morph, twist=0.4 * morph), lib.BoxMesh((32, 32, 32), 3, dirichlet=0, node_xyz=False, ghost_l2g=False, morph=morph, twist=0.4 * morph)
        rule = lib.Rule.preferred(ctx, 4); efsu, efsp = lib.EFS(ctx, rule, 4), lib.EFS(ctx, rule, 3)
        su, sp = lib.Space.from_mesh(ctx, mu, bs=3), lib.Space.from_mesh(ctx, mp, bs=1)
        op = lib.Op(ctx, lib.OP_STOKES, lib.StokesParam(1.0, 0.1, 1.2, 0.0), rule, [su, sp], [efsu, efsp], mu.coords)
        xu = 0.01 * torch.sin(0.37 * torch.arange(mu.n * 3, dtype=torch.float64, device="cuda")); xp = torch.cos(0.11 * torch.arange(mp.n, dtype=torch.float64, device="cuda"))
        yu, yp = torch.empty_like(xu), torch.empty_like(xp)
        op.function([xu, xp], [yu, yp])
        def f():
            flush.zero_()
            op.mult([xu, xp], [yu, yp])
        ms = timed(f) - timed(lambda: flush.zero_())
        out["stokes_p4p3"] = {"ms": round(ms, 4), "gdofs": round((mu.n * 3 + mp.n) / ms / 1e6, 2)}
        op.close()
print(json.dumps(out))
