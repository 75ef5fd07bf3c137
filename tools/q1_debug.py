import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
import scipy.sparse as sp
import oracle_lib as ol
from dohp_b200 import lib
ol.build_oracle(ref=False)
np.set_printoptions(linewidth=250, precision=4, suppress=True)
ctx = lib.Context(0, stream=torch.cuda.current_stream().cuda_stream)
for ne, deg in [((1, 1, 1), 1), ((2, 1, 1), 1)]:
    mesh = lib.BoxMesh(ne, deg, dirichlet=0)
    o = ol.OracleEllip(mesh, exact=0, a=1.1, b=0.9, c=1.3, p=2.0, lam=0.0, eps=0.9)
    rule = lib.Rule.preferred(ctx, mesh.p); efs = lib.EFS(ctx, rule, mesh.p); space = lib.Space.from_mesh(ctx, mesh)
    op = lib.Op(ctx, lib.OP_ELLIP, lib.EllipParam(2.0, 0.9, 0.0), rule, [space], [efs], mesh.coords)
    op.set_forcing(o.forcing(op.quadrature_points())); op.set_dirichlet(0, o.dirichlet_cache())
    rowptr, colidx = op.q1_pattern(mesh.l2g)
    x = 0.3 * np.random.default_rng(deg).standard_normal(mesh.n)
    vals = op.assemble_q1(x, mesh.l2g, rowptr, colidx)
    ref = o.jacobian_q1(x, rowptr, colidx)
    G = sp.csr_matrix((vals, colidx, rowptr), shape=(mesh.n, mesh.n)).toarray()
    R = sp.csr_matrix((ref, colidx, rowptr), shape=(mesh.n, mesh.n)).toarray()
    print("ne", ne, "e2l", mesh.e2l.reshape(mesh.nelem, -1))
    print((np.abs(G - R) > 1e-10).astype(int))
    print(G * 18); print(R * 18)
