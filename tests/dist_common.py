"""Shared by the multi-rank tests: block-partitioned ellip MatMult against the single-domain oracle, matched by node
coordinates (the global numbering of a partitioned space differs from the single-domain one)."""
import numpy as np


def field(xyz):
    return np.sin(1.3 * xyz[:, 0] + 0.4) * np.cos(0.7 * xyz[:, 1]) + 0.5 * xyz[:, 2] ** 2 + 0.1 * xyz[:, 0] * xyz[:, 1]


def key(xyz):
    return [tuple(r) for r in np.round(xyz * 1e7).astype(np.int64)]


def single_domain_reference(ne, p, morph, prm):
    import oracle_lib as ol
    from dohp_b200 import lib
    mesh = lib.BoxMesh(ne, p, **morph)
    o = ol.OracleEllip(mesh, **prm)
    x = field(mesh.node_xyz[:mesh.n])
    o.function(0.3 * x, dcache=np.zeros(mesh.nc - mesh.n))          # writes the stash at state 0.3*x (homogeneous Dirichlet data)
    y = o.mult(x)
    return dict(zip(key(mesh.node_xyz[:mesh.n]), y)), mesh.n


def single_domain_q1(ne, p, morph, prm, elast=False):
    """The single-domain oracle's embedded-Q1 matrix as {(row key, col key[, fi, fj]): value}, linearised at 0.3*field."""
    import oracle_lib as ol
    from dohp_b200 import lib
    from dohp_b200.box import q1_pattern
    mesh = lib.BoxMesh(ne, p, **morph)
    rp, ci = q1_pattern(mesh, bs=1)
    k = key(mesh.node_xyz[:mesh.n])
    rows = np.repeat(np.arange(mesh.n), np.diff(rp))
    if not elast:
        o = ol.OracleEllip(mesh, **prm)
        x = 0.3 * field(mesh.node_xyz[:mesh.n])
        vals = o.jacobian_q1(x, rp, ci, dcache=np.zeros(mesh.nc - mesh.n))
        return {(k[r], k[c]): v for r, c, v in zip(rows, ci, vals)}
    import scipy.sparse as sp
    o = ol.OracleElast(mesh, **prm)
    x = 0.05 * np.repeat(field(mesh.node_xyz[:mesh.n]), 3) * np.tile([1.0, -0.7, 0.4], mesh.n)
    rp3, ci3 = q1_pattern(mesh, bs=3)
    J = sp.csr_matrix((o.jacobian_q1(x, rp3, ci3, dcache=np.zeros((mesh.nc - mesh.n) * 3)), ci3, rp3), shape=(mesh.n * 3, mesh.n * 3)).tocoo()
    return {(k[r // 3], k[c // 3], r % 3, c % 3): v for r, c, v in zip(J.row, J.col, J.data)}
