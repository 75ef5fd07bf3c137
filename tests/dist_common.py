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
