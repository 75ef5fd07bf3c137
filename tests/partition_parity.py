"""Partitioned-space parity: a small fixed GLOBAL problem is block-partitioned over the ranks of the job (one process per GPU,
genblock's partition sets, genblock.c:286-303), every rank runs the library's fused traversals on its part with the NCCL ghost
exchange inside (VecGhostUpdate pair, fs.c:578-579, 627-628 -- for Stokes one pair per space, MatMult_Nest_StokesCoupled,
stokes.c:541-577), and the assembled result is compared with the SINGLE-DOMAIN oracle, matched by node coordinates (the global
numbering of a partitioned space differs from the single-domain one).  Also the embedded-Q1 rows: every rank assembles owned
rows + one row per ghost node (what MatSetValuesBlockedLocal stashes for the owner, fs.c:919); the ranks' entries are summed by
global (row, column) -- MatAssemblyEnd's job -- and compared with the single-domain oracle matrix.

Test infrastructure (imports the oracle): used by tests/mgpu_check.py, tests/test_gpu_partition.py and by bench.py's untimed
"parity" block, which is how the driver sees multi-GPU correctness at N = 2, 4, 8.  world == 1 runs the same code on a 1x1x1 grid."""
import numpy as np

from dist_common import field, key


def _gather(dist, world, obj):
    if world == 1:
        return [obj]
    out = [None] * world
    dist.all_gather_object(out, obj)
    return out


def _relerr_by_key(parts, ref):
    """parts: list of (keys, values) per rank; ref: dict key -> value.  Every key must appear exactly once over the ranks."""
    seen, worst = 0, 0.0
    scale = max(abs(v) for v in ref.values())
    for keys, vals in parts:
        for k, v in zip(keys, vals):
            worst = max(worst, abs(v - ref[k]))
            seen += 1
    assert seen == len(ref), ("owned dofs over the ranks %d != single-domain dofs %d" % (seen, len(ref)))
    return worst / scale


def ellip_matmult(ctx, lib, ol, dist, rank, world, procs, ne=(4, 4, 4), p=3):
    morph, prm = dict(morph=0.2, twist=0.15), dict(p=1.7, lam=0.3, eps=0.8)
    mesh = lib.BoxMesh(ne, p, procs=procs, rank=rank, **morph)
    rule = lib.Rule.preferred(ctx, p)
    efs = lib.EFS(ctx, rule, p)
    space = lib.Space.from_mesh(ctx, mesh)
    op = lib.Op(ctx, lib.OP_ELLIP, lib.EllipParam(prm["p"], prm["eps"], prm["lam"]), rule, [space], [efs], mesh.coords)
    x = field(mesh.node_xyz[:mesh.n])
    op.function(0.3 * x)                       # stash at 0.3 x (homogeneous-zero Dirichlet data, no forcing)
    y = op.mult(x)
    parts = _gather(dist, world, (key(mesh.node_xyz[:mesh.n]), y))
    err = None
    if rank == 0:
        m1 = lib.BoxMesh(ne, p, **morph)
        o = ol.OracleEllip(m1, **prm)
        x1 = field(m1.node_xyz[:m1.n])
        o.function(0.3 * x1, dcache=np.zeros(m1.nc - m1.n))
        err = _relerr_by_key(parts, dict(zip(key(m1.node_xyz[:m1.n]), o.mult(x1))))
    op.close()
    return err


def _vec_keys(xyz, bs):
    k = key(xyz)
    return [(kk, f) for kk in k for f in range(bs)]


def _elast_state(xyz):
    return 0.05 * np.repeat(field(xyz), 3) * np.tile([1.0, -0.7, 0.4], len(xyz))


def elast_matmult(ctx, lib, ol, dist, rank, world, procs, ne=(4, 4, 4), p=2):
    morph = dict(morph=0.2, twist=0.15)
    mesh = lib.BoxMesh(ne, p, procs=procs, rank=rank, **morph)
    rule = lib.Rule.preferred(ctx, p)
    efs = lib.EFS(ctx, rule, p)
    space = lib.Space.from_mesh(ctx, mesh, bs=3)
    E0, nu = 1.0, 0.3
    lam, mu = E0 * nu / ((1 + nu) * (1 - 2 * nu)), E0 / (2 * (1 + nu))
    op = lib.Op(ctx, lib.OP_ELAST, lib.ElastParam(lam, mu), rule, [space], [efs], mesh.coords)
    x = _elast_state(mesh.node_xyz[:mesh.n])
    op.function(x)
    dx = np.repeat(field(mesh.node_xyz[:mesh.n]), 3) * np.tile([0.3, 1.0, -0.5], mesh.n)
    y = op.mult(dx)
    parts = _gather(dist, world, (_vec_keys(mesh.node_xyz[:mesh.n], 3), y))
    err = None
    if rank == 0:
        m1 = lib.BoxMesh(ne, p, **morph)
        o = ol.OracleElast(m1, E0=E0, nu=nu)
        o.function(_elast_state(m1.node_xyz[:m1.n]), dcache=np.zeros((m1.nc - m1.n) * 3))
        dx1 = np.repeat(field(m1.node_xyz[:m1.n]), 3) * np.tile([0.3, 1.0, -0.5], m1.n)
        err = _relerr_by_key(parts, dict(zip(_vec_keys(m1.node_xyz[:m1.n], 3), o.mult(dx1))))
    op.close()
    return err


def stokes_matmult(ctx, lib, ol, dist, rank, world, procs, ne=(4, 4, 4), p=3):
    """Coupled Stokes Jacobian on TWO partitioned spaces (velocity degree p, bs 3, Dirichlet walls; pressure degree p-1, bs 1,
    no Dirichlet nodes): two forward and two reverse ghost exchanges per product.  Also the A, B, B^T blocks."""
    morph, prm = dict(morph=0.15, twist=0.1), dict(B=1.0, eps=0.1, p=1.2)
    mu_ = lib.BoxMesh(ne, p, procs=procs, rank=rank, **morph)
    mp_ = lib.BoxMesh(ne, p - 1, procs=procs, rank=rank, dirichlet=0, **morph)
    rule = lib.Rule.preferred(ctx, p)
    efsu, efsp = lib.EFS(ctx, rule, p), lib.EFS(ctx, rule, p - 1)
    su, sp = lib.Space.from_mesh(ctx, mu_, bs=3), lib.Space.from_mesh(ctx, mp_, bs=1)
    op = lib.Op(ctx, lib.OP_STOKES, lib.StokesParam(prm["B"], prm["eps"], prm["p"], 0.0), rule, [su, sp], [efsu, efsp], mu_.coords)

    def state(mu, mp):
        xu = 0.2 * np.repeat(field(mu.node_xyz[:mu.n]), 3) * np.tile([1.0, -0.6, 0.8], mu.n)
        xp = np.cos(1.1 * mp.node_xyz[:mp.n, 0]) + 0.3 * mp.node_xyz[:mp.n, 1] * mp.node_xyz[:mp.n, 2]
        du = np.repeat(np.sin(0.9 * mu.node_xyz[:mu.n, 1] + 0.2), 3) * np.tile([0.5, 1.0, -0.25], mu.n) + 0.1 * xu
        dp = field(mp.node_xyz[:mp.n])
        return xu, xp, du, dp

    xu, xp, du, dp = state(mu_, mp_)
    op.function([xu, xp])
    ku, kp = _vec_keys(mu_.node_xyz[:mu_.n], 3), key(mp_.node_xyz[:mp_.n])
    res = {}
    for name, block in (("J", None), ("A", lib.STOKES_A), ("B", lib.STOKES_B), ("Bt", lib.STOKES_Bt)):
        yu, yp = op.mult([du, dp], block=block)
        res[name] = (np.array(yu), np.array(yp))
    parts = _gather(dist, world, (ku, kp, res))
    err = None
    if rank == 0:
        m1u, m1p = lib.BoxMesh(ne, p, **morph), lib.BoxMesh(ne, p - 1, dirichlet=0, **morph)
        o = ol.OracleStokes(m1u, m1p, exact=1, **prm)
        x1u, x1p, d1u, d1p = state(m1u, m1p)
        o.function(x1u, x1p, dcache=np.zeros((m1u.nc - m1u.n) * 3))
        k1u, k1p = _vec_keys(m1u.node_xyz[:m1u.n], 3), key(m1p.node_xyz[:m1p.n])
        err = 0.0
        for name, block in (("J", 0), ("A", lib.STOKES_A), ("B", lib.STOKES_B), ("Bt", lib.STOKES_Bt)):
            ru, rp = o.mult(d1u, d1p, block=block)
            if name in ("J", "A", "Bt"):
                err = max(err, _relerr_by_key([(q[0], q[2][name][0]) for q in parts], dict(zip(k1u, ru))))
            if name in ("J", "B"):
                err = max(err, _relerr_by_key([(q[1], q[2][name][1]) for q in parts], dict(zip(k1p, rp))))
    op.close()
    return err


def q1_rows(ctx, lib, ol, dist, rank, world, procs, kind="ellip", ne=(4, 4, 4), p=2):
    from dist_common import single_domain_q1
    morph = dict(morph=0.2, twist=0.15)
    elast = kind == "elast"
    prm = dict(E0=1.0, nu=0.3) if elast else dict(p=1.7, lam=0.3, eps=0.8)
    mesh = lib.BoxMesh(ne, p, procs=procs, rank=rank, **morph)
    rule = lib.Rule.preferred(ctx, p)
    efs = lib.EFS(ctx, rule, p)
    bs = 3 if elast else 1
    space = lib.Space.from_mesh(ctx, mesh, bs=bs)
    if elast:
        lam, mu = prm["E0"] * prm["nu"] / ((1 + prm["nu"]) * (1 - 2 * prm["nu"])), prm["E0"] / (2 * (1 + prm["nu"]))
        op = lib.Op(ctx, lib.OP_ELAST, lib.ElastParam(lam, mu), rule, [space], [efs], mesh.coords)
        x = _elast_state(mesh.node_xyz[:mesh.n])
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
    # owner-side reduction over NCCL (MatAssemblyBegin/End, ellip.c:528-529): every rank ends with its owned rows of the global matrix
    orowptr, ocolidx = op.q1_owner_pattern(rowptr, colidx)
    ovals = op.q1_owner_reduce(vals.reshape(-1)).reshape(len(ocolidx), bs, bs)
    assert len(orowptr) == mesh.n + 1 and all(np.all(np.diff(ocolidx[orowptr[i]:orowptr[i + 1]]) > 0) for i in range(mesh.n))     # sorted, unique
    gkeys = {}
    for d in _gather(dist, world, {int(mesh.l2g[i]): kloc[i] for i in range(mesh.n)}):
        gkeys.update(d)
    orows = np.repeat(np.arange(mesh.n), np.diff(orowptr))
    otrip = {}
    for r, c, blk in zip(orows, ocolidx, ovals):
        for fi in range(bs):
            for fj in range(bs):
                kk = (kloc[r], gkeys[int(c)], fi, fj) if elast else (kloc[r], gkeys[int(c)])
                assert kk not in otrip
                otrip[kk] = blk[fi, fj]
    oparts = _gather(dist, world, otrip)
    parts = _gather(dist, world, trip)
    err = None
    if rank == 0:
        ref = single_domain_q1(ne, p, morph, prm, elast)
        scale = max(abs(v) for v in ref.values())
        # owned rows after the NCCL reduction: every entry of the global matrix exactly once over the ranks
        seen, oerr = set(), 0.0
        for t in oparts:
            for kk, v in t.items():
                assert kk not in seen, "an owned-row entry appears on two ranks"
                seen.add(kk)
                assert kk in ref or v == 0.0
                oerr = max(oerr, abs(v - ref.get(kk, 0.0)))
        assert not [kk for kk in ref if kk not in seen and ref[kk] != 0.0]
        tot = {}
        for t in parts:
            for kk, v in t.items():
                tot[kk] = tot.get(kk, 0.0) + v
        missing = [kk for kk in ref if kk not in tot and ref[kk] != 0.0]
        extra = [kk for kk in tot if kk not in ref and tot[kk] != 0.0]
        assert not missing and not extra, (len(missing), len(extra))
        err = max(max(abs(tot[kk] - v) for kk, v in ref.items() if kk in tot), oerr) / scale
    op.close()
    return err


def run_all(ctx, dist, rank, world, procs, quick=False):
    """Returns (on rank 0) {"ellip": .., "elast": .., "stokes": .., "q1": ..} = max relative errors vs the single-domain oracle."""
    import oracle_lib as ol
    from dohp_b200 import lib
    ol.build_oracle(ref=False)
    out = {}
    ne = tuple(2 * m for m in procs) if quick else tuple(max(4, 2 * m) for m in procs)
    cases_e = [(3, ne)] if quick else [(2, ne), (3, (4, 2 * procs[1], 4)), (5, tuple(2 * m for m in procs))]
    errs = [ellip_matmult(ctx, lib, ol, dist, rank, world, procs, ne=n_, p=p_) for p_, n_ in cases_e]
    out["ellip"] = max(errs) if rank == 0 else None
    out["elast"] = elast_matmult(ctx, lib, ol, dist, rank, world, procs, ne=ne, p=2)
    out["stokes"] = stokes_matmult(ctx, lib, ol, dist, rank, world, procs, ne=ne, p=3)
    errs = [q1_rows(ctx, lib, ol, dist, rank, world, procs, kind="ellip", ne=ne, p=2)]
    if not quick:
        errs.append(q1_rows(ctx, lib, ol, dist, rank, world, procs, kind="ellip", ne=(4, 2 * procs[1], 4), p=3))
    errs.append(q1_rows(ctx, lib, ol, dist, rank, world, procs, kind="elast", ne=tuple(2 * m for m in procs), p=2))
    out["q1"] = max(errs) if rank == 0 else None
    return out if rank == 0 else None
