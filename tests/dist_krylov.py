"""Distributed Newton-Krylov driver for BASELINE.json configs[3] (stokes, velocity p / pressure p-1, SNES + KSP on 1/2/4/8 GPUs):
test infrastructure that plays the role PETSc plays for the reference (src/fs/tests/stokes.c:1080-1160 with the options of
src/fs/tests/CMakeLists.txt:65-70: Newton with the cubic backtracking line search, FGMRES on the coupled matrix-free Jacobian
(MatMult_Nest_StokesCoupled, stokes.c:541-577), right-preconditioned by the upper block-triangular Schur factorisation
[A B^T; 0 S] with an inner CG on the velocity block A (rtol 1e-2) and S ~ -D, D = the pressure mass matrix scaled by 1/eta
(StokesJacobianAssemble_Pressure, stokes.c:760-794)).  The inner preconditioners are the partition-independent diagonals: the 3x3
diagonal blocks of the embedded-Q1 velocity matrix Ap (stokes.c:675-737) and diag(D) -- assembled on every rank by the library,
off-process (ghost-row) contributions added to the owners through the space's reverse ghost exchange -- so the iteration counts do
not depend on the number of GPUs, which is what the tests check.

Vectors are any array type with + - * (numpy arrays on the CPU oracle, torch CUDA tensors on the GPUs); inner products go through
an `ops` object (all-reduce over the ranks)."""
import numpy as np


class NumpyOps:
    def dot(self, a, b):
        return float(np.dot(a, b))

    def norm(self, a):
        return float(np.linalg.norm(a))

    def zeros_like(self, a):
        return np.zeros_like(a)


class TorchDistOps:
    """Owned parts of distributed vectors as torch tensors; reductions over the process group (NCCL on GPUs, gloo on CPU)."""

    def __init__(self, dist, world):
        self.dist, self.world = dist, world

    def dot(self, a, b):
        import torch
        s = torch.dot(a, b).reshape(1)
        if self.world > 1:
            self.dist.all_reduce(s)
        return float(s.item())

    def norm(self, a):
        return self.dot(a, a) ** 0.5

    def zeros_like(self, a):
        import torch
        return torch.zeros_like(a)


def cg(A, M, b, ops, rtol=1e-2, maxit=500):
    x = ops.zeros_like(b)
    r = b
    z = M(r)
    dp0 = dp = ops.norm(z)
    p, beta_old, its = None, None, 0
    while dp > rtol * dp0 and its < maxit:
        beta = ops.dot(r, z)
        p = z if p is None else z + (beta / beta_old) * p
        w = A(p)
        a = beta / ops.dot(p, w)
        x = x + a * p
        r = r - a * w
        z = M(r)
        dp = ops.norm(z)
        beta_old = beta
        its += 1
    return x, its


def fgmres(A, M, b, ops, rtol=1e-2, restart=30, maxit=300):
    """Flexible GMRES (right preconditioning, KSPFGMRES): monitored norm = unpreconditioned residual (recurrence estimate)."""
    x = ops.zeros_like(b)
    r = b
    beta0 = ops.norm(r)
    hist = [beta0]
    its = 0
    while True:
        beta = ops.norm(r)
        if beta <= rtol * beta0 or its >= maxit:
            return x, hist
        V, Z = [r * (1.0 / beta)], []
        H = np.zeros((restart + 1, restart))
        g = np.zeros(restart + 1)
        g[0] = beta
        cs, sn = np.zeros(restart), np.zeros(restart)
        k_used, done = 0, False
        for k in range(restart):
            Z.append(M(V[k]))
            w = A(Z[k])
            hk = [ops.dot(w, v) for v in V]
            for i, v in enumerate(V):
                w = w - hk[i] * v
            H[:k + 1, k] = hk
            H[k + 1, k] = ops.norm(w)
            V.append(w * (1.0 / H[k + 1, k]) if H[k + 1, k] > 0 else w)
            for i in range(k):
                t = cs[i] * H[i, k] + sn[i] * H[i + 1, k]
                H[i + 1, k] = -sn[i] * H[i, k] + cs[i] * H[i + 1, k]
                H[i, k] = t
            d = np.hypot(H[k, k], H[k + 1, k])
            cs[k], sn[k] = H[k, k] / d, H[k + 1, k] / d
            H[k, k] = d
            H[k + 1, k] = 0
            g[k + 1] = -sn[k] * g[k]
            g[k] = cs[k] * g[k]
            its += 1
            k_used = k + 1
            hist.append(abs(g[k + 1]))
            if abs(g[k + 1]) <= rtol * beta0 or its >= maxit:
                done = True
                break
        y = np.linalg.solve(np.triu(H[:k_used, :k_used]), g[:k_used])
        for i in range(k_used):
            x = x + float(y[i]) * Z[i]
        if done:
            return x, hist
        r = b - A(x)


def newton_bt(function, mult, solve, x0, ops, rtol=1e-7, maxit=30, alpha=1e-4):
    """SNESNEWTONLS with PETSc's cubic backtracking line search (see tests/petsc_like.py linesearch_bt).  `solve(x, F)` returns the
    Newton direction y ~ J(x)^-1 F and its iteration record."""
    x = x0
    F = function(x)
    fnorms, records = [ops.norm(F)], []
    while fnorms[-1] > rtol * fnorms[0] and len(fnorms) <= maxit:
        y, rec = solve(x, F)
        records.append(rec)
        fnorm = fnorms[-1]
        f = fnorm * fnorm
        initslope = ops.dot(F, mult(y))
        initslope = -abs(initslope) if initslope != 0 else -1.0
        lam = 1.0
        trial = lambda l: function(x - l * y)
        G = trial(lam)
        gnorm = ops.norm(G)
        g = gnorm * gnorm
        if not .5 * g <= .5 * f + lam * alpha * initslope:
            lamtemp = -initslope / (g - f - 2.0 * lam * initslope)
            lamprev, gprev = lam, g
            lam = .1 * lam if lamtemp <= .1 * lam else min(lamtemp, .5 * lam)
            G = trial(lam)
            gnorm = ops.norm(G)
            g = gnorm * gnorm
            count = 0
            while not .5 * g < .5 * f + lam * alpha * initslope:
                count += 1
                if lam <= 1e-12 or count > 40:
                    raise RuntimeError("line search failed")
                t1 = .5 * (g - f) - lam * initslope
                t2 = .5 * (gprev - f) - lamprev * initslope
                a = (t1 / (lam * lam) - t2 / (lamprev * lamprev)) / (lam - lamprev)
                bq = (-lamprev * t1 / (lam * lam) + lam * t2 / (lamprev * lamprev)) / (lam - lamprev)
                d = max(bq * bq - 3 * a * initslope, 0.0)
                lamtemp = -initslope / (2.0 * bq) if a == 0.0 else (-bq + d ** 0.5) / (3.0 * a)
                lamprev, gprev = lam, g
                lam = .1 * lam if lamtemp <= .1 * lam else min(lamtemp, .5 * lam)
                G = trial(lam)
                gnorm = ops.norm(G)
                g = gnorm * gnorm
        x = x - lam * y
        F = G
        fnorms.append(gnorm)
    return x, fnorms, records


def stokes_schur_solver(J, Ablock, Bt, apply_Adiag_inv, apply_Ddiag_inv, nu, ops, outer_rtol=1e-2, inner_rtol=1e-2):
    """Returns solve(x, F): FGMRES on J with the upper Schur preconditioner; x = [u (nu entries) ; p]."""
    def pc(r):
        ru, rp = r[:nu], r[nu:]
        p = -apply_Ddiag_inv(rp)                              # S^-1 ~ -diag(D)^-1
        rhs = ru - Bt(p)
        u, its = cg(Ablock, apply_Adiag_inv, rhs, ops, rtol=inner_rtol)
        pc.inner.append(its)
        return cat(u, p)

    def cat(u, p):
        if isinstance(u, np.ndarray):
            return np.concatenate([u, p])
        import torch
        return torch.cat([u, p])

    pc.inner = []

    def solve(x, F):
        pc.inner = []
        y, hist = fgmres(J, pc, F, ops, rtol=outer_rtol)
        return y, dict(outer=len(hist) - 1, inner=list(pc.inner), hist=hist)

    return solve


def run_stokes_gpu(ctx, lib, ol, dist, rank, world, procs, ne_per_gpu=(4, 4, 4), p=3, prm=None, newton_rtol=1e-7, morph=None, return_solution=False):
    """The whole configs[3] solve on the job's partition through the C ABI.  Returns (on every rank) a dict with the Newton / FGMRES /
    CG iteration record, the final residual norm and the error against the manufactured solution (Exact1, stokesexact.py:95-103)."""
    import torch
    prm = prm or dict(B=1.0, eps=0.1, p=1.2)
    morph = morph or {}
    gne = tuple(n * m for n, m in zip(ne_per_gpu, procs))
    mu = lib.BoxMesh(gne, p, procs=procs, rank=rank, **morph)
    mp = lib.BoxMesh(gne, p - 1, procs=procs, rank=rank, dirichlet=0, **morph)
    o = ol.OracleStokes(mu, mp, exact=1, **prm)                  # the user callbacks of the example: forcing, Dirichlet data, exact solution
    rule = lib.Rule.preferred(ctx, p)
    efsu, efsp = lib.EFS(ctx, rule, p), lib.EFS(ctx, rule, p - 1)
    su, sp = lib.Space.from_mesh(ctx, mu, bs=3), lib.Space.from_mesh(ctx, mp, bs=1)
    su9 = lib.Space.from_mesh(ctx, mu, bs=9)                     # carries the 3x3 diagonal blocks through the reverse ghost exchange
    op = lib.Op(ctx, lib.OP_STOKES, lib.StokesParam(prm["B"], prm["eps"], prm["p"], 0.0), rule, [su, sp], [efsu, efsp], mu.coords)
    op.set_forcing(o.forcing(op.quadrature_points()))
    op.set_dirichlet(0, o.dirichlet_cache())
    nu, npr = mu.n * 3, mp.n
    dev = torch.device("cuda", torch.cuda.current_device())
    ops = TorchDistOps(dist, world)

    def split(x):
        return x[:nu], x[nu:]

    def function(x):
        y = torch.empty_like(x)
        op.function(list(split(x)), list(split(y)))
        return y

    def blockmult(block):
        def f(x):
            y = torch.zeros(nu + npr, dtype=torch.float64, device=dev)
            op.mult(list(split(x)), list(split(y)), block=block)
            return y
        return f

    J = blockmult(None)

    def Ablock(u):
        yu, yp = torch.empty_like(u), torch.empty(npr, dtype=torch.float64, device=dev)
        op.mult([u, torch.zeros(npr, dtype=torch.float64, device=dev)], [yu, yp], block=lib.STOKES_A)
        return yu

    def Bt(pv):
        yu, yp = torch.empty(nu, dtype=torch.float64, device=dev), torch.empty_like(pv)
        op.mult([torch.zeros(nu, dtype=torch.float64, device=dev), pv], [yu, yp], block=lib.STOKES_Bt)
        return yu

    # patterns once; diagonals per Newton step
    arp, aci = op.q1_pattern(mu.l2g)
    drp, dci = op.dense_pattern(1, mp.l2g)

    def row_diag(rowptr, colidx, vals, mesh, bs2):
        """Diagonal entries/blocks of the owned + ghost rows of a local CSR, summed onto the owners."""
        rp, ci = torch.from_numpy(rowptr.astype(np.int64)).to(dev), torch.from_numpy(colidx.astype(np.int64)).to(dev)
        nrows = len(rowptr) - 1
        rows = torch.repeat_interleave(torch.arange(nrows, device=dev), rp[1:] - rp[:-1])
        rowglobal = torch.cat([mesh.rstart + torch.arange(mesh.n, device=dev), torch.from_numpy(mesh.l2g[mesh.nc:].astype(np.int64)).to(dev)])
        mask = ci == rowglobal[rows]
        assert int(mask.sum().item()) == nrows
        d = torch.zeros((mesh.nc + mesh.ngh, bs2), dtype=torch.float64, device=dev)
        rowlocal = torch.cat([torch.arange(mesh.n, device=dev), mesh.nc + torch.arange(mesh.ngh, device=dev)])
        d[rowlocal[rows[mask]]] = vals.reshape(len(colidx), bs2)[mask]
        return d

    def make_solver(x):
        xu, xp = split(x)
        avals = torch.empty(len(aci) * 9, dtype=torch.float64, device=dev)
        op.assemble_q1([xu, xp], mu.l2g, None, None, vals=avals)
        dloc = row_diag(arp, aci, avals, mu, 9)
        dg = torch.empty(mu.n * 9, dtype=torch.float64, device=dev)
        lib.chk(lib.load().dB200SpaceLocalToGlobal(su9.h, lib.ptr(dloc), lib.ptr(dg), lib.INSERT_VALUES))
        Ainv = torch.linalg.inv(dg.reshape(mu.n, 3, 3))
        dvals = torch.empty(len(dci), dtype=torch.float64, device=dev)
        op.assemble_dense(lib.DENSE_STOKES_D, [xu, xp], mp.l2g, drp, dci, vals=dvals)
        dl = row_diag(drp, dci, dvals, mp, 1)
        dd = torch.empty(mp.n, dtype=torch.float64, device=dev)
        lib.chk(lib.load().dB200SpaceLocalToGlobal(sp.h, lib.ptr(dl), lib.ptr(dd), lib.INSERT_VALUES))
        return stokes_schur_solver(J, Ablock, Bt, lambda r: torch.einsum("nij,nj->ni", Ainv, r.reshape(-1, 3)).reshape(-1), lambda r: r / dd, nu, ops)

    solvers = {}

    def solve(x, F):
        return make_solver(x)(x, F)

    x0 = torch.zeros(nu + npr, dtype=torch.float64, device=dev)
    x, fnorms, recs = newton_bt(function, J, solve, x0, ops, rtol=newton_rtol)
    # error against the manufactured solution at the owned nodes
    ue, pe = o.exact_solution(mu.node_xyz[:mu.n])[0].reshape(-1), o.exact_solution(mp.node_xyz[:mp.n])[1]
    xu, xp = split(x)
    eu = torch.tensor([float((xu.cpu().numpy() - ue).__abs__().max()), float(np.abs(ue).max())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(eu, op=dist.ReduceOp.MAX)
    res = dict(newton_its=len(fnorms) - 1, fnorms=fnorms, outer=[r["outer"] for r in recs], inner_total=[sum(r["inner"]) for r in recs],
               velocity_err_rel=float(eu[0].item() / eu[1].item()), dofs=int(mu.nglobal * 3 + mp.nglobal))
    if return_solution:                                          # owned parts of the Newton solution and the checker that knows the exact solution
        res["xu"], res["xp"], res["oracle"] = xu.cpu().numpy().copy(), xp.cpu().numpy().copy(), o
    op.close()
    return res


def run_stokes_oracle(lib_or_box, ol, ne=(3, 3, 3), p=3, prm=None, newton_rtol=1e-7, morph=None):
    """The same solve driven by the CPU oracle on a single domain (numpy vectors): the reference side of the iteration-count parity."""
    import scipy.sparse as sp
    from dohp_b200.box import build_box_space, dense_pattern, q1_pattern
    prm = prm or dict(B=1.0, eps=0.1, p=1.2)
    mu = build_box_space(ne, p, morph=morph)
    mp = build_box_space(ne, p - 1, dirichlet=(), morph=morph)
    o = ol.OracleStokes(mu, mp, exact=1, **prm)
    nu, npr = mu.n * 3, mp.n
    ops = NumpyOps()
    arp, aci = q1_pattern(mu, bs=3)
    drp, dci = dense_pattern(mp)

    def function(x):
        return np.concatenate(o.function(np.ascontiguousarray(x[:nu]), np.ascontiguousarray(x[nu:])))

    def J(x):
        return np.concatenate(o.mult(np.ascontiguousarray(x[:nu]), np.ascontiguousarray(x[nu:]), block=0))

    def Ablock(u):
        return o.mult(np.ascontiguousarray(u), np.zeros(npr), block=lib_or_box.STOKES_A)[0]

    def Bt(pv):
        return o.mult(np.zeros(nu), np.ascontiguousarray(pv), block=lib_or_box.STOKES_Bt)[0]

    def solve(x, F):
        xu = np.ascontiguousarray(x[:nu])
        avals, _ = o.jacobian_q1_velocity(xu, arp, aci)
        A = sp.bsr_matrix(sp.csr_matrix((avals, aci, arp), shape=(nu, nu)), blocksize=(3, 3))
        A.sort_indices()
        rows = np.repeat(np.arange(mu.n), np.diff(A.indptr))
        Ainv = np.linalg.inv(A.data[A.indices == rows])
        D, _ = o.jacobian_pressure(xu, drp, dci)
        dd = sp.csr_matrix((D, dci, drp), shape=(npr, npr)).diagonal()
        s = stokes_schur_solver(J, Ablock, Bt, lambda r: np.einsum("nij,nj->ni", Ainv, r.reshape(-1, 3)).reshape(-1), lambda r: r / dd, nu, ops)
        return s(x, F)

    x, fnorms, recs = newton_bt(function, J, solve, np.zeros(nu + npr), ops, rtol=newton_rtol)
    ue = o.exact_solution(mu.node_xyz[:mu.n])[0].reshape(-1)
    return dict(newton_its=len(fnorms) - 1, fnorms=fnorms, outer=[r["outer"] for r in recs], inner_total=[sum(r["inner"]) for r in recs],
                velocity_err_rel=float(np.abs(x[:nu] - ue).max() / np.abs(ue).max()), dofs=nu + npr)
