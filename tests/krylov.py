"""Minimal Newton-Krylov harness shared by the CPU oracle and the GPU path (test infrastructure): left-preconditioned
restarted GMRES in the style of PETSc's default KSP (classical Gram-Schmidt, preconditioned residual norm, rtol 1e-5)
and a full-step Newton iteration as SNES would run the examples (rtol 1e-10 on ||F||, ellip.c:705).  Both paths run the
SAME code with different operator callbacks, so iteration counts and residual histories can be compared."""
import numpy as np


def gmres(A, M, b, rtol=1e-5, restart=30, maxit=200):
    """Solve A x = b with left preconditioner M (callables).  Returns x and the list of preconditioned residual norms."""
    x = np.zeros_like(b)
    hist = []
    r = M(b - A(x))
    beta0 = np.linalg.norm(r)
    hist.append(beta0)
    if beta0 == 0:
        return x, hist
    it = 0
    while it < maxit:
        beta = np.linalg.norm(r)
        V = [r / beta]
        H = np.zeros((restart + 1, restart))
        g = np.zeros(restart + 1)
        g[0] = beta
        cs, sn = np.zeros(restart), np.zeros(restart)
        k_used = 0
        for k in range(restart):
            w = M(A(V[k]))
            for i in range(k + 1):
                H[i, k] = np.dot(w, V[i])
            for i in range(k + 1):
                w = w - H[i, k] * V[i]
            H[k + 1, k] = np.linalg.norm(w)
            V.append(w / H[k + 1, k] if H[k + 1, k] > 0 else w)
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
            it += 1
            k_used = k + 1
            hist.append(abs(g[k + 1]))
            if abs(g[k + 1]) <= rtol * beta0 or it >= maxit:
                break
        y = np.linalg.solve(np.triu(H[:k_used, :k_used]), g[:k_used])
        for i in range(k_used):
            x = x + y[i] * V[i]
        if hist[-1] <= rtol * beta0 or it >= maxit:
            break
        r = M(b - A(x))
    return x, hist


def newton(function, mult, make_pc, x0, rtol=1e-10, maxit=20, ksp_rtol=1e-5):
    """x <- x - J(x)^-1 F(x).  `function(x)` must leave the stash at x so that `mult` is J(x); `make_pc(x)` returns M."""
    x = x0.copy()
    fnorms, ksp_its = [], []
    F = function(x)
    f0 = np.linalg.norm(F)
    fnorms.append(f0)
    for _ in range(maxit):
        if fnorms[-1] <= rtol * f0:
            break
        M = make_pc(x)
        dx, hist = gmres(mult, M, F, rtol=ksp_rtol)
        ksp_its.append(len(hist) - 1)
        x = x - dx
        F = function(x)
        fnorms.append(np.linalg.norm(F))
    return x, fnorms, ksp_its
