"""PETSc-faithful Krylov / Newton drivers for replaying the reference's golden solver histories (test infrastructure).

The reference examples are driven by PETSc SNES/KSP (src/fs/tests/ellip.c:700-747, elast.c:518-536) and their goldens
(src/fs/tests/refout/*.refout) print what -snes_monitor_short / -ksp_monitor_short saw.  PETSc is not installed here, so the
algorithms PETSc runs with the options of src/fs/tests/CMakeLists.txt:53-60 are restated:

* KSPGMRES defaults: left preconditioning, restart 30, classical Gram-Schmidt without refinement, the monitored quantity is the
  preconditioned residual norm (the recurrence estimate inside a cycle, the recomputed residual at the start of every cycle, so
  the entry of a restart iteration is the recomputed one), KSPConvergedDefault: rnorm <= max(rtol * rnorm_0, abstol), rtol 1e-5.
* KSPCG defaults: left preconditioning, the monitored quantity is ||B r||_2 (KSP_NORM_PRECONDITIONED).
* SNESNEWTONLS with -snes_linesearch_type basic: x <- x - J^-1 F, convergence ||F|| <= max(rtol ||F_0||, abstol) with the
  tolerances the examples set (SNESSetTolerances(snes, default 1e-50, 1e-10, default 1e-8, ...), ellip.c:705) incl. the step test
  ||dx|| < stol ||x||.
* PCILU on a BAIJ matrix with the natural ordering: block ILU(0) (MatILUFactorSymbolic_SeqBAIJ / MatLUFactorNumeric_SeqBAIJ_N).
* the format of the short monitors (PetscViewerASCIIPrintf in KSPMonitorDefaultShort / SNESMonitorDefaultShort): %g above 1e-9,
  %5.3e above 1e-11, "< 1.e-11" below.

Both the CPU oracle and the GPU library are driven through these with their own operator callbacks."""
import re

import numpy as np


def fmt_short(x):
    """The number as the short monitors print it."""
    if x > 1e-9:
        return "%g" % x
    if x > 1e-11:
        return "%5.3e" % x
    return "< 1.e-11"


def gmres(A, M, b, rtol=1e-5, abstol=1e-50, restart=30, maxit=10000):
    """Left-preconditioned restarted GMRES, zero initial guess.  Returns x and the monitored norms (entry k = iteration k)."""
    x = np.zeros_like(b)
    hist = []
    its = 0
    r = M(b)                                   # zero initial guess: B (b - A 0)
    rnorm0 = None
    while True:
        beta = np.linalg.norm(r)
        if rnorm0 is None:
            rnorm0 = beta
            hist.append(beta)
        else:
            hist[-1] = beta                    # the restart iteration is monitored with the recomputed residual
        if beta <= max(rtol * rnorm0, abstol) or its >= maxit:
            return x, hist
        V = [r / beta]
        H = np.zeros((restart + 1, restart))
        g = np.zeros(restart + 1)
        g[0] = beta
        cs, sn = np.zeros(restart), np.zeros(restart)
        k_used, done = 0, False
        for k in range(restart):
            w = M(A(V[k]))
            hk = np.array([np.dot(w, v) for v in V])          # classical Gram-Schmidt: all dots against the unmodified w (VecMDot)
            for i, v in enumerate(V):                         # VecMAXPY
                w = w - hk[i] * v
            H[:k + 1, k] = hk
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
            its += 1
            k_used = k + 1
            hist.append(abs(g[k + 1]))
            if abs(g[k + 1]) <= max(rtol * rnorm0, abstol) or its >= maxit:
                done = True
                break
        y = np.linalg.solve(np.triu(H[:k_used, :k_used]), g[:k_used])
        for i in range(k_used):
            x = x + y[i] * V[i]
        if done:
            return x, hist
        r = M(b - A(x))


def cg(A, M, b, rtol=1e-5, abstol=1e-50, maxit=10000):
    """KSPCG: preconditioned conjugate gradients, monitored norm ||B r||_2, zero initial guess."""
    x = np.zeros_like(b)
    r = b.copy()
    z = M(r)
    dp = np.linalg.norm(z)
    hist = [dp]
    dp0 = dp
    beta_old, p = None, None
    for _ in range(maxit):
        if dp <= max(rtol * dp0, abstol):
            break
        beta = np.dot(r, z)
        p = z.copy() if p is None else z + (beta / beta_old) * p
        w = A(p)
        a = beta / np.dot(p, w)
        x = x + a * p
        r = r - a * w
        z = M(r)
        dp = np.linalg.norm(z)
        hist.append(dp)
        beta_old = beta
    return x, hist


def linesearch_bt(function, mult, x, F, fnorm, y, alpha=1e-4, minlambda=1e-12, max_its=40):
    """SNESLineSearchApply_BT (PETSc's default Newton line search: cubic backtracking) for the update x - lambda y.  `mult` is the
    Jacobian at x (used for the initial slope before any trial evaluation overwrites the stash).  Returns (x_new, F_new, ||F_new||, lambda)."""
    f = fnorm * fnorm
    initslope = np.dot(F, mult(y))
    if initslope > 0:
        initslope = -initslope
    if initslope == 0:
        initslope = -1.0
    lam = 1.0
    w = x - lam * y
    G = function(w)
    gnorm = np.linalg.norm(G)
    g = gnorm * gnorm
    if .5 * g <= .5 * f + lam * alpha * initslope:
        return w, G, gnorm, lam
    lamtemp = -initslope / (g - f - 2.0 * lam * initslope)      # quadratic fit
    lamprev, gprev = lam, g
    lam = .1 * lam if lamtemp <= .1 * lam else min(lamtemp, .5 * lam)
    w = x - lam * y
    G = function(w)
    gnorm = np.linalg.norm(G)
    g = gnorm * gnorm
    if .5 * g < .5 * f + lam * alpha * initslope:
        return w, G, gnorm, lam
    for _ in range(max_its):                                     # cubic fits
        if lam <= minlambda:
            raise RuntimeError("line search failed")
        t1 = .5 * (g - f) - lam * initslope
        t2 = .5 * (gprev - f) - lamprev * initslope
        a = (t1 / (lam * lam) - t2 / (lamprev * lamprev)) / (lam - lamprev)
        b = (-lamprev * t1 / (lam * lam) + lam * t2 / (lamprev * lamprev)) / (lam - lamprev)
        d = max(b * b - 3 * a * initslope, 0.0)
        lamtemp = -initslope / (2.0 * b) if a == 0.0 else (-b + np.sqrt(d)) / (3.0 * a)
        lamprev, gprev = lam, g
        lam = .1 * lam if lamtemp <= .1 * lam else min(lamtemp, .5 * lam)
        w = x - lam * y
        G = function(w)
        gnorm = np.linalg.norm(G)
        g = gnorm * gnorm
        if .5 * g < .5 * f + lam * alpha * initslope:
            return w, G, gnorm, lam
    raise RuntimeError("line search failed")


def newton(function, mult, make_pc, x0, linesearch="bt", rtol=1e-10, abstol=1e-50, stol=1e-8, maxit=50, ksp=gmres, ksp_rtol=1e-5):
    """SNESNEWTONLS.  `function(x)` leaves the stash at x so that `mult` is J(x); `make_pc(x)` returns the preconditioner application
    for the Jacobian at x.  linesearch: "bt" (PETSc's default) or "basic" (-snes_linesearch_type basic: the full step).
    Returns x, the SNES function norms and the KSP histories."""
    x = x0.copy()
    F = function(x)
    fnorms, ksp_hists = [np.linalg.norm(F)], []
    f0 = fnorms[0]
    for _ in range(maxit):
        if fnorms[-1] <= max(rtol * f0, abstol):
            break
        M = make_pc(x)
        y, hist = ksp(mult, M, F, rtol=ksp_rtol)
        ksp_hists.append(hist)
        if linesearch == "basic":
            x = x - y
            F = function(x)
            fn, lam = np.linalg.norm(F), 1.0
        else:
            x, F, fn, lam = linesearch_bt(function, mult, x, F, fnorms[-1], y)
        fnorms.append(fn)
        if lam * np.linalg.norm(y) < stol * np.linalg.norm(x) and fn > max(rtol * f0, abstol):
            break                                              # SNES_CONVERGED_SNORM_RELATIVE
    return x, fnorms, ksp_hists


def newton_basic(function, mult, make_pc, x0, **kw):
    return newton(function, mult, make_pc, x0, linesearch="basic", **kw)


def block_ilu0(rowptr, colidx, vals, bs):
    """ILU(0) of a block CSR matrix (sorted columns) in the matrix's own ordering; returns the application z = (LU)^-1 r.
    vals: [nnzb, bs, bs].  The factor keeps L (unit block diagonal) and U with the inverted diagonal blocks, as PETSc's BAIJ does.
    Compiled restatement oracle/petsc_ilu.c (liboracle.so); `block_ilu0_numpy` below is the same algorithm in numpy (slow reference)."""
    import oracle_lib as ol
    lib = ol.oracle()
    n = len(rowptr) - 1
    rp, ci = np.ascontiguousarray(rowptr, dtype=np.int32), np.ascontiguousarray(colidx, dtype=np.int32)
    a = np.array(vals, dtype=np.float64).reshape(-1, bs, bs).copy()
    dinv = np.zeros((max(n, 1), bs, bs))
    err = lib.or_bilu0_factor(n, bs, ol.iptr(rp), ol.iptr(ci), ol.dptr(a), ol.dptr(dinv))
    if err:
        raise RuntimeError("or_bilu0_factor: error %d (1 zero pivot, 2 unsorted row / missing diagonal)" % err)

    def solve(r):
        r = np.ascontiguousarray(r, dtype=np.float64)
        y = np.empty(n * bs)
        assert r.size == n * bs and lib.or_bilu0_solve(n, bs, ol.iptr(rp), ol.iptr(ci), ol.dptr(a), ol.dptr(dinv), ol.dptr(r), ol.dptr(y)) == 0
        return y

    return solve


def block_ilu0_numpy(rowptr, colidx, vals, bs):
    """The algorithm of block_ilu0 in plain numpy."""
    n = len(rowptr) - 1
    a = np.array(vals, dtype=np.float64).reshape(-1, bs, bs).copy()
    pos = [dict((int(c), int(k)) for k, c in zip(range(rowptr[i], rowptr[i + 1]), colidx[rowptr[i]:rowptr[i + 1]])) for i in range(n)]
    diag = np.array([pos[i][i] for i in range(n)])
    dinv = np.zeros((n, bs, bs))
    for i in range(n):
        for k in range(rowptr[i], rowptr[i + 1]):
            c = int(colidx[k])
            if c >= i:
                break
            a[k] = a[k] @ dinv[c]                               # L_ic = A_ic U_cc^-1
            lic = a[k]
            for kk in range(diag[c] + 1, rowptr[c + 1]):        # row c of U, columns > c
                j = int(colidx[kk])
                p = pos[i].get(j)
                if p is not None:
                    a[p] = a[p] - lic @ a[kk]
        dinv[i] = np.linalg.inv(a[diag[i]])

    def solve(r):
        y = np.array(r, dtype=np.float64).reshape(n, bs).copy()
        for i in range(n):
            for k in range(rowptr[i], diag[i]):
                y[i] -= a[k] @ y[colidx[k]]
        for i in range(n - 1, -1, -1):
            for k in range(diag[i] + 1, rowptr[i + 1]):
                y[i] -= a[k] @ y[colidx[k]]
            y[i] = dinv[i] @ y[i]
        return y.reshape(-1)

    return solve


def parse_refout(path):
    """-> (snes norms as printed, [ksp histories as printed]) ; a file without SNES lines gives one KSP history."""
    snes, ksps = [], []
    for line in open(path):
        m = re.match(r"\s*(\d+) SNES Function norm (.*\S)\s*$", line)
        if m:
            snes.append(m.group(2).strip())
            continue
        m = re.match(r"\s*(\d+) KSP Residual norm (.*\S)\s*$", line)
        if m:
            if int(m.group(1)) == 0:
                ksps.append([])
            ksps[-1].append(m.group(2).strip())
    return snes, ksps


def same_as_printed(value, printed, last_digit=1):
    """Does `value` print as the golden does, allowing `last_digit` units in the last printed place (PETSc's run and ours differ in
    summation order and LU pivoting: about 1e-12 relative in the operator, amplified along a Krylov history)?"""
    if printed.startswith("<"):
        return value <= 1.0000005e-11
    mine = fmt_short(value)
    if mine == printed:
        return True
    if mine.startswith("<"):
        return False
    ref = float(printed)
    if ref == 0:
        return value == 0
    nd = 6 if ref > 1e-9 else 4                                # %g prints 6 significant digits (trailing zeros dropped), %5.3e four
    ulp = 10.0 ** (np.floor(np.log10(abs(ref))) - (nd - 1))
    return abs(value - ref) <= (last_digit + 0.5) * ulp
