"""Prints the CPU-oracle replay of the reference's Stokes goldens in PETSc's own output format, so that it can be diffed against
src/fs/tests/refout/stokes-q3q2-{fgmres,scr}.refout (copies under tests/golden/refout/).

    python tools/stokes_replay_print.py fgmres > profiles/r4_stokes_q3q2_fgmres_replay.out
    python tools/stokes_replay_print.py scr    > profiles/r4_stokes_q3q2_scr_replay.out"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import oracle_lib as ol                                    # noqa: E402
import petsc_like as pl                                    # noqa: E402
import test_stokes_refout_replay as T                      # noqa: E402

ol.build_oracle(ref=False)
ol.oracle().or_set_num_threads(os.cpu_count() or 1)
which = sys.argv[1] if len(sys.argv) > 1 else "fgmres"
prm = dict(B=1.0, eps=0.1, p=1.2) if which == "fgmres" else dict(B=1.0, eps=1.0, p=2.0)
P = T.Problem(prm, "rcm")
out = []
# the two header lines (stokes.c:1117-1135): residual at the interpolated exact solution; J(soln) soln + F(0)
su, sp_ = P.o.nodal_solution()
soln = np.concatenate([su, sp_])
F0 = P.function(np.zeros(P.nu + P.np))
out.append("Norm of discrete residual for exact solution %g" % np.linalg.norm(P.function(soln)))
out.append("Norm of discrete linear residual at exact solution %g" % np.linalg.norm(P.J(soln) + F0))
if which == "fgmres":
    snes, ksps, inners = T.fgmres_newton(P)
    for it, fn in enumerate(snes):
        out.append("%3d SNES Function norm %s " % (it, pl.fmt_short(fn)))
        if it < len(ksps):
            for k, rn in enumerate(ksps[it]):
                out.append("  %3d KSP Residual norm %s " % (k, pl.fmt_short(rn)))
                if k < len(inners[it]):
                    out.append("      Linear solve converged due to CONVERGED_RTOL iterations %d" % inners[it][k])
else:
    nu = P.nu
    x = np.zeros(P.nu + P.np)
    F = P.function(x)
    Mu, Mp = P.preconditioners(x)
    inner = []

    def Ainv(b):
        y, h = pl.gmres(P.A, Mu, b, rtol=1e-5)
        inner.append(len(h) - 1)
        return y
    yu = Ainv(F[:nu])
    yp, hist = pl.cg(lambda p: -P.B(Ainv(P.Bt(p))), Mp, F[nu:] - P.B(yu), rtol=1e-5)
    yu = Ainv(F[:nu] - P.Bt(yp))
    y = np.concatenate([yu, yp])
    y[nu:] -= y[nu:].mean()
    P.x = x - y
    P.F = P.function(P.x)
    out.append("  0 SNES Function norm %s " % pl.fmt_short(np.linalg.norm(F)))
    out.append("      Linear solve converged due to CONVERGED_RTOL iterations %d" % inner[0])
    out.append("      Residual norms for fieldsplit_p_ solve.")
    for k, rn in enumerate(hist):
        out.append("    %3d KSP Residual norm %s " % (k, pl.fmt_short(rn)))
        out.append("      Linear solve converged due to CONVERGED_RTOL iterations %d" % inner[1 + k])
    out.append("  1 SNES Function norm %s " % pl.fmt_short(np.linalg.norm(P.F)))
# closing lines (stokes.c:1152-1170)
xu, xp = P.x[:P.nu].copy(), P.x[P.nu:].copy()
xp -= xp.mean()
spn = sp_ - sp_.mean()
en, gn, pn = P.o.error_norms(xu, xp)
for label, v in (("Algebraic residual       ", ol.norms_algebraic_scaled(P.F)), ("Interpolation residual   ", ol.norms_algebraic_scaled(np.concatenate([xu - su, xp - spn]))),
                 ("Pointwise velocity error ", en), ("Pointwise gradient error ", gn), ("Pointwise pressure error ", pn)):
    out.append("%s |x|_1 %8.2e  |x|_2 %8.2e  |x|_inf %8.2e" % (label, v[0], v[1], v[2]))
print("\n".join(out))
