"""stokes-q3q2-fgmres.refout replayed with the GPU LIBRARY as the operator (through the C ABI): the matrix-free residual, the coupled
Jacobian and its A / B^T blocks, the embedded-Q1 velocity block Ap (dB200OpAssembleQ1) and the pressure block D (dB200OpAssembleDense) on
the reference's default RCM numbering (dohp_b200.mesh.genblock_hex_space), driven by the PETSc-faithful Newton / FGMRES / field-split
restatement of tests/test_stokes_refout_replay.py (which does the same on the CPU oracle, verified).  Every SNES norm, all 80 FGMRES
residual norms to the printed digits, all 70 inner CG iteration counts and the closing lines of the golden must come out.

STATUS: written after this round's GPU budget was spent, so it has NEVER RUN ON A DEVICE.  It is marked xfail(strict=False) for that
reason alone: a pass shows up as XPASS, a failure cannot stop the rest of the suite.  Every library call it makes is exercised by the
verified tests (test_gpu_elast_stokes.py, test_gpu_q1.py, test_gpu_mesh.py) against the oracle to 1e-12, and its own Python path runs on the
CPU against an oracle-backed mock of lib.Op with the same shapes and layouts (tools/mock_gpu_stokes_replay.py: passes, also with random relative noise of 1e-13 on every operator output -- the GPU differs from the
oracle by about 1e-15)."""
import numpy as np
import pytest
import scipy.sparse as sp

import oracle_lib as ol
import petsc_like as pl
from dohp_b200 import lib
from dohp_b200.mesh import genblock_hex_space
from test_gpu_elast_stokes import make_stokes
from test_stokes_refout_replay import check, closing, fgmres_newton, parse

pytestmark = [pytest.mark.gpu, pytest.mark.xfail(strict=False, reason="never run on a device: written after the round's GPU budget was spent")]


class GpuProblem:
    """the interface of test_stokes_refout_replay.Problem with libdohp_b200.so as the operator"""

    def __init__(self, ctx, prm):
        self.mu, self.mp = genblock_hex_space((4, 4, 4), 3), genblock_hex_space((4, 4, 4), 2, dirichlet=False)
        self.o = ol.OracleStokes(self.mu, self.mp, exact=1, **prm)          # the example's user callbacks: forcing, Dirichlet data, exact solution
        self.op = op = make_stokes(ctx, self.mu, self.mp, self.o)
        self.nu, self.np = nu, npr = self.mu.n * 3, self.mp.n
        self.arp, self.aci = op.q1_pattern(self.mu.l2g)                     # block CSR of the velocity space
        self.drp, self.dci = op.dense_pattern(1, self.mp.l2g)
        self.function = lambda x: np.concatenate(op.function([np.ascontiguousarray(x[:nu]), np.ascontiguousarray(x[nu:])]))
        self.J = lambda x: np.concatenate(op.mult([np.ascontiguousarray(x[:nu]), np.ascontiguousarray(x[nu:])]))
        self.A = lambda u: op.mult([np.ascontiguousarray(u), None], block=lib.STOKES_A)[0].copy()
        self.Bt = lambda p: op.mult([None, np.ascontiguousarray(p)], block=lib.STOKES_Bt)[0].copy()

    def preconditioners(self, x):
        state = [np.ascontiguousarray(x[:self.nu]), np.zeros(self.np)]
        avals = self.op.assemble_q1(state, self.mu.l2g, self.arp, self.aci).reshape(-1, 3, 3)
        Ap = sp.bsr_matrix((avals, self.aci, self.arp), shape=(self.nu, self.nu))
        Ap.sort_indices()
        D = self.op.assemble_dense(lib.DENSE_STOKES_D, state, self.mp.l2g, self.drp, self.dci)
        Dp = sp.csr_matrix((D, self.dci, self.drp), shape=(self.np, self.np))
        Dp.sort_indices()
        return pl.block_ilu0(Ap.indptr, Ap.indices, Ap.data, 3), pl.block_ilu0(Dp.indptr, Dp.indices, Dp.data.reshape(-1, 1, 1), 1)


def test_stokes_q3q2_fgmres_golden_on_the_gpu_library():
    import torch
    ol.build_oracle(ref=False)
    ctx = lib.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    try:
        gs, gk, gi = parse("stokes-q3q2-fgmres")
        P = GpuProblem(ctx, dict(B=1.0, eps=0.1, p=1.2))
        snes, ksps, inners = fgmres_newton(P)
        check(snes, gs, "SNES")
        assert [len(k) for k in ksps] == [len(k) for k in gk]
        for it, (k, g) in enumerate(zip(ksps, gk)):
            check(k, g, "KSP of Newton step %d" % it)
        assert inners == gi
        closing(P, "stokes-q3q2-fgmres")
        P.op.close()
    finally:
        ctx.close()
