"""Runs tests/test_gpu_zzz_stokes_golden.py's code path on the CPU with a mock of lib.Op built on the oracle (shapes and layouts of
the real ctypes wrapper): catches Python-level mistakes of the never-run GPU test."""
import sys
import numpy as np, scipy.sparse as sp
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import oracle_lib as ol
from dohp_b200 import lib
from dohp_b200.box import dense_pattern, q1_pattern
ol.build_oracle(ref=False)

class MockOp:
    def __init__(self, mu, mp, o):
        self.mu, self.mp, self.o = mu, mp, o
    def function(self, xs):
        return list(o2(self.o.function(xs[0], xs[1])))
    def mult(self, xs, block=None):
        b = {None: 0, lib.STOKES_A: 1, lib.STOKES_B: 2, lib.STOKES_Bt: 3}[block]
        return list(o2(self.o.mult(xs[0], xs[1], block=b)))
    def q1_pattern(self, l2g):
        return q1_pattern(self.mu, bs=1)
    def assemble_q1(self, state, l2g, rowptr, colidx):
        rp, ci = q1_pattern(self.mu, bs=3)
        vals, _ = self.o.jacobian_q1_velocity(state[0], rp, ci)
        n3 = self.mu.n * 3
        B = sp.bsr_matrix(sp.csr_matrix((vals, ci, rp), shape=(n3, n3)), blocksize=(3, 3)); B.sort_indices()
        assert np.array_equal(B.indptr, rowptr) and np.array_equal(B.indices, colidx)
        return B.data.reshape(-1).copy()
    def dense_pattern(self, field, l2g):
        return dense_pattern(self.mp)
    def assemble_dense(self, kind, state, l2g, rowptr, colidx):
        assert kind == lib.DENSE_STOKES_D
        return self.o.jacobian_pressure(state[0], rowptr, colidx)[0]
    def close(self):
        pass
def o2(t):
    return [np.array(t[0]), np.array(t[1])]

import test_gpu_zzz_stokes_golden as T
T.make_stokes = lambda ctx, mu, mp, o: MockOp(mu, mp, o)
from test_stokes_refout_replay import check, closing, fgmres_newton, parse
gs, gk, gi = parse("stokes-q3q2-fgmres")
P = T.GpuProblem(None, dict(B=1.0, eps=0.1, p=1.2))
snes, ksps, inners = fgmres_newton(P)
check(snes, gs, "SNES")
for it, (k, g) in enumerate(zip(ksps, gk)):
    check(k, g, "KSP %d" % it)
assert inners == gi
closing(P, "stokes-q3q2-fgmres")
print("mock GPU replay: OK", len(snes), sum(map(len, ksps)), sum(map(len, inners)))
