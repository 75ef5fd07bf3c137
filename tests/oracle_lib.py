"""ctypes access to the TEST-ONLY oracle libraries.

* ``oracle/liboracle.so``  -- this repo's plain-C restatement of the reference hot path (oracle/oracle.c)
* ``oracle/_ref/libref_driver.so`` -- the reference's own dJacobi layer compiled unmodified (oracle/Makefile)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REFERENCE = "/root/reference"

APPLY_INTERP, APPLY_INTERP_TRANSPOSE, APPLY_GRAD, APPLY_GRAD_TRANSPOSE = 40001, 40002, 40003, 40004
INSERT_VALUES, ADD_VALUES = 1, 2
FAST, SPARSE, SELF = 0, 1, 2
GAUSS, LOBATTO = 0, 1
HOMOGENEOUS, INHOMOGENEOUS = 0, 1

c_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
c_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build_oracle(ref=None):
    """Compile the oracle (and, when /root/reference is present, oracle/_ref).  Building the checker is not using it."""
    targets = ["oracle"]
    if ref is None:
        ref = os.path.isdir(REFERENCE)
    if ref:
        targets.append("ref")
        if os.path.exists(os.path.join(ROOT, "dohp_b200", "libdohp_b200.so")):
            targets.append("refbind")        # the reference-side dJacobi binding (integration/jacobi_b200.c) on the reference's own test program
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR] + targets)


def have_ref():
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libref_driver.so"))


_oracle = None
_ref = None


def dptr(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_double))


def iptr(a):
    if a is None:
        return None
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_int))


class OrSpace(C.Structure):
    _fields_ = [("nelem", C.c_int), ("bs", C.c_int), ("n", C.c_int), ("nc", C.c_int), ("ngh", C.c_int),
                ("e2l", C.POINTER(C.c_int))]


class OrEllipParam(C.Structure):
    _fields_ = [("p", C.c_double), ("epsilon", C.c_double), ("lambda_", C.c_double), ("exact", C.c_int),
                ("a", C.c_double), ("b", C.c_double), ("c", C.c_double)]


class OrElastParam(C.Structure):
    _fields_ = [("lambda_", C.c_double), ("mu", C.c_double), ("exact", C.c_int),
                ("a", C.c_double), ("b", C.c_double), ("c", C.c_double), ("scale", C.c_double)]


class OrStokesParam(C.Structure):
    _fields_ = [("B", C.c_double), ("eps", C.c_double), ("p", C.c_double), ("nu", C.c_double), ("exact", C.c_int),
                ("a", C.c_double), ("b", C.c_double), ("c", C.c_double), ("scale", C.c_double)]


def oracle():
    global _oracle
    if _oracle is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build_oracle()
        _oracle = C.CDLL(path)
    return _oracle


def ref():
    global _ref
    if _ref is None:
        path = os.path.join(ORACLE_DIR, "_ref", "libref_driver.so")
        if not os.path.exists(path):
            build_oracle()
        _ref = C.CDLL(path)
        _ref.RefInit(0)
    return _ref


def _check(err, what):
    if err:
        raise RuntimeError("%s failed with error %d" % (what, err))


class OracleEFS:
    """(basis degree, rule) pair of the restatement: mirrors dJacobiGetEFS on rules from dQuadratureGetRules."""

    def __init__(self, bdeg, rorder=None, method=FAST, family=GAUSS):
        lib = oracle()
        bdeg = [bdeg] * 3 if np.isscalar(bdeg) else list(bdeg)
        if rorder is None:
            rorder = [2 * b for b in bdeg]          # dFSGetPreferredQuadratureRuleSet, fsruleset.c:32-37
        rorder = [rorder] * 3 if np.isscalar(rorder) else list(rorder)
        self.h = C.c_void_p()
        _check(lib.or_efs_create((C.c_int * 3)(*bdeg), (C.c_int * 3)(*rorder), method, family, C.byref(self.h)), "or_efs_create")
        P = (C.c_int * 3)()
        Q = (C.c_int * 3)()
        lib.or_efs_get_sizes(self.h, P, Q)
        self.P, self.Q = list(P), list(Q)
        self.np, self.nq = int(np.prod(self.P)), int(np.prod(self.Q))
        self.method, self.family = method, family

    def tables(self, d):
        P, Q = self.P[d], self.Q[d]
        interp, deriv, node, weight = np.zeros((Q, P)), np.zeros((Q, P)), np.zeros(P), np.zeros(P)
        oracle().or_efs_get_tables(self.h, d, dptr(interp), dptr(deriv), dptr(node), dptr(weight))
        return interp, deriv, node, weight

    def rule(self, d):
        c, w = np.zeros(self.Q[d]), np.zeros(self.Q[d])
        oracle().or_efs_get_rule(self.h, d, dptr(c), dptr(w))
        return c, w

    def patches(self):
        n, ps = C.c_int(), C.c_int()
        ind, wt = C.POINTER(C.c_int)(), C.POINTER(C.c_double)()
        oracle().or_efs_get_patches(self.h, C.byref(n), C.byref(ps), C.byref(ind), C.byref(wt))
        N = n.value * ps.value
        return n.value, ps.value, np.ctypeslib.as_array(ind, (N,)).copy(), np.ctypeslib.as_array(wt, (N,)).copy()

    def apply(self, jinv, D, x, amode, imode=INSERT_VALUES, out=None, nelem=None):
        """Batched dEFSApply: x is [nelem, ...] element-contiguous."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        nelem = x.shape[0] if nelem is None else nelem
        shape = {APPLY_INTERP: (nelem, self.nq, D), APPLY_GRAD: (nelem, self.nq, D, 3),
                 APPLY_INTERP_TRANSPOSE: (nelem, self.np, D), APPLY_GRAD_TRANSPOSE: (nelem, self.np, D)}[amode]
        if out is None:
            out = np.zeros(shape)
        _check(oracle().or_efs_apply_batch(self.h, nelem, dptr(jinv), D, dptr(x), dptr(out), amode, imode), "or_efs_apply_batch")
        return out

    def compute_physical(self, jac):
        jac = np.ascontiguousarray(jac)
        nelem = jac.shape[0]
        jinv, jw = np.zeros((nelem, self.nq, 9)), np.zeros((nelem, self.nq))
        _check(oracle().or_rule_compute_physical_batch(self.h, nelem, dptr(jac), dptr(jinv), dptr(jw)), "or_rule_compute_physical")
        return jinv, jw

    def explicit_sparse(self, cjinv=None, eoffset=0):
        npatch, Q, qidx, _ = self.patches()
        maxP = self.np
        P = C.c_int()
        eidx = np.zeros(npatch * maxP, dtype=np.int32)
        interp, deriv = np.zeros(npatch * Q * maxP), np.zeros(npatch * Q * maxP * 3)
        _check(oracle().or_efs_get_explicit_sparse(self.h, npatch, Q, iptr(qidx), dptr(cjinv), eoffset, C.byref(P), iptr(eidx), dptr(interp), dptr(deriv)), "or_efs_get_explicit_sparse")
        P = P.value
        return P, eidx[:npatch * P].reshape(npatch, P), interp[:npatch * Q * P].reshape(npatch, Q, P), deriv[:npatch * Q * P * 3].reshape(npatch, Q, P, 3)


class RefEFS:
    """The reference's own dEFS/dRule pair, through oracle/ref_driver.c."""

    def __init__(self, bdeg, rorder=None, method=FAST, family=GAUSS):
        lib = ref()
        if family != getattr(RefEFS, "_family", GAUSS):
            lib.RefInit(family)
            RefEFS._family = family
        bdeg = [bdeg] * 3 if np.isscalar(bdeg) else list(bdeg)
        if rorder is None:
            rorder = [2 * b for b in bdeg]
        rorder = [rorder] * 3 if np.isscalar(rorder) else list(rorder)
        self.efs, self.rule_h = C.c_void_p(), C.c_void_p()
        _check(lib.RefGetEFS((C.c_int * 3)(*bdeg), (C.c_int * 3)(*rorder), method, C.byref(self.efs), C.byref(self.rule_h)), "RefGetEFS")
        P = (C.c_int * 3)()
        Q = (C.c_int * 3)()
        lib.RefEFSGetPQ(self.efs, P, Q)
        self.P, self.Q = list(P), list(Q)
        self.np, self.nq = int(np.prod(self.P)), int(np.prod(self.Q))

    def tables(self, d):
        P, Q = self.P[d], self.Q[d]
        interp, deriv, node, weight = np.zeros((Q, P)), np.zeros((Q, P)), np.zeros(P), np.zeros(P)
        ref().RefEFSGetTables(self.efs, d, dptr(interp), dptr(deriv), dptr(node), dptr(weight))
        return interp, deriv, node, weight

    def rule(self, d):
        c, w = np.zeros(self.Q[d]), np.zeros(self.Q[d])
        _check(ref().RefRuleGetTensorNodeWeight(self.rule_h, d, dptr(c), dptr(w)), "RefRuleGetTensorNodeWeight")
        return c, w

    def patches(self):
        n, ps = C.c_int(), C.c_int()
        ind, wt = C.POINTER(C.c_int)(), C.POINTER(C.c_double)()
        _check(ref().RefRuleGetPatches(self.rule_h, C.byref(n), C.byref(ps), C.byref(ind), C.byref(wt)), "RefRuleGetPatches")
        N = n.value * ps.value
        return n.value, ps.value, np.ctypeslib.as_array(ind, (N,)).copy(), np.ctypeslib.as_array(wt, (N,)).copy()

    def apply(self, jinv, D, x, amode, imode=INSERT_VALUES, out=None, nthreads=1):
        x = np.ascontiguousarray(x, dtype=np.float64)
        nelem = x.shape[0]
        shape = {APPLY_INTERP: (nelem, self.nq, D), APPLY_GRAD: (nelem, self.nq, D, 3),
                 APPLY_INTERP_TRANSPOSE: (nelem, self.np, D), APPLY_GRAD_TRANSPOSE: (nelem, self.np, D)}[amode]
        if out is None:
            out = np.zeros(shape)
        _check(ref().RefEFSApplyBatch(self.efs, nelem, dptr(jinv), D, dptr(x), dptr(out), amode, imode, nthreads), "RefEFSApplyBatch")
        return out

    def compute_physical(self, jac):
        jac = np.ascontiguousarray(jac)
        nelem = jac.shape[0]
        jinv, jw = np.zeros((nelem, self.nq, 9)), np.zeros((nelem, self.nq))
        _check(ref().RefRuleComputePhysicalBatch(self.rule_h, nelem, dptr(jac), dptr(jinv), dptr(jw)), "RefRuleComputePhysicalBatch")
        return jinv, jw

    def explicit_sparse(self, cjinv=None, eoffset=0):
        npatch, Q, qidx, _ = self.patches()
        maxP = self.np
        P = C.c_int()
        eidx = np.zeros(npatch * maxP, dtype=np.int32)
        interp, deriv = np.zeros(npatch * Q * maxP), np.zeros(npatch * Q * maxP * 3)
        _check(ref().RefEFSGetExplicitSparse(self.efs, npatch, Q, iptr(qidx), dptr(cjinv), eoffset, C.byref(P), iptr(eidx), dptr(interp), dptr(deriv)), "RefEFSGetExplicitSparse")
        P = P.value
        return P, eidx[:npatch * P].reshape(npatch, P), interp[:npatch * Q * P].reshape(npatch, Q, P), deriv[:npatch * Q * P * 3].reshape(npatch, Q, P, 3)


# ---- FS-level oracle drivers (operate on dohp_b200.box.BoxSpace index arrays) ---------------------------------------------

def or_space(space, bs=None):
    s = OrSpace(space.nelem, space.bs if bs is None else bs, space.n, space.nc, space.ngh, iptr(space.e2l))
    s._keep = space.e2l
    return s


def norms_algebraic_scaled(r, bs=1):
    """dNormsAlgebraicScaled (src/util/util.c:154-168): l^p norms scaled to mimic L^p, clipped at 1e-14 -> 1e-99"""
    n = max(len(r) // bs, 1)
    v = [np.abs(r).sum() / n, np.linalg.norm(r) / np.sqrt(n), np.abs(r).max() if len(r) else 0.0]
    return [x if x >= 1e-14 else 1e-99 for x in v]


def global_to_local(space, g, dcache=None, hmode=HOMOGENEOUS, bs=None):
    bs = space.bs if bs is None else bs
    s = or_space(space, bs)
    loc = np.zeros(space.nloc * bs)
    _check(oracle().or_fs_global_to_local(C.byref(s), dptr(np.ascontiguousarray(g, dtype=np.float64)), dptr(dcache), hmode, dptr(loc)), "or_fs_global_to_local")
    return loc


class OracleEllip:
    """The reference's ellip example restated: residual, matrix-free Jacobian and Q1 assembly (ellip.c:407-536)."""

    def __init__(self, space, p=2.0, eps=1.0, lam=0.0, exact=0, a=1.0, b=1.0, c=1.0, jac_qmethod=SPARSE):
        self.space = space
        self.prm = OrEllipParam(p, eps, lam, exact, a, b, c)
        deg = space.p
        self.efs = OracleEFS(deg)                                   # FAST rule of order 2*deg
        self.cefs = OracleEFS(1, rorder=2 * deg)                    # coordinate space on the same rule
        self.efs_s = OracleEFS(deg, method=jac_qmethod)
        self.cefs_s = OracleEFS(1, rorder=2 * deg, method=jac_qmethod)
        self.stash = np.zeros((space.nelem, self.efs.nq, 5))
        self.s = or_space(space, 1)
        self.coords = np.ascontiguousarray(space.coords)

    def exact_solution(self, xyz):
        u, du = np.zeros(len(xyz)), np.zeros((len(xyz), 3))
        lib = oracle()
        for i, x in enumerate(np.ascontiguousarray(xyz)):
            uu, dd = np.zeros(1), np.zeros(3)
            lib.or_ellip_exact_solution(C.byref(self.prm), dptr(x), dptr(uu), dptr(dd))
            u[i], du[i] = uu[0], dd
        return u, du

    def forcing(self, xq):
        """The example's rhs callback evaluated at quadrature points xq[...,3] (ellip.c:361)."""
        xq = np.ascontiguousarray(xq, dtype=np.float64).reshape(-1, 3)
        f = np.zeros(len(xq))
        lib = oracle()
        ff = np.zeros(1)
        for i, x in enumerate(xq):
            lib.or_ellip_exact_forcing(C.byref(self.prm), dptr(x), dptr(ff))
            f[i] = ff[0]
        return f

    def dirichlet_cache(self):
        """Exact solution at the Dirichlet nodes (ellip.c:717-723)."""
        sp = self.space
        return self.exact_solution(sp.node_xyz[sp.n:sp.nc])[0]

    def function_local(self, xloc):
        yloc = np.zeros(self.space.nloc)
        _check(oracle().or_ellip_function(self.cefs.h, self.efs.h, C.byref(self.s), dptr(self.coords), C.byref(self.prm),
                                          dptr(xloc), dptr(yloc), dptr(self.stash)), "or_ellip_function")
        return yloc

    def function(self, gx, dcache=None):
        """F(x): inhomogeneous Dirichlet data, result restricted to the explicit dofs (single rank)."""
        if dcache is None:
            dcache = self.dirichlet_cache()
        xloc = global_to_local(self.space, gx, dcache, INHOMOGENEOUS, 1)
        return self.function_local(xloc)[:self.space.n]

    def mult_local(self, xloc):
        yloc = np.zeros(self.space.nloc)
        _check(oracle().or_ellip_mult(self.cefs.h, self.efs.h, C.byref(self.s), dptr(self.coords), C.byref(self.prm),
                                      dptr(self.stash), dptr(xloc), dptr(yloc)), "or_ellip_mult")
        return yloc

    def mult(self, gx):
        xloc = global_to_local(self.space, gx, None, HOMOGENEOUS, 1)
        return self.mult_local(xloc)[:self.space.n]

    def error_norms(self, gx, dcache=None):
        """EllipErrorNorms (ellip.c:538-564): the 'Pointwise solution error' and 'Pointwise gradient error' triples of the example's output."""
        if dcache is None:
            dcache = self.dirichlet_cache()
        xloc = global_to_local(self.space, gx, dcache, INHOMOGENEOUS, 1)
        en, gn = np.zeros(3), np.zeros(3)
        _check(oracle().or_ellip_error_norms(self.cefs.h, self.efs.h, C.byref(self.s), dptr(self.coords), C.byref(self.prm), dptr(xloc), dptr(en), dptr(gn)), "or_ellip_error_norms")
        return en, gn

    def nodal_solution(self):
        """EllipGetNodalSolutionVector (ellip.c:566-592): the exact solution at the explicit nodes"""
        return self.exact_solution(self.space.node_xyz[:self.space.n])[0]

    def jacobian_q1(self, gx, rowptr, colidx, dcache=None):
        if dcache is None:
            dcache = self.dirichlet_cache()
        xloc = global_to_local(self.space, gx, dcache, INHOMOGENEOUS, 1)
        vals = np.zeros(len(colidx))
        _check(oracle().or_ellip_jacobian_q1(self.cefs_s.h, self.efs_s.h, C.byref(self.s), dptr(self.coords), C.byref(self.prm),
                                             dptr(xloc), iptr(self.space.l2g), iptr(rowptr), iptr(colidx), dptr(vals)), "or_ellip_jacobian_q1")
        return vals


class OrProjParam(C.Structure):
    _fields_ = [("exact", C.c_int), ("a", C.c_double), ("b", C.c_double), ("c", C.c_double)]


# the two-hex mesh of the projection test (createHexMesh, ex1.c:53-102): parametric (default) or -affine vertices, iMesh hex order
EX1_VERTS_PARAMETRIC = np.array([0.1, 0.2, 0.3, 1.3, 0.2, 0.1, 1.1, 1.2, 0.3, 0.3, 1.2, 0.1, 0.1, 0.2, 1.3, 1.3, 0.2, 1.1, 1.1, 1.2, 1.3, 0.3, 1.2, 1.1,
                                 2.1, 0.2, 0.3, 2.3, 1.2, 0.1, 2.1, 1.2, 1.3, 2.3, 0.2, 1.1]).reshape(12, 3)
EX1_VERTS_AFFINE = np.array([0, 0, 0, 1, 0, 0, 1, 1, 0, 0, 1, 0, 0, 0, 1, 1, 0, 1, 1, 1, 1, 0, 1, 1, 2, 0, 0, 2, 1, 0, 2, 1, 1, 2, 0, 1], dtype=np.float64).reshape(12, 3)
EX1_HEXES = np.array([[0, 1, 2, 3, 4, 5, 6, 7], [1, 2, 6, 5, 8, 9, 10, 11]])
# the faces and edges the test creates explicitly, in its creation order (fconn / econn, ex1.c:66-70 = cunit.c:83-86)
EX1_FACES = np.array([0, 1, 2, 3, 1, 2, 6, 5, 2, 3, 7, 6, 0, 3, 7, 4, 0, 1, 5, 4, 5, 6, 7, 4, 8, 11, 5, 1, 2, 1, 8, 9, 8, 9, 10, 11, 11, 10, 6, 5, 9, 10, 6, 2]).reshape(11, 4)
EX1_EDGES = np.array([0, 1, 4, 5, 7, 6, 3, 2, 1, 8, 5, 11, 6, 10, 2, 9, 0, 3, 3, 7, 7, 4, 4, 0, 1, 2, 2, 6, 6, 5, 5, 1, 8, 11, 11, 10, 10, 9, 9, 8]).reshape(20, 2)


class OracleProj:
    """The reference's projection test restated (src/fs/tests/ex1.c): residual (u - f) jw, the mass operator, the Q1 mass matrix and
    the pointwise residual norms the test prints."""

    def __init__(self, space, exact=0, frequency=(1.0, 1.0, 1.0), proj_qmethod=FAST, jac_qmethod=SPARSE):
        self.space = space
        self.prm = OrProjParam(exact, *frequency)
        deg = space.p
        self.efs = OracleEFS(deg, method=proj_qmethod)
        self.cefs = OracleEFS(1, rorder=2 * deg, method=proj_qmethod)
        self.efs_s = OracleEFS(deg, method=jac_qmethod)
        self.cefs_s = OracleEFS(1, rorder=2 * deg, method=jac_qmethod)
        self.s = or_space(space, 1)
        self.coords = np.ascontiguousarray(space.coords)

    def target(self, xq):
        """exact.function at points xq[...,3] (ex1.c:33-51): what the residual subtracts at the quadrature points."""
        xq = np.ascontiguousarray(xq, dtype=np.float64).reshape(-1, 3)
        f, ff, dd = np.zeros(len(xq)), np.zeros(1), np.zeros(3)
        lib = oracle()
        for i, x in enumerate(xq):
            lib.or_proj_exact(C.byref(self.prm), dptr(x), dptr(ff), dptr(dd))
            f[i] = ff[0]
        return f

    def _local(self, gx):
        return global_to_local(self.space, gx, np.zeros(self.space.nc - self.space.n), INHOMOGENEOUS, 1)

    def function(self, gx):
        yloc = np.zeros(self.space.nloc)
        _check(oracle().or_proj_function(self.cefs.h, self.efs.h, C.byref(self.s), dptr(self.coords), C.byref(self.prm), dptr(self._local(gx)), dptr(yloc)), "or_proj_function")
        return yloc[:self.space.n]

    def mult(self, gx):
        yloc = np.zeros(self.space.nloc)
        _check(oracle().or_proj_mult(self.cefs.h, self.efs.h, C.byref(self.s), dptr(self.coords), dptr(self._local(gx)), dptr(yloc)), "or_proj_mult")
        return yloc[:self.space.n]

    def residual_norms(self, gx):
        res, gres = np.zeros(3), np.zeros(3)
        _check(oracle().or_proj_residual_norms(self.cefs.h, self.efs.h, C.byref(self.s), dptr(self.coords), C.byref(self.prm), dptr(self._local(gx)), dptr(res), dptr(gres)),
               "or_proj_residual_norms")
        return res, gres

    def jacobian_q1(self, rowptr, colidx):
        vals = np.zeros(len(colidx))
        _check(oracle().or_proj_jacobian_q1(self.cefs_s.h, self.efs_s.h, C.byref(self.s), dptr(self.coords), iptr(self.space.l2g), iptr(rowptr), iptr(colidx), dptr(vals)),
               "or_proj_jacobian_q1")
        return vals


class RefEllipMult:
    """EllipShellMatMult (ellip.c:442-472) with the REFERENCE's own per-element kernels: oracle/ref_driver.c RefEllipMatMult
    calls the unmodified dEFSApply / dRuleComputePhysical of oracle/_ref around the restated gather / pointwise / scatter
    loop.  `stash` defaults to the linear Poisson store (eta = 1, sqrt_mdeta_Du = 0, lambda e^u = 0: ellip_p = 2, lambda = 0)."""

    def __init__(self, space, stash=None):
        self.space = space
        deg = space.p
        self.efs = RefEFS(deg)
        self.cefs = RefEFS(1, rorder=2 * deg)
        if stash is None:
            stash = np.zeros((space.nelem, self.efs.nq, 5))
            stash[:, :, 0] = 1.0
        self.stash = np.ascontiguousarray(stash, dtype=np.float64)
        self.e2l = np.ascontiguousarray(space.e2l, dtype=np.int32)
        self.coords = np.ascontiguousarray(space.coords, dtype=np.float64)

    def mult_local(self, xloc, nthreads=1):
        yloc = np.zeros(self.space.nloc)
        _check(ref().RefEllipMatMult(self.cefs.efs, self.efs.efs, self.efs.rule_h, self.space.nelem, self.space.nloc, iptr(self.e2l), dptr(self.coords),
                                     dptr(self.stash), dptr(np.ascontiguousarray(xloc, dtype=np.float64)), dptr(yloc), nthreads), "RefEllipMatMult")
        return yloc


class OracleElast:
    """The reference's elast example restated (elast.c:220-407): St-Venant-Kirchhoff residual, matrix-free Jacobian, Q1 assembly."""

    def __init__(self, space, E0=1.0, nu=0.3, exact=0, a=1.0, b=1.0, c=1.0, scale=1.0, jac_qmethod=SPARSE):
        self.space = space
        lam, mu = E0 * nu / ((1 + nu) * (1 - 2 * nu)), E0 / (2 * (1 + nu))            # ElastSetMaterial, elast.c:62-70
        self.lam, self.mu = lam, mu
        self.prm = OrElastParam(lam, mu, exact, a, b, c, scale)
        deg = space.p
        self.efs, self.cefs = OracleEFS(deg), OracleEFS(1, rorder=2 * deg)
        self.efs_s, self.cefs_s = OracleEFS(deg, method=jac_qmethod), OracleEFS(1, rorder=2 * deg, method=jac_qmethod)
        self.stash = np.zeros((space.nelem, self.efs.nq, 9))
        self.s = or_space(space, 3)
        self.coords = np.ascontiguousarray(space.coords)

    def exact_solution(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
        u, du = np.zeros((len(xyz), 3)), np.zeros((len(xyz), 9))
        for i, x in enumerate(xyz):
            oracle().or_elast_exact_solution(C.byref(self.prm), dptr(x), dptr(u[i]), dptr(du[i]))
        return u, du

    def forcing(self, xq):
        xq = np.ascontiguousarray(xq, dtype=np.float64).reshape(-1, 3)
        f = np.zeros((len(xq), 3))
        for i, x in enumerate(xq):
            oracle().or_elast_exact_forcing(C.byref(self.prm), dptr(x), dptr(f[i]))
        return f

    def dirichlet_cache(self):
        sp = self.space
        return np.ascontiguousarray(self.exact_solution(sp.node_xyz[sp.n:sp.nc])[0].reshape(-1))

    def function(self, gx, dcache=None):
        if dcache is None:
            dcache = self.dirichlet_cache()
        xloc = global_to_local(self.space, gx, dcache, INHOMOGENEOUS, 3)
        yloc = np.zeros(self.space.nloc * 3)
        _check(oracle().or_elast_function(self.cefs.h, self.efs.h, C.byref(self.s), dptr(self.coords), C.byref(self.prm), dptr(xloc), dptr(yloc), dptr(self.stash)), "or_elast_function")
        return yloc[:self.space.n * 3]

    def mult(self, gx):
        xloc = global_to_local(self.space, gx, None, HOMOGENEOUS, 3)
        yloc = np.zeros(self.space.nloc * 3)
        _check(oracle().or_elast_mult(self.cefs.h, self.efs.h, C.byref(self.s), dptr(self.coords), C.byref(self.prm), dptr(self.stash), dptr(xloc), dptr(yloc)), "or_elast_mult")
        return yloc[:self.space.n * 3]

    def error_norms(self, gx, dcache=None):
        """ElastErrorNorms, elast.c:409-438"""
        if dcache is None:
            dcache = self.dirichlet_cache()
        xloc = global_to_local(self.space, gx, dcache, INHOMOGENEOUS, 3)
        en, gn = np.zeros(3), np.zeros(3)
        _check(oracle().or_elast_error_norms(self.cefs.h, self.efs.h, C.byref(self.s), dptr(self.coords), C.byref(self.prm), dptr(xloc), dptr(en), dptr(gn)), "or_elast_error_norms")
        return en, gn

    def nodal_solution(self):
        return self.exact_solution(self.space.node_xyz[:self.space.n])[0].reshape(-1)

    def jacobian_q1(self, gx, rowptr, colidx, dcache=None):
        if dcache is None:
            dcache = self.dirichlet_cache()
        xloc = global_to_local(self.space, gx, dcache, INHOMOGENEOUS, 3)
        vals = np.zeros(len(colidx))
        _check(oracle().or_elast_jacobian_q1(self.cefs_s.h, self.efs_s.h, C.byref(self.s), dptr(self.coords), C.byref(self.prm),
                                             dptr(xloc), iptr(self.space.l2g), iptr(rowptr), iptr(colidx), dptr(vals)), "or_elast_jacobian_q1")
        return vals


class OracleStokes:
    """The reference's stokes example restated (stokes.c:447-673): velocity space `su` (degree p, bs 3), pressure space
    `sp` (degree p - codim, bs 1), both tabulated on the velocity rule (stokes.c:292-298)."""

    def __init__(self, su, sp, B=1.0, eps=1.0, p=2.0, nu=0.0, exact=1, a=1.0, b=1.0, c=1.0, scale=1.0):
        self.su, self.sp = su, sp
        self.prm = OrStokesParam(B, eps, p, nu, exact, a, b, c, scale)
        self.efsu, self.efsp = OracleEFS(su.p), OracleEFS(sp.p, rorder=2 * su.p)
        self.cefs = OracleEFS(1, rorder=2 * su.p)
        self.stash = np.zeros((su.nelem, self.efsu.nq, 8))
        self.osu, self.osp = or_space(su, 3), or_space(sp, 1)
        self.coords = np.ascontiguousarray(su.coords)

    def exact_solution(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
        u, p = np.zeros((len(xyz), 3)), np.zeros(len(xyz))
        du, dp, pp = np.zeros(9), np.zeros(3), np.zeros(1)
        for i, x in enumerate(xyz):
            oracle().or_stokes_exact_solution(C.byref(self.prm), dptr(x), dptr(u[i]), dptr(du), dptr(pp), dptr(dp))
            p[i] = pp[0]
        return u, p

    def forcing(self, xq):
        xq = np.ascontiguousarray(xq, dtype=np.float64).reshape(-1, 3)
        f = np.zeros((len(xq), 4))
        fp = np.zeros(1)
        for i, x in enumerate(xq):
            oracle().or_stokes_exact_forcing(C.byref(self.prm), dptr(x), dptr(f[i, :3]), dptr(fp))
            f[i, 3] = fp[0]
        return f

    def dirichlet_cache(self):
        s = self.su
        return np.ascontiguousarray(self.exact_solution(s.node_xyz[s.n:s.nc])[0].reshape(-1))

    def function(self, xu, xp, dcache=None):
        if dcache is None:
            dcache = self.dirichlet_cache()
        xul = global_to_local(self.su, xu, dcache, INHOMOGENEOUS, 3)
        xpl = global_to_local(self.sp, xp, np.zeros(self.sp.nc - self.sp.n), INHOMOGENEOUS, 1)
        yu, yp = np.zeros(self.su.nloc * 3), np.zeros(self.sp.nloc)
        _check(oracle().or_stokes_function(self.cefs.h, self.efsu.h, self.efsp.h, C.byref(self.osu), C.byref(self.osp), dptr(self.coords), C.byref(self.prm),
                                           dptr(xul), dptr(xpl), dptr(yu), dptr(yp), dptr(self.stash)), "or_stokes_function")
        return yu[:self.su.n * 3], yp[:self.sp.n]

    def mult(self, xu, xp, block=0):
        xul = global_to_local(self.su, xu if xu is not None else np.zeros(self.su.n * 3), None, HOMOGENEOUS, 3)
        xpl = global_to_local(self.sp, xp if xp is not None else np.zeros(self.sp.n), None, HOMOGENEOUS, 1)
        yu, yp = np.zeros(self.su.nloc * 3), np.zeros(self.sp.nloc)
        _check(oracle().or_stokes_mult(self.cefs.h, self.efsu.h, self.efsp.h, C.byref(self.osu), C.byref(self.osp), dptr(self.coords), C.byref(self.prm),
                                       dptr(self.stash), block, dptr(xul), dptr(xpl), dptr(yu), dptr(yp)), "or_stokes_mult")
        return yu[:self.su.n * 3], yp[:self.sp.n]

    def error_norms(self, xu, xp, dcache=None, alldirichlet=True):
        """StokesErrorNorms, stokes.c:835-885: (velocity, velocity gradient, pressure) error triples"""
        if dcache is None:
            dcache = self.dirichlet_cache()
        xul = global_to_local(self.su, xu, dcache, INHOMOGENEOUS, 3)
        xpl = global_to_local(self.sp, xp, np.zeros(self.sp.nc - self.sp.n), INHOMOGENEOUS, 1)
        en, gn, pn = np.zeros(3), np.zeros(3), np.zeros(3)
        _check(oracle().or_stokes_error_norms(self.cefs.h, self.efsu.h, self.efsp.h, C.byref(self.osu), C.byref(self.osp), dptr(self.coords), C.byref(self.prm),
                                              dptr(xul), dptr(xpl), int(alldirichlet), dptr(en), dptr(gn), dptr(pn)), "or_stokes_error_norms")
        return en, gn, pn

    def nodal_solution(self):
        """StokesGetSolutionVector, stokes.c:926-944: exact velocity and pressure at the explicit nodes"""
        return self.exact_solution(self.su.node_xyz[:self.su.n])[0].reshape(-1), self.exact_solution(self.sp.node_xyz[:self.sp.n])[1]

    def jacobian_q1_velocity(self, xu, rowptr, colidx, dcache=None):
        """StokesJacobianAssemble_Velocity (stokes.c:675-737): scalar-CSR values of the velocity block Ap and the Mdiag vector."""
        if dcache is None:
            dcache = self.dirichlet_cache()
        efs_s, cefs_s = OracleEFS(self.su.p, method=SPARSE), OracleEFS(1, rorder=2 * self.su.p, method=SPARSE)
        xul = global_to_local(self.su, xu, dcache, INHOMOGENEOUS, 3)
        vals, md = np.zeros(len(colidx)), np.zeros(self.su.nloc * 3)
        _check(oracle().or_stokes_jacobian_q1_velocity(cefs_s.h, efs_s.h, C.byref(self.osu), dptr(self.coords), C.byref(self.prm), dptr(xul),
                                                       iptr(self.su.l2g), iptr(rowptr), iptr(colidx), dptr(vals), dptr(md)), "or_stokes_jacobian_q1_velocity")
        return vals, md[:self.su.n * 3]

    def jacobian_pressure(self, xu, rowptr, colidx, dcache=None):
        """StokesJacobianAssemble_Pressure (stokes.c:739-794): values of D (mass/eta) and Daux (Neumann Laplacian) on the
        element-dense pressure pattern."""
        if dcache is None:
            dcache = self.dirichlet_cache()
        efsu_s, cefs_s = OracleEFS(self.su.p, method=SPARSE), OracleEFS(1, rorder=2 * self.su.p, method=SPARSE)
        efsp_s = OracleEFS(self.sp.p, rorder=2 * self.su.p, method=SPARSE)
        xul = global_to_local(self.su, xu, dcache, INHOMOGENEOUS, 3)
        D, Da = np.zeros(len(colidx)), np.zeros(len(colidx))
        _check(oracle().or_stokes_jacobian_pressure(cefs_s.h, efsu_s.h, efsp_s.h, C.byref(self.osu), C.byref(self.osp), dptr(self.coords), C.byref(self.prm),
                                                    dptr(xul), iptr(self.sp.l2g), iptr(rowptr), iptr(colidx), dptr(D), dptr(Da)), "or_stokes_jacobian_pressure")
        return D, Da


def rule_1d(method, family, order):
    """or_rule_1d: TensorGetRule restated (tensorquad.c:51-118)."""
    size = C.c_int()
    c, w = np.zeros(order + 2), np.zeros(order + 2)
    _check(oracle().or_rule_1d(method, family, order, C.byref(size), dptr(c), dptr(w)), "or_rule_1d")
    return c[:size.value].copy(), w[:size.value].copy()


def basis_1d(P, qcoord):
    """or_basis_1d: TensorBasisCreate restated (tensor.c:22-72)."""
    qcoord = np.ascontiguousarray(qcoord, dtype=np.float64)
    Q = len(qcoord)
    interp, deriv, node, weight = np.zeros((Q, P)), np.zeros((Q, P)), np.zeros(P), np.zeros(P)
    _check(oracle().or_basis_1d(P, Q, dptr(qcoord), dptr(interp), dptr(deriv), dptr(node), dptr(weight)), "or_basis_1d")
    return interp, deriv, node, weight


# ---- Line / Quad dEFS ---------------------------------------------------------------------------------------------
def lowdim_apply(degrees, x, amode, D, rorder=None, method=FAST, family=GAUSS):
    """dEFSApply_Tensor_Line / _Quad (efstopo.c:290-366) with TensorMult_Line / _Quad (efstopo.c:453-516) and TensorRuleNoMapping
    (efstopo.c:518-546) restated in numpy on the oracle's 1-D tables; pinned against the reference's own functions by
    tests/test_oracle_lowdim.py.  x: [nelem][prod P][D] (INTERP, GRAD) or [nelem][prod Q][D] (INTERP_TRANSPOSE)."""
    dim = len(degrees)
    rorder = [2 * d for d in degrees] if rorder is None else rorder
    rules = [rule_1d(method, family, r) for r in rorder]
    tabs = [basis_1d(d + 1, r[0]) for d, r in zip(degrees, rules)]
    B, Dm = [t[0] for t in tabs], [t[1] for t in tabs]
    P, Q = [b.shape[1] for b in B], [b.shape[0] for b in B]
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    if dim == 1:
        if amode == APPLY_INTERP:
            return np.einsum("li,eid->eld", B[0], x.reshape(n, P[0], D))
        if amode == APPLY_INTERP_TRANSPOSE:
            return np.einsum("li,eld->eid", B[0], x.reshape(n, Q[0], D))
        if amode == APPLY_GRAD:
            g = np.zeros((n, Q[0], D, 3))                       # stride 3 on every topology (TensorRuleNoMapping); unused directions: zero here, garbage upstream
            g[..., 0] = np.einsum("li,eid->eld", Dm[0], x.reshape(n, P[0], D))
            return g
    else:
        if amode == APPLY_INTERP:
            return np.einsum("ai,bj,eijd->eabd", B[0], B[1], x.reshape(n, P[0], P[1], D)).reshape(n, Q[0] * Q[1], D)
        if amode == APPLY_INTERP_TRANSPOSE:
            return np.einsum("ai,bj,eabd->eijd", B[0], B[1], x.reshape(n, Q[0], Q[1], D)).reshape(n, P[0] * P[1], D)
        if amode == APPLY_GRAD:
            g0 = np.einsum("ai,bj,eijd->eabd", Dm[0], B[1], x.reshape(n, P[0], P[1], D))
            g1 = np.einsum("ai,bj,eijd->eabd", B[0], Dm[1], x.reshape(n, P[0], P[1], D))
            return np.stack([g0, g1, np.zeros_like(g0)], axis=-1).reshape(n, Q[0] * Q[1], D, 3)
    raise ValueError("invalid dApplyMode %d on a %d-dimensional dEFS (efstopo.c:321-325, 361-363)" % (amode, dim))


def ref_lowdim_apply(degrees, x, amode, D, rorder=None, method=FAST):
    """The reference's own dEFSApply on a dTOPO_LINE / dTOPO_QUAD dEFS (oracle/_ref)."""
    dim = len(degrees)
    rorder = [2 * d for d in degrees] if rorder is None else rorder
    bd, ro = list(degrees) + [0] * (3 - dim), list(rorder) + [0] * (3 - dim)
    efs, rule = C.c_void_p(), C.c_void_p()
    _check(ref().RefGetEFSDim(dim, (C.c_int * 3)(*bd), (C.c_int * 3)(*ro), method, C.byref(efs), C.byref(rule)), "RefGetEFSDim")
    P, Q = (C.c_int * 3)(), (C.c_int * 3)()
    ref().RefEFSGetPQDim(efs, dim, P, Q)
    npts, nq = int(np.prod(list(P))), int(np.prod(list(Q)))
    x = np.ascontiguousarray(x, dtype=np.float64)
    n = x.shape[0]
    out = np.zeros({APPLY_INTERP: (n, nq, D), APPLY_GRAD: (n, nq, D, 3)}.get(amode, (n, npts, D)))
    err = ref().RefEFSApplyBatchDim(efs, dim, n, D, dptr(x), dptr(out), amode, INSERT_VALUES)
    return err, out, list(P)[:dim], list(Q)[:dim]


# the six rotations + translations of dQuadratureGetFacetRules_Tensor_Private (tensorquad.c:321-328)
FACET_JAC = np.array([[[1, 0, 0], [0, 0, -1], [0, 1, 0]], [[0, 0, 1], [1, 0, 0], [0, 1, 0]], [[-1, 0, 0], [0, 0, 1], [0, 1, 0]],
                      [[0, 0, -1], [-1, 0, 0], [0, 1, 0]], [[0, 1, 0], [1, 0, 0], [0, 0, -1]], [[1, 0, 0], [0, 1, 0], [0, 0, 1]]], dtype=np.float64)
FACET_TRANSLATION = np.array([[0, -1, 0], [1, 0, 0], [0, 1, 0], [-1, 0, 0], [0, 0, -1], [0, 0, 1]], dtype=np.float64)


def facet_rule_nodes(facet, coord, weight):
    """dRuleGetNodeWeight_Transformed (tensorquad.c:225-243) on the QUAD rule's tensor nodes: x = jac X_ref + translation."""
    X = np.array([[a, b, 0.0] for a in coord[0] for b in coord[1]])
    return X @ FACET_JAC[facet].T + FACET_TRANSLATION[facet], np.array([wa * wb for wa in weight[0] for wb in weight[1]])
