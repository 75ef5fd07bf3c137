"""ctypes binding of libdohp_b200.so (include/dohp_b200.h) and thin host-side mirrors of the reference's objects.

The names follow the reference: ``Rule`` = dRule, ``EFS`` = dEFS (``apply`` = dEFSApply over a batch of elements),
``Space`` = dFS (E, local layout, ghost scatter), ``Op`` = one example's dRulesetIterator traversal with its stash
(``function`` = XFunction, ``mult`` = XShellMatMult, ``assemble_q1`` = XJacobian).  Buffers may be numpy arrays (host)
or torch CUDA tensors (device); the C library detects which.  Nothing here computes: without the built CUDA library
every call fails loudly (there is no CPU path).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.environ.get("DOHP_B200_LIB", os.path.join(HERE, "libdohp_b200.so"))      # override: kernel-variant experiments

APPLY_INTERP, APPLY_INTERP_TRANSPOSE, APPLY_GRAD, APPLY_GRAD_TRANSPOSE = 40001, 40002, 40003, 40004
INSERT_VALUES, ADD_VALUES = 1, 2
FAST, SPARSE, SELF = 0, 1, 2
GAUSS, LOBATTO = 0, 1
HOMOGENEOUS, INHOMOGENEOUS = 0, 1
OP_ELLIP, OP_ELAST, OP_STOKES, OP_PROJECTION = 1, 2, 3, 4
DENSE_ELLIP_FAST, DENSE_STOKES_D, DENSE_STOKES_DAUX, DENSE_ELAST_FAST = 1, 2, 3, 4
STOKES_J, STOKES_A, STOKES_B, STOKES_Bt = 0, 1, 2, 3
WALLS, TOP, BOTTOM = 1, 2, 4          # dirichlet_mask bits: genblock sets 100, 200, 300

# every symbol include/dohp_b200.h declares (tests/test_abi.py checks the library exports all of them)
SYMBOLS = [
    "dB200GetLastError", "dB200Init", "dB200NcclUniqueId", "dB200InitComm", "dB200Destroy", "dB200SetStream", "dB200Synchronize",
    "dB200NumLaunches", "dB200Rule1D", "dB200Basis1D", "dB200RuleCreate", "dB200RuleDestroy", "dB200EFSCreate", "dB200EFSDestroy",
    "dB200EFSGetSizes", "dB200RuleCreateDim", "dB200EFSCreateDim", "dB200FacetRuleNodes", "dB200EFSApplyBatch", "dB200RuleComputePhysicalBatch", "dB200SpaceCreate", "dB200SpaceDestroy",
    "dB200SpaceSetHalo", "dB200SpaceHaloMode", "dB200SpaceGlobalToLocal", "dB200SpaceLocalToGlobal", "dB200SpaceRotationApply", "dB200OpCreate", "dB200OpDestroy",
    "dB200OpGetQuadraturePoints", "dB200OpSetForcing", "dB200OpSetDirichlet", "dB200OpFunction", "dB200OpMult", "dB200OpMultBlock", "dB200OpMultAsync", "dB200StreamJoin",
    "dB200OpGetStash", "dB200OpQ1Pattern", "dB200OpAssembleQ1", "dB200OpStokesLumpedMass", "dB200OpQ1SetReduced", "dB200OpQ1OwnerPattern", "dB200OpQ1OwnerReduce", "dB200OpDensePattern", "dB200OpAssembleDense", "dB200BoxMeshCreate", "dB200BoxMeshDestroy", "dB200MeasureFP64Peak",
]


class EllipParam(C.Structure):
    _fields_ = [("p", C.c_double), ("epsilon", C.c_double), ("lam", C.c_double)]


class ElastParam(C.Structure):
    _fields_ = [("lam", C.c_double), ("mu", C.c_double)]


class StokesParam(C.Structure):
    _fields_ = [("B", C.c_double), ("eps", C.c_double), ("p", C.c_double), ("nu", C.c_double)]


class BoxSpec(C.Structure):
    _fields_ = [("ne", C.c_int * 3), ("p", C.c_int), ("procs", C.c_int * 3), ("rank", C.c_int), ("dirichlet_mask", C.c_int),
                ("box", C.c_double * 6), ("morph", C.c_double), ("twist", C.c_double), ("flags", C.c_int)]


class BoxMeshC(C.Structure):
    _fields_ = [("nelem", C.c_int), ("n", C.c_int), ("nc", C.c_int), ("ngh", C.c_int), ("rstart", C.c_longlong), ("nglobal", C.c_longlong),
                ("e2l", C.POINTER(C.c_int)), ("l2g", C.POINTER(C.c_int)), ("coords", C.POINTER(C.c_double)), ("node_xyz", C.POINTER(C.c_double)),
                ("nneigh", C.c_int), ("neigh_rank", C.POINTER(C.c_int)), ("sendcount", C.POINTER(C.c_int)), ("recvcount", C.POINTER(C.c_int)),
                ("sendidx", C.POINTER(C.POINTER(C.c_int))), ("recvidx", C.POINTER(C.POINTER(C.c_int)))]


_lib = None


def build(verbose=False):
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    import subprocess
    r = subprocess.run(["make", "-j", str(os.cpu_count() or 4), "-C", os.path.join(HERE, "csrc")], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libdohp_b200.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    if verbose:
        print(r.stdout[-2000:])
    return LIBPATH


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIBPATH):
            raise RuntimeError("%s is missing: run `make -C dohp_b200/csrc` (or __graft_entry__.build()); there is no CPU fallback" % LIBPATH)
        _lib = C.CDLL(LIBPATH)
        _lib.dB200GetLastError.restype = C.c_char_p
    return _lib


class DohpError(RuntimeError):
    pass


def chk(err):
    if err:
        raise DohpError("dohp_b200 error %d: %s" % (err, load().dB200GetLastError().decode()))


def ptr(a):
    """Raw pointer of a numpy array (host) or a torch tensor (host or device); None -> NULL."""
    if a is None:
        return C.c_void_p(0)
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        assert a.dtype in (np.float64, np.int32), "the C ABI takes double and int (32-bit) arrays only, got %s" % a.dtype
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        assert a.is_contiguous()
        assert str(a.dtype) in ("torch.float64", "torch.int32"), "the C ABI takes double and int (32-bit) arrays only, got %s" % a.dtype
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def rule_1d(method, family, order):
    size = C.c_int()
    coord, weight = np.zeros(order + 2), np.zeros(order + 2)
    chk(load().dB200Rule1D(method, family, order, C.byref(size), ptr(coord), ptr(weight)))
    return coord[:size.value].copy(), weight[:size.value].copy()


def basis_1d(P, qcoord):
    Q = len(qcoord)
    interp, deriv, node, weight = np.zeros((Q, P)), np.zeros((Q, P)), np.zeros(P), np.zeros(P)
    chk(load().dB200Basis1D(P, Q, ptr(_f64(qcoord)), ptr(interp), ptr(deriv), ptr(node), ptr(weight)))
    return interp, deriv, node, weight


class Context:
    def __init__(self, device=0, rank=0, nranks=1, nccl_id=None, stream=None):
        self.h = C.c_void_p()
        if nranks > 1:
            chk(load().dB200InitComm(device, rank, nranks, C.c_char_p(nccl_id), C.byref(self.h)))
        else:
            chk(load().dB200Init(device, C.byref(self.h)))
        self.rank, self.nranks = rank, nranks
        if stream is not None:
            chk(load().dB200SetStream(self.h, C.c_void_p(stream)))

    @staticmethod
    def nccl_unique_id():
        buf = C.create_string_buffer(128)
        chk(load().dB200NcclUniqueId(buf))
        return buf.raw

    def synchronize(self):
        chk(load().dB200Synchronize(self.h))

    def stream_join(self):
        chk(load().dB200StreamJoin(self.h))

    def num_launches(self):
        n = C.c_longlong()
        chk(load().dB200NumLaunches(self.h, C.byref(n)))
        return n.value

    def fp64_peak_tflops(self):
        t = C.c_double()
        chk(load().dB200MeasureFP64Peak(self.h, C.byref(t)))
        return t.value

    def close(self):
        if self.h:
            chk(load().dB200Destroy(C.byref(self.h)))


class Rule:
    """dRule: tensor-product rule on a hex; `coord` / `weight` are one 1-D rule (isotropic) or a list of three (anisotropic)."""

    def __init__(self, ctx, coord, weight):
        aniso = isinstance(coord, (list, tuple)) and len(coord) == 3 and not np.isscalar(coord[0])
        self.coords = [_f64(c) for c in (coord if aniso else [coord] * 3)]
        self.weights = [_f64(w) for w in (weight if aniso else [weight] * 3)]
        self.ctx, self.coord, self.weight = ctx, self.coords[0], self.weights[0]
        self.Q3 = [len(c) for c in self.coords]
        self.Q, self.nq = self.Q3[0], int(np.prod(self.Q3))
        cp = (C.c_void_p * 3)(*[c.ctypes.data for c in self.coords])
        wp = (C.c_void_p * 3)(*[w.ctypes.data for w in self.weights])
        self.h = C.c_void_p()
        chk(load().dB200RuleCreate(ctx.h, (C.c_int * 3)(*self.Q3), cp, wp, C.byref(self.h)))

    @classmethod
    def preferred(cls, ctx, degree, method=FAST, family=GAUSS):
        """dFSGetPreferredQuadratureRuleSet: order = 2*degree per direction (src/fs/interface/fsruleset.c:32-37)."""
        return cls(ctx, *rule_1d(method, family, 2 * degree))

    def compute_physical(self, jac, jinv=None, jw=None):
        jac = _f64(jac)
        nelem = jac.size // (self.nq * 9)
        jinv = np.empty((nelem, self.nq, 9)) if jinv is None else jinv
        jw = np.empty((nelem, self.nq)) if jw is None else jw
        chk(load().dB200RuleComputePhysicalBatch(self.h, nelem, ptr(jac), ptr(jinv), ptr(jw)))
        return jinv, jw


class EFS:
    """dEFS: GLL nodal basis of `degree` (an int, or three for an anisotropic element) tabulated on `rule` (TensorBasisCreate arithmetic
    done by dB200Basis1D unless the caller brings its own tables, e.g. the reference's: one table or a list of three)."""

    def __init__(self, ctx, rule, degree, interp=None, deriv=None):
        degs = [degree] * 3 if np.isscalar(degree) else list(degree)
        self.ctx, self.rule, self.P3 = ctx, rule, [d + 1 for d in degs]
        self.P, self.np = self.P3[0], int(np.prod(self.P3))
        if interp is None:
            tabs = [basis_1d(self.P3[d], rule.coords[d]) for d in range(3)]
            self.interps, self.derivs = [t[0] for t in tabs], [t[1] for t in tabs]
            self.node, self.nweight = tabs[0][2], tabs[0][3]
        else:
            aniso = isinstance(interp, (list, tuple))
            self.interps = [_f64(a) for a in (interp if aniso else [interp] * 3)]
            self.derivs = [_f64(a) for a in (deriv if aniso else [deriv] * 3)]
        self.interps, self.derivs = [_f64(a) for a in self.interps], [_f64(a) for a in self.derivs]
        self.interp, self.deriv = self.interps[0], self.derivs[0]
        ip = (C.c_void_p * 3)(*[a.ctypes.data for a in self.interps])
        dp = (C.c_void_p * 3)(*[a.ctypes.data for a in self.derivs])
        self.h = C.c_void_p()
        chk(load().dB200EFSCreate(ctx.h, rule.h, (C.c_int * 3)(*self.P3), ip, dp, C.byref(self.h)))

    def apply(self, jinv, D, x, amode, imode=INSERT_VALUES, out=None, nelem=None):
        """dEFSApply over nelem elements; x/out follow the reference layouts."""
        P3, Q3 = self.np, self.rule.nq
        fwd = amode in (APPLY_INTERP, APPLY_GRAD)
        g = 3 if amode in (APPLY_GRAD, APPLY_GRAD_TRANSPOSE) else 1
        if isinstance(x, np.ndarray):
            x = _f64(x)
        size = x.size if isinstance(x, np.ndarray) else x.numel()
        if nelem is None:
            nelem = size // (D * (P3 if fwd else Q3 * g))
        if out is None:
            out = np.zeros((nelem, Q3, D, 3) if amode == APPLY_GRAD else (nelem, Q3, D) if amode == APPLY_INTERP else (nelem, P3, D))
        if isinstance(jinv, np.ndarray):
            jinv = _f64(jinv)
        chk(load().dB200EFSApplyBatch(self.h, nelem, ptr(jinv), D, ptr(x), ptr(out), amode, imode))
        return out


class RuleDim:
    """dRule on dTOPO_LINE (dim 1) / dTOPO_QUAD (dim 2): `dim` 1-D rules."""

    def __init__(self, ctx, coord, weight):
        self.ctx, self.dim = ctx, len(coord)
        self.coords, self.weights = [_f64(c) for c in coord], [_f64(w) for w in weight]
        self.Q = [len(c) for c in self.coords]
        self.nq = int(np.prod(self.Q))
        cp = (C.c_void_p * self.dim)(*[a.ctypes.data for a in self.coords])
        wp = (C.c_void_p * self.dim)(*[a.ctypes.data for a in self.weights])
        self.h = C.c_void_p()
        chk(load().dB200RuleCreateDim(ctx.h, self.dim, (C.c_int * self.dim)(*self.Q), cp, wp, C.byref(self.h)))

    @classmethod
    def preferred(cls, ctx, degrees, method=FAST, family=GAUSS):
        r = [rule_1d(method, family, 2 * d) for d in degrees]
        return cls(ctx, [a[0] for a in r], [a[1] for a in r])


class EFSDim:
    """dEFS on a Line / Quad (dEFSApply_Tensor_Line / _Quad, efstopo.c:290-366): INTERP, INTERP_TRANSPOSE and the reference-space GRAD."""

    def __init__(self, ctx, rule, degrees):
        self.ctx, self.rule, self.dim = ctx, rule, rule.dim
        self.Ps = [d + 1 for d in degrees]
        tabs = [basis_1d(self.Ps[d], rule.coords[d]) for d in range(self.dim)]
        self.interps, self.derivs = [_f64(t[0]) for t in tabs], [_f64(t[1]) for t in tabs]
        self.np = int(np.prod(self.Ps))
        ip = (C.c_void_p * self.dim)(*[a.ctypes.data for a in self.interps])
        dp = (C.c_void_p * self.dim)(*[a.ctypes.data for a in self.derivs])
        self.h = C.c_void_p()
        chk(load().dB200EFSCreateDim(ctx.h, rule.h, self.dim, (C.c_int * self.dim)(*self.Ps), ip, dp, C.byref(self.h)))

    def apply(self, D, x, amode, imode=INSERT_VALUES, out=None, jinv=None):
        x = _f64(x)
        nelem = x.shape[0]
        if out is None:
            out = np.zeros({APPLY_INTERP: (nelem, self.rule.nq, D), APPLY_GRAD: (nelem, self.rule.nq, D, 3)}.get(amode, (nelem, self.np, D)))
        chk(load().dB200EFSApplyBatch(self.h, nelem, ptr(jinv), D, ptr(x), ptr(out), amode, imode))
        return out


def facet_rule_nodes(facet, coord, weight):
    """dQuadratureGetFacetRules: points / weights of a QUAD rule on hex facet `facet` in the hex's reference coordinates."""
    c, w = [_f64(a) for a in coord], [_f64(a) for a in weight]
    Q = [len(a) for a in c]
    nodes, wts = np.zeros((Q[0] * Q[1], 3)), np.zeros(Q[0] * Q[1])
    chk(load().dB200FacetRuleNodes(facet, (C.c_int * 2)(*Q), (C.c_void_p * 2)(*[a.ctypes.data for a in c]), (C.c_void_p * 2)(*[a.ctypes.data for a in w]), ptr(nodes), ptr(wts)))
    return nodes, wts


class Space:
    """dFS as the kernels see it."""

    def __init__(self, ctx, nelem, P, bs, n, nc, ngh, e2l, halos=None):
        self.ctx, self.nelem, self.P, self.bs, self.n, self.nc, self.ngh = ctx, nelem, P, bs, n, nc, ngh
        self.h = C.c_void_p()
        e2l = _i32(e2l)
        assert e2l.size == nelem * P ** 3
        chk(load().dB200SpaceCreate(ctx.h, nelem, P, bs, n, nc, ngh, ptr(e2l), C.byref(self.h)))
        if halos:
            nn = len(halos)
            ranks = (C.c_int * nn)(*[h[0] for h in halos])
            send = [_i32(h[1]) for h in halos]
            recv = [_i32(h[2]) for h in halos]
            sc = (C.c_int * nn)(*[len(s) for s in send])
            rc = (C.c_int * nn)(*[len(r) for r in recv])
            sp = (C.c_void_p * nn)(*[s.ctypes.data for s in send])
            rp = (C.c_void_p * nn)(*[r.ctypes.data for r in recv])
            chk(load().dB200SpaceSetHalo(self.h, nn, ranks, sc, sp, rc, rp))

    @classmethod
    def from_mesh(cls, ctx, mesh, bs=1):
        return cls(ctx, mesh.nelem, mesh.p + 1, bs, mesh.n, mesh.nc, mesh.ngh, mesh.e2l, halos=mesh.halos)

    def global_to_local(self, g, dirichlet=None, hmode=HOMOGENEOUS):
        local = np.empty((self.nc + self.ngh) * self.bs)
        chk(load().dB200SpaceGlobalToLocal(self.h, ptr(_f64(g)), ptr(None if dirichlet is None else _f64(dirichlet)), hmode, ptr(local)))
        return local

    def halo_mode(self):
        """'none', 'nccl' (grouped send/recv) or 'p2p' (direct stores into the neighbours' memory over NVLink)."""
        m = C.c_int()
        chk(load().dB200SpaceHaloMode(self.h, C.byref(m)))
        return ("none", "nccl", "p2p")[m.value]

    def rotation_apply(self, idx, rmat, ns, local, strong=None, reverse=False, hmode=HOMOGENEOUS):
        """dFSRotationApplyLocal (fsrot.c): in place on the local vector."""
        idx, ns, rmat = _i32(idx), _i32(ns), _f64(rmat)
        chk(load().dB200SpaceRotationApply(self.h, len(idx), ptr(idx), ptr(rmat), ptr(ns), ptr(None if strong is None else _f64(strong)), ptr(local), int(reverse), hmode))
        return local

    def local_to_global(self, local, g=None, imode=INSERT_VALUES):
        g = np.zeros(self.n * self.bs) if g is None else g
        chk(load().dB200SpaceLocalToGlobal(self.h, ptr(_f64(local)), ptr(g), imode))
        return g


class Op:
    """One example's traversal: rule + spaces + dEFS + expanded geometry + pointwise physics + stash."""

    def __init__(self, ctx, kind, param, rule, spaces, efs, coords):
        self.ctx, self.kind, self.param, self.rule, self.spaces, self.efs = ctx, kind, param, rule, list(spaces), list(efs)
        ns = len(self.spaces)
        sp = (C.c_void_p * ns)(*[s.h for s in self.spaces])
        ep = (C.c_void_p * ns)(*[e.h for e in self.efs])
        self.coords = _f64(coords) if isinstance(coords, np.ndarray) else coords
        self.h = C.c_void_p()
        chk(load().dB200OpCreate(ctx.h, kind, None if param is None else C.byref(param), rule.h, ns, sp, ep, ptr(self.coords), C.byref(self.h)))
        self.nelem, self.Q = self.spaces[0].nelem, rule.Q

    def _vecs(self, xs, ys):
        ns = len(self.spaces)
        xs = list(xs) if isinstance(xs, (list, tuple)) else [xs]
        if ys is None:
            ys = [np.zeros(s.n * s.bs) for s in self.spaces]
        ys = list(ys) if isinstance(ys, (list, tuple)) else [ys]
        xs = [_f64(x) if isinstance(x, np.ndarray) else x for x in xs]
        xp = (C.c_void_p * ns)(*[ptr(x).value or 0 for x in xs])
        yp = (C.c_void_p * ns)(*[ptr(y).value or 0 for y in ys])
        return xs, ys, xp, yp

    def quadrature_points(self):
        xq = np.empty((self.nelem, self.Q ** 3, 3))
        chk(load().dB200OpGetQuadraturePoints(self.h, ptr(xq)))
        return xq

    def set_forcing(self, f):
        chk(load().dB200OpSetForcing(self.h, ptr(None if f is None else _f64(f))))

    def set_dirichlet(self, ifield, vals):
        chk(load().dB200OpSetDirichlet(self.h, ifield, ptr(_f64(vals))))

    def function(self, xs, ys=None):
        xs, ys, xp, yp = self._vecs(xs, ys)
        chk(load().dB200OpFunction(self.h, xp, yp))
        return ys[0] if len(ys) == 1 else ys

    def mult(self, xs, ys=None, block=None):
        xs, ys, xp, yp = self._vecs(xs, ys)
        if block is None:
            chk(load().dB200OpMult(self.h, xp, yp))
        else:
            chk(load().dB200OpMultBlock(self.h, block, xp, yp))
        return ys[0] if len(ys) == 1 else ys

    def mult_async(self, xs, ys):
        """dB200OpMultAsync: host (pinned) vectors, returns immediately; Context.synchronize() / stream_join() completes it."""
        xs, ys, xp, yp = self._vecs(xs, ys)
        chk(load().dB200OpMultAsync(self.h, xp, yp))
        return ys[0] if len(ys) == 1 else ys

    def stash(self):
        ns = C.c_int()
        chk(load().dB200OpGetStash(self.h, C.byref(ns), ptr(None)))
        st = np.empty((self.nelem, self.Q ** 3, ns.value))
        chk(load().dB200OpGetStash(self.h, C.byref(ns), ptr(st)))
        return st

    def q1_pattern(self, l2g):
        l2g = _i32(l2g)
        nrows, nnz = C.c_longlong(), C.c_longlong()
        chk(load().dB200OpQ1Pattern(self.h, ptr(l2g), C.byref(nrows), C.byref(nnz), ptr(None), ptr(None)))
        rowptr, colidx = np.empty(nrows.value + 1, dtype=np.int32), np.empty(nnz.value, dtype=np.int32)
        chk(load().dB200OpQ1Pattern(self.h, ptr(l2g), C.byref(nrows), C.byref(nnz), ptr(rowptr), ptr(colidx)))
        self._q1_nnzb = nnz.value
        return rowptr, colidx

    def assemble_q1(self, xs, l2g, rowptr, colidx, vals=None):
        """Values of the block CSR (bs*bs per block, row-major = BAIJ; AIJ for bs=1) at linearisation state xs."""
        xs, _, xp, _ = self._vecs(xs, None)
        bs = self.spaces[0].bs
        if vals is None:
            nnzb = len(colidx) if colidx is not None else self._q1_nnzb
            vals = np.empty(nnzb * bs * bs)
        rp, ci = (None, None) if rowptr is None else (rowptr, colidx) if hasattr(rowptr, 'data_ptr') else (_i32(rowptr), _i32(colidx))       # None: the pattern cached by q1_pattern
        chk(load().dB200OpAssembleQ1(self.h, xp, ptr(l2g if hasattr(l2g, 'data_ptr') else _i32(l2g)), ptr(rp), ptr(ci), ptr(vals)))
        return vals

    def q1_set_reduced(self, flag):
        chk(load().dB200OpQ1SetReduced(self.h, int(flag)))

    def q1_owner_pattern(self, rowptr, colidx):
        """Rows of the global matrix this rank owns after MatAssemblyEnd (collective): (orowptr, ocolidx), global columns."""
        onnz = C.c_longlong()
        rp, ci = _i32(rowptr), _i32(colidx)
        chk(load().dB200OpQ1OwnerPattern(self.h, ptr(rp), ptr(ci), C.byref(onnz), ptr(None), ptr(None)))
        orowptr, ocolidx = np.empty(self.spaces[0].n + 1, dtype=np.int32), np.empty(onnz.value, dtype=np.int32)
        chk(load().dB200OpQ1OwnerPattern(self.h, ptr(rp), ptr(ci), C.byref(onnz), ptr(orowptr), ptr(ocolidx)))
        self._q1_onnz = onnz.value
        return orowptr, ocolidx

    def q1_owner_reduce(self, vals, ovals=None):
        """Owner-side sum of the ranks' off-process row contributions (collective): values of the owned rows."""
        bs = self.spaces[0].bs
        ovals = np.empty(self._q1_onnz * bs * bs) if ovals is None else ovals
        chk(load().dB200OpQ1OwnerReduce(self.h, ptr(vals), ptr(ovals)))
        return ovals

    def dense_pattern(self, ifield, l2g):
        """Element-dense CSR pattern of field `ifield` (dEFSGetExplicitSparse_Basic path, jacobi.c:611-643)."""
        l2g = _i32(l2g)
        nrows, nnz = C.c_longlong(), C.c_longlong()
        chk(load().dB200OpDensePattern(self.h, ifield, ptr(l2g), C.byref(nrows), C.byref(nnz), ptr(None), ptr(None)))
        rowptr, colidx = np.empty(nrows.value + 1, dtype=np.int32), np.empty(nnz.value, dtype=np.int32)
        chk(load().dB200OpDensePattern(self.h, ifield, ptr(l2g), C.byref(nrows), C.byref(nnz), ptr(rowptr), ptr(colidx)))
        return rowptr, colidx

    def assemble_dense(self, which, xs, l2g, rowptr, colidx, vals=None):
        xs, _, xp, _ = self._vecs(xs, None)
        vals = np.empty(len(colidx) * (9 if which == DENSE_ELAST_FAST else 1)) if vals is None else vals
        chk(load().dB200OpAssembleDense(self.h, which, xp, ptr(_i32(l2g)), ptr(_i32(rowptr)), ptr(_i32(colidx)), ptr(vals)))
        return vals

    def stokes_lumped_mass(self, out=None):
        """Mdiag of StokesJacobianAssemble_Velocity (stokes.c:721-733)."""
        out = np.empty(self.spaces[0].n * 3) if out is None else out
        chk(load().dB200OpStokesLumpedMass(self.h, ptr(out)))
        return out

    def close(self):
        if self.h:
            chk(load().dB200OpDestroy(C.byref(self.h)))


class BoxMesh:
    """genblock box + dFSBuildSpace_Cont numbering (natural order) from the C builder (csrc/boxmesh.cpp)."""

    def __init__(self, ne, p, procs=(1, 1, 1), rank=0, dirichlet=WALLS | TOP | BOTTOM, box=(-1, 1, -1, 1, -1, 1), morph=0.0, twist=0.0, node_xyz=True, ghost_l2g=True):
        spec = BoxSpec()
        spec.ne[:] = list(ne)
        spec.p = p
        spec.procs[:] = list(procs)
        spec.rank = rank
        spec.dirichlet_mask = dirichlet
        spec.box[:] = list(box)
        spec.morph, spec.twist = morph, twist
        spec.flags = (0 if node_xyz else 1) | (0 if ghost_l2g else 2)
        m = C.POINTER(BoxMeshC)()
        chk(load().dB200BoxMeshCreate(C.byref(spec), C.byref(m)))
        c = m.contents
        self.ne, self.p, self.procs, self.rank, self.bs = tuple(ne), p, tuple(procs), rank, 1
        self.nelem, self.n, self.nc, self.ngh, self.rstart, self.nglobal = c.nelem, c.n, c.nc, c.ngh, c.rstart, c.nglobal
        d3, nloc = (p + 1) ** 3, c.nc + c.ngh
        self._c = m                       # the arrays below are views of the C mesh; freed in __del__
        self.e2l = np.ctypeslib.as_array(c.e2l, shape=(max(c.nelem * d3, 1),))[:c.nelem * d3]
        self.l2g = np.ctypeslib.as_array(c.l2g, shape=(max(nloc, 1),))[:nloc] if c.l2g else None     # None: > 2^31 global nodes, matrix-free only
        self.coords = np.ctypeslib.as_array(c.coords, shape=(max(c.nelem, 1), 8, 3))[:c.nelem]
        self.node_xyz = np.ctypeslib.as_array(c.node_xyz, shape=(max(nloc, 1), 3))[:nloc] if node_xyz else None
        self.halos = []
        for i in range(c.nneigh):
            s = np.ctypeslib.as_array(c.sendidx[i], shape=(max(c.sendcount[i], 1),))[:c.sendcount[i]].copy()
            r = np.ctypeslib.as_array(c.recvidx[i], shape=(max(c.recvcount[i], 1),))[:c.recvcount[i]].copy()
            self.halos.append((c.neigh_rank[i], s, r))

    def __del__(self):
        m = getattr(self, "_c", None)
        if m is not None and _lib is not None:
            self.e2l = self.l2g = self.coords = self.node_xyz = None
            _lib.dB200BoxMeshDestroy(C.byref(m))
            self._c = None

    @property
    def nloc(self):
        return self.nc + self.ngh
