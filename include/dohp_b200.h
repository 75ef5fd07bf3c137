/* dohp_b200.h — C ABI of the B200-native (sm_100a) implementation of Dohp's matrix-free element hot path.
 *
 * Plain C: no PETSc, MOAB or torch types.  Every function returns dErr-style int (0 = success, like PetscErrorCode,
 * /root/reference/include/dohptype.h:63); dB200GetLastError() gives the message a dERROR traceback would have printed.
 * The caller owns every buffer it passes; the library owns device memory behind opaque handles.  Any `double*`/`int*`
 * DATA argument may point to host memory (pageable or pinned) or to device memory of the context's GPU: the library
 * detects which (cudaPointerGetAttributes) and stages host buffers through HBM itself.  Work is enqueued on the
 * context's stream; calls that received a host OUTPUT pointer return after the copy-back completed, calls with device
 * outputs return asynchronously (dB200Synchronize).
 *
 * Citations are to the reference tree (jedbrown/dohp @ 005e3590); each entry point names the reference interface it
 * replaces.  The granularity differs from the reference on purpose: dEFSApply is one element per call
 * (src/jacobi/interface/jacobi.c:489-506) and the physics is inline C between dRulesetIteratorGetPatchApplied and
 * CommitPatchApplied (src/fs/tests/ellip.c:456-468); a GPU needs the whole traversal per call, so
 * the boundary is "one call per dRulesetIterator traversal" with the pointwise physics selected by `kind`.
 *
 * Layout conventions are the reference's (SURVEY.md 8 header): element-local node (i,j,k) -> ((i*P1+j)*P2+k)*D+d,
 * k fastest, component innermost; quadrature point q = (qi*Q1+qj)*Q2+qk; gradients [q][d][e] (direction e innermost);
 * jinv [q][xi-direction][x-direction] = d xi_k / d x_e (src/jacobi/impls/tensor/efstopo.c:557-591).
 */
#ifndef DOHP_B200_H
#define DOHP_B200_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct dB200Ctx_s   *dB200Ctx;     /* dInitialize + the GPU/stream/communicator                              */
typedef struct dB200Rule_s  *dB200Rule;    /* dRule   (tensor-product hex rule; include/dohpjacimpl.h:17-33)         */
typedef struct dB200EFS_s   *dB200EFS;     /* dEFS    (dEFS_Tensor = 3 TensorBasis + rule; tensorimpl.h:17-30)       */
typedef struct dB200Space_s *dB200Space;   /* dFS     (E as index array, local layout, ghost scatter; dohpfsimpl.h)  */
typedef struct dB200Op_s    *dB200Op;      /* the dRulesetIterator + stash + pointwise physics of one example        */

/* dApplyMode, include/dohpjacobi.h:101-108; InsertMode = PETSc's */
enum { dB200_APPLY_INTERP = 40001, dB200_APPLY_INTERP_TRANSPOSE, dB200_APPLY_GRAD, dB200_APPLY_GRAD_TRANSPOSE };
enum { dB200_INSERT_VALUES = 1, dB200_ADD_VALUES = 2 };
/* dQuadratureMethod (include/dohpjacobi.h:140-148) and dGaussFamily (:215) */
enum { dB200_QMETHOD_FAST = 0, dB200_QMETHOD_SPARSE = 1, dB200_QMETHOD_SELF = 2 };
enum { dB200_GAUSS_LEGENDRE = 0, dB200_GAUSS_LOBATTO = 1 };
/* dFSHomogeneousMode, include/dohpfs.h:48 */
enum { dB200_HOMOGENEOUS = 0, dB200_INHOMOGENEOUS = 1 };

const char *dB200GetLastError(void);

/* ---- context ------------------------------------------------------------------------------------------------------ */
/* dInitialize (src/util/dohpinit.c:7-18).  Fails (never falls back to a CPU path) when `device` is not an sm_100 GPU. */
int dB200Init(int device, dB200Ctx *ctx);
/* Multi-GPU: one process per GPU.  `nccl_unique_id` is the 128-byte ncclUniqueId produced by dB200NcclUniqueId on rank 0
 * and broadcast by the host program (torch.distributed / MPI); replaces the MPI communicator of dFSCreate. */
int dB200NcclUniqueId(void *id128);
int dB200InitComm(int device, int rank, int nranks, const void *nccl_unique_id, dB200Ctx *ctx);
int dB200Destroy(dB200Ctx *ctx);
int dB200SetStream(dB200Ctx ctx, void *cuda_stream);      /* run on the caller's stream (e.g. torch's current stream) */
int dB200Synchronize(dB200Ctx ctx);
int dB200NumLaunches(dB200Ctx ctx, long long *nlaunches); /* kernels launched by this context so far */

/* ---- 1-D tables (host arithmetic; setup only) --------------------------------------------------------------------- */
/* TensorGetRule, src/jacobi/quadrature/impls/tensor/tensorquad.c:51-118: FAST size=1+order/2 Gauss(-Lobatto);
 * SPARSE size=order, 2-point Gauss in every GLL sub-interval; SELF = GLL.  coord/weight need room for `order+2`. */
int dB200Rule1D(int method, int family, int order, int *size, double *coord, double *weight);
/* TensorBasisCreate, src/jacobi/impls/tensor/tensor.c:22-109: GLL nodal basis with P nodes tabulated at qcoord[Q]:
 * interp[Q][P], deriv[Q][P] = interp . D_gll, node[P], weight[P] (any output may be NULL).  2 <= P <= 14. */
int dB200Basis1D(int P, int Q, const double *qcoord, double *interp, double *deriv, double *node, double *weight);

/* ---- dRule / dEFS ----------------------------------------------------------------------------------------------------- */
int dB200RuleCreate(dB200Ctx ctx, const int Q[3], const double *const coord[3], const double *const weight[3], dB200Rule *rule);
int dB200RuleDestroy(dB200Rule *rule);
/* Tables are the caller's (the reference's TensorBasis arrays can be passed as they are): interp[d] and deriv[d] are
 * [Q[d]][P[d]] row-major.  The GPU kernels evaluate derivatives by collocation on the rule's points, which is exact for
 * the polynomial bases the reference builds; EFSCreate verifies deriv == Dcolloc . interp to 1e-12 and errors otherwise. */
int dB200EFSCreate(dB200Ctx ctx, dB200Rule rule, const int P[3], const double *const interp[3], const double *const deriv[3], dB200EFS *efs);
int dB200EFSDestroy(dB200EFS *efs);
/* dEFSGetSizes, src/jacobi/impls/tensor/efstopo.c:290-304 */
int dB200EFSGetSizes(dB200EFS efs, int P[3], int Q[3]);
/* dEFSApply (include/dohpjacobi.h:177; dEFSApply_Tensor_Hex efstopo.c:368-423) over `nelem` elements stored
 * back to back.  jinv: [nelem][Q][9] or NULL (GRAD with NULL = reference-space gradient, TensorRuleNoMapping
 * efstopo.c:518-546; GRAD_TRANSPOSE with NULL is an error as upstream, efstopo.c:600). */
int dB200EFSApplyBatch(dB200EFS efs, int nelem, const double *jinv, int D, const double *in, double *out, int amode, int imode);
/* Lower-dimensional topologies (what facet integrals / weak boundaries evaluate).  dim = 1: dTOPO_LINE, 2: dTOPO_QUAD
 * (dQuadratureGetRules_Tensor_Private tensorquad.c:186-195, dJacobiGetEFS_Tensor tensor.c:522-531): Q/coord/weight and P/interp/deriv
 * carry `dim` entries.  dB200EFSApplyBatch on such a dEFS follows dEFSApply_Tensor_Line / _Quad (efstopo.c:290-366): layouts
 * in[(i*P1+j)][D], out[(qi*Q1+qj)][D]; GRAD out [q][D][3] in reference coordinates -- upstream's TensorRuleNoMapping (efstopo.c:518-546)
 * keeps the 3-direction stride on every topology and reads the directions >= dim out of bounds (garbage); here they are written as
 * exact zeros.  jinv must be NULL (upstream: "Not implemented", efstopo.c:638,653) and GRAD_TRANSPOSE is an invalid mode
 * (efstopo.c:321-325, 361-363). */
int dB200RuleCreateDim(dB200Ctx ctx, int dim, const int Q[], const double *const coord[], const double *const weight[], dB200Rule *rule);
int dB200EFSCreateDim(dB200Ctx ctx, dB200Rule rule, int dim, const int P[], const double *const interp[], const double *const deriv[], dB200EFS *efs);
/* dQuadratureGetFacetRules (src/jacobi/quadrature/impls/tensor/tensorquad.c:309-355, 225-243): points of the QUAD rule (Q[2], coord,
 * weight) on hex facet 0..5, in the hex's reference coordinates: nodes[Q0*Q1][3], weights[Q0*Q1] (weights may be NULL).  A hex dEFS whose
 * tables are tabulated at these points (dB200Basis1D at the facet's coordinates, Q = 1 in the normal direction) gives traces and normal
 * derivatives through dB200EFSApplyBatch. */
int dB200FacetRuleNodes(int facet, const int Q[2], const double *const coord[2], const double *const weight[2], double *nodes, double *weights);
/* dRuleComputePhysical (src/jacobi/interface/jacobi.c:383-414) over nelem elements: jac [nelem][Q][9] from a
 * coordinate GRAD with NULL jinv -> jinv [nelem][Q][9], jw [nelem][Q] = w0 w1 w2 det.  det<=0 anywhere -> error. */
int dB200RuleComputePhysicalBatch(dB200Rule rule, int nelem, const double *jac, double *jinv, double *jw);

/* ---- dFS ------------------------------------------------------------------------------------------------------------------ */
/* A continuous-Galerkin space as the kernels see it: nelem hexes of P^3 nodes each, block size
 * bs, the boolean element-assembly matrix E as e2l[nelem*P^3] (expanded node -> local node, src/jacobi/impls/tensor/
 * tensor.c:360-468) and the local layout [explicit n | dirichlet nc-n | ghost ngh] (src/vec/vecd.c:123-151).
 * Global vectors passed to the operator calls hold the n*bs owned explicit dofs (VecDohp global form). */
int dB200SpaceCreate(dB200Ctx ctx, int nelem, int P, int bs, int n, int nc, int ngh, const int *e2l, dB200Space *space);
int dB200SpaceDestroy(dB200Space *space);
/* VecGhost scatter of the closure vector (src/fs/interface/fs.c:578-579 forward INSERT, :627-628 reverse ADD):
 * neighbour r sends the nodes in sendidx (local indices < n that the neighbour ghosts) and receives recvidx (local
 * indices >= nc); both sides list a shared node set in the same order.  COLLECTIVE over the ranks of the communicator when the context has
 * one: the neighbours' ghost buffers are mapped over NVLink (cudaIpc) so that the exchanges are direct stores into peer memory
 * (csrc/p2p.cu; DOHP_B200_P2P=0 or unmappable peers: grouped ncclSend/ncclRecv). */
int dB200SpaceSetHalo(dB200Space space, int nneigh, const int *ranks, const int *sendcount, const int *const *sendidx,
                      const int *recvcount, const int *const *recvidx);
/* How this space exchanges ghosts: 0 = no neighbours, 1 = grouped ncclSend/ncclRecv, 2 = direct stores into peer memory over NVLink. */
int dB200SpaceHaloMode(dB200Space space, int *mode);
/* dFSGlobalToExpanded followed by dFSExpandedToGlobal pieces, exposed for tests:
 *   local[(nc+ngh)*bs] <- global (HOMOGENEOUS: zero Dirichlet part; INHOMOGENEOUS: dirichlet[(nc-n)*bs]) + ghost update */
int dB200SpaceGlobalToLocal(dB200Space space, const double *g, const double *dirichlet, int hmode, double *local);
/* global[n*bs] (+)= owned part of local after the reverse ADD of ghosts */
int dB200SpaceLocalToGlobal(dB200Space space, const double *local, double *g, int imode);

/* dFSRotation (slip boundaries; src/fs/interface/fsrot.c:4-37, dFSRotationApplyLocal :132-198): rotate the bs-blocks of the n local
 * nodes idx[] of local[(nc+ngh)*bs] by rmat[n][bs][bs] (row-major, orthogonal; the rows are e.g. normal and tangents) and enforce the
 * first ns[i] rotated components strongly: FORWARD (reverse = 0): v <- R v, then those components = 0 (HOMOGENEOUS) or the next ns[i]
 * values of strong[] (INHOMOGENEOUS); REVERSE: those components are set likewise, then v <- R^T v.  This is the behaviour the
 * reference's header documents; its loop body (fsrot.c:160-186) indexes rmat by node number, counts with the strong VALUES and does not
 * form R^T v, and no upstream test exercises it: parity-unpinned. */
int dB200SpaceRotationApply(dB200Space space, int n, const int *idx, const double *rmat, const int *ns, const double *strong, double *local,
                            int reverse, int hmode);

/* ---- operators ------------------------------------------------------------------------------------------------------------ */
enum { dB200_OP_ELLIP = 1, dB200_OP_ELAST = 2, dB200_OP_STOKES = 3,
       /* L2 projection of src/fs/tests/ex1.c (ProjResidual{,2,3}, ex1.c:294-430) on one scalar space, param = NULL: dB200OpFunction gives
        * B^T W (B x - f) with f = the forcing array at the quadrature points (dB200OpSetForcing; NULL = 0), dB200OpMult the mass operator
        * B^T W B x (no stash), dB200OpAssembleQ1 the Q1 mass matrix of ProjJacobian1 (ex1.c:432-482; the state vector is ignored) */
       dB200_OP_PROJECTION = 4 };
/* EllipParam, src/fs/tests/ellip.c:119-160 (defaults p=2, epsilon=1, lambda=0 => linear Poisson) */
typedef struct { double p, epsilon, lambda; } dB200EllipParam;
/* ElastParam, src/fs/tests/elast.c:62-70: Lame parameters from E0, nu */
typedef struct { double lambda, mu; } dB200ElastParam;
/* struct StokesRheology {B, eps, p, nu}, src/fs/tests/stokes.c:56-61 (defaults 1, 1, 2, 0: stokes.c:122-125) */
typedef struct { double B, eps, p, nu; } dB200StokesParam;
/* Stokes block selector (StokesShellMatMult_All_IorA, stokes.c:579-673) */
enum { dB200_STOKES_J = 0, dB200_STOKES_A = 1, dB200_STOKES_B = 2, dB200_STOKES_Bt = 3 };

/* XGetRegionIterator (ellip.c:289-316, elast.c:165-190, stokes.c:283-304): binds the rule, the spaces (stokes:
 * spaces[0]=velocity bs 3, spaces[1]=pressure bs 1, both tabulated on the velocity rule), their dEFS, and the expanded
 * Q1 geometry vector coords[nelem][8][3] in tensor order (dFSGetGeometryVectorExpanded, src/fs/interface/
 * fsgeom.c:111-170, 216-227).  Allocates the stash (dRulesetIteratorAddStash, fsrulesetit.c:779-798). Checks det J > 0
 * at every quadrature point once here (dRuleComputePhysical errors on it at every traversal, jacobi.c:397). */
int dB200OpCreate(dB200Ctx ctx, int kind, const void *param, dB200Rule rule, int nspaces, const dB200Space *spaces,
                  const dB200EFS *efs, const double *coords, dB200Op *op);
int dB200OpDestroy(dB200Op *op);
/* Physical coordinates of the quadrature points, xq[nelem][Q][3] (the `x` the reference hands to the user's forcing
 * function, fsrulesetit.c:549 link 0 INTERP). */
int dB200OpGetQuadraturePoints(dB200Op op, double *xq);
/* Forcing evaluated by the caller at those points, f[nelem][Q][nf] (ellip nf=1: f; elast nf=3; stokes nf=4: fu[3],fp):
 * replaces the `rhs`/`forcing` function pointers of the examples (ellip.c:361, elast.c:238, stokes.c:466). NULL = 0. */
int dB200OpSetForcing(dB200Op op, const double *f);
/* Dirichlet values of field `ifield`, vals[(nc-n)*bs] (the dcache of dFSDirichletProject/dFSInhomogeneousDirichletCommit,
 * src/fs/interface/fs.c:530-541,666-693). */
int dB200OpSetDirichlet(dB200Op op, int ifield, const double *vals);
/* Residual: XFunction (ellip.c:407-440, elast.c:282-309, stokes.c:507-539): gx INHOMOGENEOUS, gy overwritten, stash
 * written.  gx[i]/gy[i] are the global vectors of space i (n_i*bs_i doubles). */
int dB200OpFunction(dB200Op op, const double *const gx[], double *const gy[]);
/* Matrix-free Jacobian product: XShellMatMult (ellip.c:442-472, elast.c:311-341, MatMult_Nest_StokesCoupled
 * stokes.c:541-577): HOMOGENEOUS, reads the stash, gy overwritten (VecZeroEntries + ADD of every element). */
int dB200OpMult(dB200Op op, const double *const gx[], double *const gy[]);
/* Stokes blocks A, B, Bt (stokes.c:579-673); for J same as dB200OpMult.  Unused vectors may be NULL. */
int dB200OpMultBlock(dB200Op op, int block, const double *const gx[], double *const gy[]);
/* dB200OpMult for HOST vectors in pinned memory without blocking: upload on a copy stream, traversal on the context's stream,
 * download on a second copy stream, ordered by events.  Products of different operators overlap (both PCIe directions and the
 * kernels); a second product on the same operator waits for the first.  gx/gy must stay valid until dB200Synchronize(ctx)
 * (host) or until work enqueued after dB200StreamJoin(ctx) on the context's stream (device). */
int dB200OpMultAsync(dB200Op op, const double *const gx[], double *const gy[]);
int dB200StreamJoin(dB200Ctx ctx);
/* Stash access for tests (layout [nelem][Q][nstash] as the reference's per-node stash struct, e.g. struct EllipStore
 * ellip.c:119-123: eta, sqrt_mdeta_Du[3], lambda_exp_u). */
int dB200OpGetStash(dB200Op op, int *nstash, double *stash);

/* ---- embedded-Q1 preconditioner (XJacobian: ellip.c:474-536, elast.c:343-407) ------------------------------------------ */
/* CSR pattern of the matrix dFSCreateMatrix + the patch loop + MatAssembly produce (src/fs/interface/fs.c:770-813,
 * 826-930): per scalar row the sorted union of the columns of every Q1 patch containing the row's node, zeros kept,
 * Dirichlet rows/columns dropped (negative ltog, fsbuild.c:31-34).  l2g[nc+ngh] is the ltog map of space 0; columns are
 * global indices.  Block CSR for bs > 1 (one block per coupled node pair = BAIJ(bs), elast.c:479).
 * Rows: the n owned nodes, then -- on a partitioned space -- one row per ghost node (local node nc+g, global row
 * l2g[nc+g]) holding this rank's contributions to a row owned by a neighbour: what MatSetValuesBlockedLocal stashes until
 * MatAssemblyEnd (fs.c:919) and what MatSetPreallocationCOO/MatSetValuesCOO accept as off-process entries.  nrows = n+ngh.
 * Returns sizes first (rowptr/colidx NULL), then fills caller arrays.  The pattern (and the row records of the fast assembly path)
 * is cached per operator together with a checksum of l2g: a call with a different map rebuilds it; dB200OpAssembleQ1 on the cached
 * pattern with another map is an error.  A map whose ghost entries are all negative (BoxSpec.flags bit 1) is rejected when ngh > 0. */
int dB200OpQ1Pattern(dB200Op op, const int *l2g, long long *nrows, long long *nnz, int *rowptr, int *colidx);
/* Assemble values at linearisation state gx into vals[nnz*bs*bs] (zeroed by the callee, MatZeroEntries ellip.c:489).
 * rowptr/colidx NULL = the pattern cached by dB200OpQ1Pattern (fast path: the per-node row records are kept); otherwise
 * any block CSR with sorted columns that contains the couplings.  Ghost values of gx are exchanged first (fs.c:578-579).
 * Degrees 1..10. */
int dB200OpAssembleQ1(dB200Op op, const double *const gx[], const int *l2g, const int *rowptr, const int *colidx, double *vals);
/* -dfs_assemble_reduced (dFSMatSetValuesBlockedExpanded, fs.c:905-917): only the entries (row*bs+k, col*bs+k) of every block are
 * inserted; here the off-diagonal entries of the blocks stay zero (and the off-diagonal component pairs are not computed). */
int dB200OpQ1SetReduced(dB200Op op, int flag);
/* MatAssemblyBegin/End of the preconditioning matrix (ellip.c:528-529, elast.c:401-402, stokes.c:735-736) on a partitioned space:
 * the ghost rows [n, n+ngh) of the local block CSR are this rank's contributions to rows a neighbour owns (the MatSetValues stash,
 * fs.c:919).  OwnerPattern (once per pattern; collective over the ranks of the communicator) sends every ghost row's column list to
 * the node's owner over NCCL and merges it into the owner's row: orowptr[n+1], ocolidx[onnz] = the n owned rows of the GLOBAL matrix
 * (sorted global columns), sizes first with NULL arrays.  rowptr/colidx = the local pattern of dB200OpQ1Pattern (needed on the first
 * call).  OwnerReduce (collective, every assembly): vals[(local nnz)*bs*bs] from dB200OpAssembleQ1 -> pack the ghost rows ->
 * grouped ncclSend/ncclRecv -> ADD into ovals[onnz*bs*bs], neighbour by neighbour in rank order (deterministic sums).
 * Without a communicator (one rank) the pattern and the values are copied through. */
int dB200OpQ1OwnerPattern(dB200Op op, const int *rowptr, const int *colidx, long long *onnz, int *orowptr, int *ocolidx);
int dB200OpQ1OwnerReduce(dB200Op op, const double *vals, double *ovals);

/* Stokes (StokesJacobianAssemble_Velocity, src/fs/tests/stokes.c:675-737): dB200OpQ1Pattern / dB200OpAssembleQ1 on a
 * dB200_OP_STOKES operator give the velocity block Ap (BAIJ-3 on spaces[0], StokesPointwiseJacobian_A incl. the nu penalty,
 * gx[0] = velocity state, INHOMOGENEOUS); dB200OpStokesLumpedMass gives the Mdiag vector of the same traversal
 * (stokes.c:721-733), mdiag[n_u*3], overwritten.  The pressure blocks D / Daux (stokes.c:739-794) come from the element-dense
 * path below (dB200_DENSE_STOKES_D / _DAUX). */
int dB200OpStokesLumpedMass(dB200Op op, double *mdiag);

/* ---- element-dense assembly: the dEFSGetExplicitSparse_Basic path (src/jacobi/interface/jacobi.c:611-643) --------------------
 * Used by the reference where the rule's patches do not line up with Q1 sub-elements: every pair of nodes sharing an element
 * couples.  dB200_DENSE_ELLIP_FAST: EllipJacobian with -ellip_jac_qmethod fast (ellip.c:474-536; = the exact Jacobian matrix,
 * ellip-morph-qfast.refout).  dB200_DENSE_STOKES_D / _DAUX: StokesJacobianAssemble_Pressure (stokes.c:739-794): the scaled
 * pressure mass matrix and the Neumann Laplacian on the pressure space (spaces[1]), state gx[0] = velocity.
 * dB200_DENSE_ELAST_FAST: ElastJacobian with -elast_jac_qmethod fast (elast.c:343-407 on the FAST rule): BAIJ(3) blocks on the same node
 * pattern, vals[nnz*9], one pass of the kernels per component pair.
 * Pattern: scalar CSR, rows as for dB200OpQ1Pattern (owned, then ghost rows), field = 0 (ellip, elast) or 1 (stokes pressure).
 * Matrix-field degree <= 4: the element matrix is formed in shared memory; higher degrees: column by column (any degree). */
enum { dB200_DENSE_ELLIP_FAST = 1, dB200_DENSE_STOKES_D = 2, dB200_DENSE_STOKES_DAUX = 3, dB200_DENSE_ELAST_FAST = 4 };
int dB200OpDensePattern(dB200Op op, int ifield, const int *l2g, long long *nrows, long long *nnz, int *rowptr, int *colidx);
int dB200OpAssembleDense(dB200Op op, int which, const double *const gx[], const int *l2g, const int *rowptr, const int *colidx, double *vals);

/* ---- synthetic structured meshes (genblock, src/fs/mesh/interface/genblock.c:208-420, + dFSBuildSpace_Cont numbering,
 *      src/fs/impls/cont/cont.c:93-385, natural ordering) — setup helper for benchmarks and tests ----------------------- */
typedef struct {
  int ne[3];            /* global element grid */
  int p;                /* polynomial degree */
  int procs[3], rank;   /* block partition (genblock.c:257-304), part id (I*N+J)*P+K */
  int dirichlet_mask;   /* bit0: set 100 (x/y walls), bit1: set 200 (z+), bit2: set 300 (z-) */
  double box[6];        /* x0,x1,y0,y1,z0,z1 */
  double morph, twist;  /* Morph(), src/fs/tests/ellip.c:173-181 */
  int flags;            /* bit0: skip node_xyz; bit1: skip the ghost entries of l2g (only Q1 assembly reads them) */
} dB200BoxSpec;
typedef struct {
  int nelem, n, nc, ngh;
  long long rstart, nglobal;
  int *e2l;             /* [nelem*(p+1)^3] */
  int *l2g;             /* [nc+ngh], negative for Dirichlet */
  double *coords;       /* [nelem][8][3] */
  double *node_xyz;     /* [nc+ngh][3] nodal coordinates (dFSGetNodalCoordinates, fsgeom.c:20-106) */
  int nneigh, *neigh_rank, *sendcount, *recvcount, **sendidx, **recvidx;
} dB200BoxMesh;
int dB200BoxMeshCreate(const dB200BoxSpec *spec, dB200BoxMesh **mesh);
int dB200BoxMeshDestroy(dB200BoxMesh **mesh);

/* FP64 FMA peak of the device measured by a register-resident DFMA loop (TFLOP/s): the roofline denominator that
 * MEASURED_PEAKS.json does not carry. */
int dB200MeasureFP64Peak(dB200Ctx ctx, double *tflops);

#ifdef __cplusplus
}
#endif
#endif
