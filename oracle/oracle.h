/* TEST INFRASTRUCTURE — CPU oracle for dohp-b200.
 *
 * Plain-C restatement of the reference's matrix-free hot path (SURVEY.md §8a rows A1–A15, Q1-a..e).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library; the product
 * (dohp_b200/, include/) never does.  Parity status: PINNED — validated against oracle/_ref (the reference's own
 * dJacobi layer, which reproduces nine upstream golden refouts byte-for-byte) and against the known answers of the
 * reference's ellip/elast/stokes goldens (tests/test_oracle_*.py).
 *
 * All citations are relative to /root/reference.  Layout conventions (SURVEY.md §8 header): element-local node
 * (i,j,k) -> ((i*P1+j)*P2+k)*D+d, k fastest, component innermost; gradients [q][d][e] with direction e innermost;
 * jinv [q][xi-direction][x-direction].
 */
#ifndef DOHP_ORACLE_H
#define DOHP_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

/* dApplyMode / InsertMode values (include/dohpjacobi.h:101-108; PETSc InsertMode) */
enum {OR_APPLY_INTERP = 40001, OR_APPLY_INTERP_TRANSPOSE, OR_APPLY_GRAD, OR_APPLY_GRAD_TRANSPOSE};
enum {OR_INSERT_VALUES = 1, OR_ADD_VALUES = 2};
/* dQuadratureMethod (include/dohpjacobi.h:140-148), dGaussFamily (:215) */
enum {OR_QMETHOD_FAST = 0, OR_QMETHOD_SPARSE = 1, OR_QMETHOD_SELF = 2};
enum {OR_GAUSS_LEGENDRE = 0, OR_GAUSS_LOBATTO = 1};

typedef struct OrEFS_s *OrEFS;

/* --- A1/A9: 1-D tables ------------------------------------------------------------------------------------------- */
int or_gauss_lobatto(int np,double *z,double *w);                       /* polylib.c:272-295 (alpha=beta=0) */
int or_gauss_legendre(int np,double *z,double *w);                      /* polylib.c:182-196 */
int or_rule_1d(int method,int family,int order,int *size,double *coord,double *weight);   /* tensorquad.c:51-118 */
int or_basis_1d(int P,int Q,const double *qcoord,double *interp,double *deriv,double *node,double *weight); /* tensor.c:22-72 */

/* --- A2: element function space = (basis degree, rule) ------------------------------------------------------------ */
int or_efs_create(const int bdeg[3],const int rorder[3],int method,int family,OrEFS *efs);
int or_efs_destroy(OrEFS *efs);
int or_efs_get_sizes(OrEFS efs,int P[3],int Q[3]);
int or_efs_get_tables(OrEFS efs,int dir,double *interp,double *deriv,double *node,double *weight);
int or_efs_get_rule(OrEFS efs,int dir,double *coord,double *weight);
int or_efs_get_patches(OrEFS efs,int *npatches,int *patchsize,const int **patchind,const double **patchweight); /* tensorquad.c:120-169 */

/* --- A3-A7 -------------------------------------------------------------------------------------------------------- */
int or_tensor_mult_hex(int D,const int P[3],const int Q[3],const double *const A[3],const double *in,double *out,int imode);
int or_efs_apply(OrEFS efs,const double *jinv,int D,const double *in,double *out,int amode,int imode);
int or_efs_apply_batch(OrEFS efs,int nelem,const double *jinv,int D,const double *in,double *out,int amode,int imode);
int or_rule_compute_physical(OrEFS efs,const double *jac,double *jinv,double *jw);
int or_rule_compute_physical_batch(OrEFS efs,int nelem,const double *jac,double *jinv,double *jw);

/* --- Q1-a / Q1-b --------------------------------------------------------------------------------------------------- */
int or_efs_get_explicit(OrEFS efs,const double *jinv,double *interp,double *deriv);             /* jacobi.c:508-559 */
int or_efs_get_explicit_sparse(OrEFS efs,int npatches,int Q,const int *qidx,const double *cjinv,int eoffset,
                               int *P,int *eidx,double *interp,double *deriv);                 /* efstopo.c:660-717, jacobi.c:592-643 */

/* --- A11: function-space data flow on index arrays ---------------------------------------------------------------- */
/* A space is described by: nelem elements of np = P0*P1*P2 nodes each, block size bs, the boolean element-assembly
 * matrix E as the index array e2l[nelem*np] (expanded node -> local node) and the local layout
 * [explicit n | dirichlet nc-n | ghost ngh] (src/vec/vecd.c:123-151). */
typedef struct {
  int nelem,bs;
  int n,nc,ngh;
  const int *e2l;
} OrSpace;
enum {OR_HOMOGENEOUS = 0, OR_INHOMOGENEOUS = 1};
int or_fs_global_to_local(const OrSpace *s,const double *g,const double *dcache,int hmode,double *local); /* fs.c:543-579 (ghost part left to caller) */
int or_fs_local_to_expanded(const OrSpace *s,int np,const double *local,double *x);                    /* fs.c:502-521 */
int or_fs_expanded_to_local(const OrSpace *s,int np,const double *x,double *local,int imode);          /* fs.c:481-500 */

/* --- A12-A15: operators --------------------------------------------------------------------------------------------- */
typedef struct {
  double p,epsilon,lambda;        /* ellip.c:156-158 */
  int    exact;                   /* 0..4, ellip.c:246-268 */
  double a,b,c;                   /* ellip.c:199 */
} OrEllipParam;
#define OR_ELLIP_STORE 5          /* doubles per quadrature point: eta, sqrt_mdeta_Du[3], lambda_exp_u (ellip.c:119-123) */

void or_ellip_exact_solution(const OrEllipParam *prm,const double x[3],double u[1],double du[3]);
void or_ellip_exact_forcing(const OrEllipParam *prm,const double x[3],double f[1]);

/* coords: expanded geometry vector [nelem][8][3] in tensor order (fsgeom.c:150-152); xloc/yloc local-form vectors.
 * yloc is ACCUMULATED into (callers zero it: ellip.c:421,452). */
int or_ellip_function(OrEFS cefs,OrEFS efs,const OrSpace *s,const double *coords,const OrEllipParam *prm,
                      const double *xloc,double *yloc,double *stash);                                  /* ellip.c:407-440 */
int or_ellip_mult(OrEFS cefs,OrEFS efs,const OrSpace *s,const double *coords,const OrEllipParam *prm,
                  const double *stash,const double *xloc,double *yloc);                                /* ellip.c:442-472 */
/* EllipErrorNorms, ellip.c:538-564 (dNormsUpdate, src/util/util.c:119-152): the pointwise error lines every example prints after its solve */
int or_ellip_error_norms(OrEFS cefs,OrEFS efs,const OrSpace *s,const double *coords,const OrEllipParam *prm,const double *xloc,double enorm[3],double gnorm[3]);
/* Q1 assembly (ellip.c:474-536): cefs_s/efs_s are built on the SPARSE rule.  Element matrices are added into CSR
 * (rowptr,colidx over explicit dofs 0..n) through l2g (local node -> global row or <0 for Dirichlet, fsbuild.c:27-37).
 * qmethod_basic!=0 forces the dense dEFSGetExplicitSparse_Basic route. */
int or_ellip_jacobian_q1(OrEFS cefs_s,OrEFS efs_s,const OrSpace *s,const double *coords,const OrEllipParam *prm,
                         const double *xloc,const int *l2g,const int *rowptr,const int *colidx,double *vals);

/* L2 projection, the dFS integration test src/fs/tests/ex1.c: residual v = (u - f) jw (three API styles, ex1.c:294-430), its derivative
 * (the mass operator -snes_mf differences), the pointwise residual norms the test prints (ex1.c:484-532) and the Q1 mass matrix of
 * ProjJacobian1 (ex1.c:432-482).  Pinned by fs-mf-projection-v{1,3}.refout and fs-mf-op-proj-{0,1}.refout. */
typedef struct {
  int    exact;                   /* 0 transcendental (ex1.c:33-46), 1 linear (ex1.c:48-51) */
  double a,b,c;                   /* -frequency, ex1.c:643 */
} OrProjParam;
void or_proj_exact(const OrProjParam *prm,const double x[3],double f[1],double df[3]);
int or_proj_function(OrEFS cefs,OrEFS efs,const OrSpace *s,const double *coords,const OrProjParam *prm,const double *xloc,double *yloc);
int or_proj_mult(OrEFS cefs,OrEFS efs,const OrSpace *s,const double *coords,const double *xloc,double *yloc);
int or_proj_residual_norms(OrEFS cefs,OrEFS efs,const OrSpace *s,const double *coords,const OrProjParam *prm,const double *xloc,double res[3],double gres[3]);
int or_proj_jacobian_q1(OrEFS cefs_s,OrEFS efs_s,const OrSpace *s,const double *coords,const int *l2g,const int *rowptr,const int *colidx,double *vals);

/* elasticity (elast.c:220-407) */
typedef struct {
  double lambda,mu;               /* from E0, nu: elast.c:62-70 */
  int    exact;                   /* 0 (elastexact.py:66-72) or -1 for zero forcing */
  double a,b,c,scale;
} OrElastParam;
#define OR_ELAST_STORE 9          /* F stored as Du (9): elast.c:39-41 */
void or_elast_exact_solution(const OrElastParam *prm,const double x[3],double u[3],double du[9]);
void or_elast_exact_forcing(const OrElastParam *prm,const double x[3],double f[3]);
int or_elast_function(OrEFS cefs,OrEFS efs,const OrSpace *s,const double *coords,const OrElastParam *prm,
                      const double *xloc,double *yloc,double *stash);
int or_elast_mult(OrEFS cefs,OrEFS efs,const OrSpace *s,const double *coords,const OrElastParam *prm,
                  const double *stash,const double *xloc,double *yloc);
int or_elast_error_norms(OrEFS cefs,OrEFS efs,const OrSpace *s,const double *coords,const OrElastParam *prm,const double *xloc,double enorm[3],double gnorm[3]);   /* elast.c:409-438 */
int or_elast_jacobian_q1(OrEFS cefs_s,OrEFS efs_s,const OrSpace *s,const double *coords,const OrElastParam *prm,
                         const double *xloc,const int *l2g,const int *rowptr,const int *colidx,double *vals);

/* Stokes (stokes.c:447-577) */
typedef struct {
  double B,eps,p,nu;              /* struct StokesRheology: stokes.c:56-61,122-125 */
  int    exact;                   /* 0,1 (stokesexact.py) or -1 for zero forcing */
  double a,b,c,scale;
} OrStokesParam;
#define OR_STOKES_STORE 8         /* eta, deta, Du[6] : stokesimpl.h:28-31 */
void or_stokes_exact_solution(const OrStokesParam *prm,const double x[3],double u[3],double du[9],double p[1],double dp[3]);
void or_stokes_exact_forcing(const OrStokesParam *prm,const double x[3],double fu[3],double fp[1]);
int or_stokes_function(OrEFS cefs,OrEFS efsu,OrEFS efsp,const OrSpace *su,const OrSpace *sp,const double *coords,const OrStokesParam *prm,
                       const double *xu,const double *xp,double *yu,double *yp,double *stash);
/* block: 0 full coupled J (stokes.c:541-577); 1 A; 2 B; 3 Bt (stokes.c:579-673) */
int or_stokes_mult(OrEFS cefs,OrEFS efsu,OrEFS efsp,const OrSpace *su,const OrSpace *sp,const double *coords,const OrStokesParam *prm,
                   const double *stash,int block,const double *xu,const double *xp,double *yu,double *yp);

/* StokesErrorNorms, stokes.c:835-885: velocity / velocity-gradient / pressure error integrals (local-form inputs; alldirichlet: pressure up to a constant) */
int or_stokes_error_norms(OrEFS cefs,OrEFS efsu,OrEFS efsp,const OrSpace *su,const OrSpace *sp,const double *coords,const OrStokesParam *prm,
                          const double *xu,const double *xp,int alldirichlet,double enorm[3],double gnorm[3],double pnorm[3]);

/* StokesJacobianAssemble_Velocity, stokes.c:675-737: the velocity block Ap (BAIJ-3 values into the given block pattern) and the
 * "lumped mass" vector Mdiag (local form [nloc*3], added into).  Parity UNPINNED upstream: no reference test prints either. */
int or_stokes_jacobian_q1_velocity(OrEFS cefs_s,OrEFS efsu_s,const OrSpace *su,const double *coords,const OrStokesParam *prm,
                                   const double *xuloc,const int *l2g,const int *rowptr,const int *colidx,double *vals,double *mdiag_loc);
/* StokesJacobianAssemble_Pressure, stokes.c:739-794 (D and Daux, either may be NULL).  D (and the velocity block above) are pinned by
 * the reference's Stokes goldens: their ILU(0) produces every printed residual of stokes-q3q2-{fgmres,scr}.refout
 * (tests/test_stokes_refout_replay.py); Daux (PCLSC only) and Mdiag are exercised by no upstream test */
int or_stokes_jacobian_pressure(OrEFS cefs_s,OrEFS efsu_s,OrEFS efsp_s,const OrSpace *su,const OrSpace *sp,const double *coords,const OrStokesParam *prm,
                                const double *xuloc,const int *l2gp,const int *rowptr,const int *colidx,double *valsD,double *valsDaux);
int or_num_threads(void);
void or_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
