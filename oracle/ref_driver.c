/* TEST INFRASTRUCTURE (oracle).  A thin, ctypes-friendly driver (ours) over the reference's *own* dJacobi layer as
 * compiled into oracle/_ref/libdohpref.so by oracle/Makefile.  Every arithmetic result that leaves this file was
 * produced by unmodified reference code (dEFSApply, dRuleComputePhysical, dEFSGetExplicitSparse, TensorBasisCreate,
 * TensorGetRule, polylib); this file only creates the objects the way the reference's tests do
 * (src/jacobi/tests/ex3.c:22-53) and loops the per-element entry points over a batch.
 *
 * Used (a) to validate oracle/oracle.c, the plain-C restatement, and (b) as the `cpu_baseline.kind == "reference"`
 * timing arm of bench.py.  Never linked or loaded by the product.
 */
#include <dohpjacobi.h>
#include <dohpjacimpl.h>
#include <dohpsys.h>
#include <dohp.h>
#include "tensorimpl.h"     /* reference-private: lets us read the 1-D tables the reference computed */
#undef _TRULE_H
#include <pthread.h>
#include <unistd.h>

/* Static-range parallel-for on pthreads (this image has no libgomp). */
typedef struct {int lo,hi,err; void (*body)(int,int,void*,int*); void *ctx; } par_task;
static void *par_trampoline(void *vt) { par_task *t = vt; t->body(t->lo,t->hi,t->ctx,&t->err); return NULL; }
static int par_for(int nthreads,int n,void (*body)(int,int,void*,int*),void *ctx)
{
  if (nthreads < 1) nthreads = 1;
  if (nthreads > n) nthreads = n > 0 ? n : 1;
  pthread_t th[nthreads];
  par_task  task[nthreads];
  int err = 0;
  for (int t=0; t<nthreads; t++) {
    task[t].lo = (int)((long)n*t/nthreads); task[t].hi = (int)((long)n*(t+1)/nthreads);
    task[t].err = 0; task[t].body = body; task[t].ctx = ctx;
    if (t) pthread_create(&th[t],NULL,par_trampoline,&task[t]);
  }
  par_trampoline(&task[0]);
  for (int t=1; t<nthreads; t++) pthread_join(th[t],NULL);
  for (int t=0; t<nthreads; t++) err |= task[t].err;
  return err;
}

/* tensorquad.h cannot be included next to tensorimpl.h (both instantiate khash "tensor"); restate the two layouts we
 * read (src/jacobi/quadrature/impls/tensor/tensorquad.h:12-27). */
struct ref_TensorRule { dInt size; dReal *weight,*coord; };
typedef struct { dRuleHEADER; dEntTopology topo; dInt npatches,patchsize; dInt *patchind; dReal *patchweight; struct ref_TensorRule *trule[3]; } ref_dRule_Tensor;

static dJacobi jac;

int RefInit(int gauss_family_lobatto)
{
  dErr err;
  if (jac) {err = dJacobiDestroy(&jac);dCHK(err);}
  OracleShimClearOptions();
  OracleShimSetOption("-dquad_tensor_gauss_family",gauss_family_lobatto ? "lobatto" : "gauss");
  err = dJacobiCreate(PETSC_COMM_WORLD,&jac);dCHK(err);
  err = dJacobiSetFromOptions(jac);dCHK(err);
  return 0;
}

/* bdeg: basis degree per direction; rorder: rule order per direction (dFSGetPreferredQuadratureRuleSet passes
 * 2*degree, src/fs/interface/fsruleset.c:32-37); qmethod: 0 FAST, 1 SPARSE, 2 SELF. */
int RefGetEFS(const int bdeg[3],const int rorder[3],int qmethod,void **efs_out,void **rule_out)
{
  dErr err;
  dEntTopology topo[1] = {dTOPO_HEX};
  dPolynomialOrder bd[1],rd[1];
  dQuadrature quad;
  dRule *rules;
  dEFS *efs;
  if (!jac) {err = RefInit(0);dCHK(err);}
  bd[0] = dPolynomialOrderCreate(0,bdeg[0],bdeg[1],bdeg[2]);
  rd[0] = dPolynomialOrderCreate(0,rorder[0],rorder[1],rorder[2]);
  err = dJacobiGetQuadrature(jac,(dQuadratureMethod)qmethod,&quad);dCHK(err);
  err = dQuadratureGetRules(quad,1,topo,rd,&rules);dCHK(err);
  err = dJacobiGetEFS(jac,1,topo,bd,rules,&efs);dCHK(err);
  *efs_out = efs[0];
  *rule_out = rules[0];
  err = dFree(rules);dCHK(err);
  err = dFree(efs);dCHK(err);
  return 0;
}

/* dTOPO_LINE (dim 1) / dTOPO_QUAD (dim 2) / dTOPO_HEX (dim 3) dEFS and rule, as dJacobiGetEFS_Tensor hands them out (tensor.c:503-545). */
int RefGetEFSDim(int dim,const int bdeg[3],const int rorder[3],int qmethod,void **efs_out,void **rule_out)
{
  dErr err;
  dEntTopology topo[1] = {dim == 1 ? dTOPO_LINE : dim == 2 ? dTOPO_QUAD : dTOPO_HEX};
  dPolynomialOrder bd[1],rd[1];
  dQuadrature quad;
  dRule *rules;
  dEFS *efs;
  if (!jac) {err = RefInit(0);dCHK(err);}
  bd[0] = dPolynomialOrderCreate(0,bdeg[0],bdeg[1],bdeg[2]);
  rd[0] = dPolynomialOrderCreate(0,rorder[0],rorder[1],rorder[2]);
  err = dJacobiGetQuadrature(jac,(dQuadratureMethod)qmethod,&quad);dCHK(err);
  err = dQuadratureGetRules(quad,1,topo,rd,&rules);dCHK(err);
  err = dJacobiGetEFS(jac,1,topo,bd,rules,&efs);dCHK(err);
  *efs_out = efs[0];
  *rule_out = rules[0];
  err = dFree(rules);dCHK(err);
  err = dFree(efs);dCHK(err);
  return 0;
}
int RefEFSGetPQDim(void *vefs,int dim,int P[3],int Q[3])
{
  dEFS_Tensor *e = vefs;
  for (int i=0; i<3; i++) {P[i] = i < dim ? e->basis[i]->P : 1; Q[i] = i < dim ? e->basis[i]->Q : 1;}
  return 0;
}
/* nelem dEFSApply calls on a Line / Quad / Hex dEFS with NULL jinv (the only mapping upstream implements below dimension 3).  GRAD output
 * keeps the stride [q][D][3] on every topology (TensorRuleNoMapping, efstopo.c:518-546); directions >= dim are read out of bounds upstream. */
int RefEFSApplyBatchDim(void *vefs,int dim,int nelem,int D,const double *in,double *out,int amode,int imode)
{
  int P[3],Q[3],np,nq;
  RefEFSGetPQDim(vefs,dim,P,Q);
  np = P[0]*P[1]*P[2]; nq = Q[0]*Q[1]*Q[2];
  long is = (amode == dAPPLY_INTERP || amode == dAPPLY_GRAD) ? (long)np*D : (long)nq*D;
  long os = (amode == dAPPLY_INTERP) ? (long)nq*D : (amode == dAPPLY_GRAD ? (long)nq*D*3 : (long)np*D);
  for (int e=0; e<nelem; e++) {dErr err = dEFSApply((dEFS)vefs,NULL,D,in+e*is,out+e*os,(dApplyMode)amode,(InsertMode)imode); if (err) return err;}
  return 0;
}

int RefEFSGetPQ(void *vefs,int P[3],int Q[3])
{
  dEFS_Tensor *e = vefs;
  for (int i=0; i<3; i++) {P[i] = e->basis[i]->P; Q[i] = e->basis[i]->Q;}
  return 0;
}

/* Copies of the tables computed by the reference's TensorBasisCreate (tensor.c:22-109). */
int RefEFSGetTables(void *vefs,int dir,double *interp,double *deriv,double *node,double *weight)
{
  dEFS_Tensor *e = vefs;
  struct _TensorBasis *b = e->basis[dir];
  memcpy(interp,b->interp,sizeof(double)*b->P*b->Q);
  memcpy(deriv,b->deriv,sizeof(double)*b->P*b->Q);
  memcpy(node,b->node,sizeof(double)*b->P);
  memcpy(weight,b->weight,sizeof(double)*b->P);
  return 0;
}

int RefRuleGetTensorNodeWeight(void *vrule,int dir,double *coord,double *weight)
{
  dErr err;
  dInt dim,nnodes[3];
  const dReal *c[3],*w[3];
  err = dRuleGetTensorNodeWeight((dRule)vrule,&dim,nnodes,c,w);dCHK(err);
  memcpy(coord,c[dir],sizeof(double)*nnodes[dir]);
  memcpy(weight,w[dir],sizeof(double)*nnodes[dir]);
  return 0;
}

int RefRuleGetPatches(void *vrule,int *npatches,int *patchsize,const int **ind,const double **weight)
{ return dRuleGetPatches((dRule)vrule,npatches,patchsize,ind,weight); }

/* nelem independent dEFSApply calls; element e reads in+e*instride, jinv+e*Q*9 (or NULL), writes out+e*outstride. */
typedef struct {dEFS efs; const double *jinv,*in; double *out; int D,amode,imode,nq; long instride,outstride;} apply_ctx;
static void apply_body(int lo,int hi,void *vc,int *err)
{
  apply_ctx *c = vc;
  for (int e=lo; e<hi; e++)
    *err |= dEFSApply(c->efs,c->jinv ? c->jinv+(long)e*c->nq*9 : NULL,c->D,c->in+e*c->instride,c->out+e*c->outstride,(dApplyMode)c->amode,(InsertMode)c->imode);
}
int RefEFSApplyBatch(void *vefs,int nelem,const double *jinv,int D,const double *in,double *out,int amode,int imode,int nthreads)
{
  int P[3],Q[3],np,nq;
  RefEFSGetPQ(vefs,P,Q);
  np = P[0]*P[1]*P[2]; nq = Q[0]*Q[1]*Q[2];
  apply_ctx c = {(dEFS)vefs,jinv,in,out,D,amode,imode,nq,
                 (amode == dAPPLY_INTERP || amode == dAPPLY_GRAD) ? (long)np*D : (amode == dAPPLY_INTERP_TRANSPOSE ? (long)nq*D : (long)nq*D*3),
                 (amode == dAPPLY_INTERP) ? (long)nq*D : (amode == dAPPLY_GRAD ? (long)nq*D*3 : (long)np*D)};
  return par_for(nthreads,nelem,apply_body,&c);
}

int RefRuleComputePhysicalBatch(void *vrule,int nelem,const double *jac_,double *jinv,double *jw)
{
  dErr err;
  dInt nq;
  err = dRuleGetSize((dRule)vrule,NULL,&nq);dCHK(err);
  for (int e=0; e<nelem; e++) {
    err = dRuleComputePhysical((dRule)vrule,jac_+(long)e*nq*9,jinv+(long)e*nq*9,jw+(long)e*nq);
    if (err) return err;
  }
  return 0;
}

int RefEFSGetExplicitSparse(void *vefs,int npatches,int Q,const int *qidx,const double *cjinv,int eoffset,int *P,int *eidx,double *interp,double *deriv)
{ return dEFSGetExplicitSparse((dEFS)vefs,npatches,Q,qidx,cjinv,eoffset,P,eidx,interp,deriv); }

int RefEFSGetExplicit(void *vefs,const double *jinv,int *Q,int *P,double *interp,double *deriv)
{
  dErr err;
  const dReal *b,*d;
  err = dEFSGetExplicit((dEFS)vefs,jinv,Q,P,&b,&d);dCHK(err);
  memcpy(interp,b,sizeof(double)*(*Q)*(*P));
  memcpy(deriv,d,sizeof(double)*(*Q)*(*P)*3);
  err = dEFSRestoreExplicit((dEFS)vefs,jinv,Q,P,&b,&d);dCHK(err);
  return 0;
}

/* CPU-baseline kernel: the reference's per-element work of one matrix-free ellip MatMult, element by element, exactly
 * the call sequence of dRulesetIteratorGetPatchApplied/CommitPatchApplied for (coords, field) links
 * (src/fs/interface/fsrulesetit.c:545-556,631-637) around the reference's pointwise Jacobian (restated inline from
 * src/fs/tests/ellip.c:369-377 because it is a static function of an example program).  Gather/scatter through the
 * boolean E use the index array e2l (fs.c:481-521 with E boolean).
 *   cefs: coordinate EFS (P=2,Q), efs: field EFS; coords: [nelem][8][3]; stash: [nelem][Q][5] (eta, sqrt_mdeta_Du[3], lambda_exp_u)
 *   xloc/yloc: local-form vectors (Dirichlet entries of xloc already zeroed, fs.c:562-571).
 * Threads own disjoint element ranges and private copies of y (summed at the end) so no atomics are needed; with
 * nthreads==1 this is the faithful single-process reference loop.
 */
typedef struct {dEFS cefs,efs; dRule rule; const int *e2l; const double *coords,*stash,*xloc; double *yloc; int np,nq,nloc,nthreads; double **ypriv;} mm_ctx;
static void mm_body(int lo,int hi,void *vc,int *errout)
{
  mm_ctx *c = vc;
  const int np = c->np,nq = c->nq;
  double *xq,*dxq,*cjinv,*jw,*ue,*u,*du,*v,*dv,*ye,*y = c->yloc;
  if (dMallocA5(nq*3,&xq,nq*9,&dxq,nq*9,&cjinv,nq,&jw,np,&ue) || dMallocA5(nq,&u,nq*3,&du,nq,&v,nq*3,&dv,np,&ye)) {*errout |= 1; return;}
  for (int e=lo; e<hi; e++) {
    dErr err = 0;
    const int *idx = c->e2l + (long)e*np;
    for (int i=0; i<np; i++) ue[i] = c->xloc[idx[i]];
    err |= dEFSApply(c->cefs,NULL,3,c->coords+(long)e*24,xq,dAPPLY_INTERP,INSERT_VALUES);
    err |= dEFSApply(c->cefs,NULL,3,c->coords+(long)e*24,dxq,dAPPLY_GRAD,INSERT_VALUES);
    err |= dRuleComputePhysical(c->rule,dxq,cjinv,jw);
    err |= dEFSApply(c->efs,cjinv,1,ue,u,dAPPLY_INTERP,INSERT_VALUES);
    err |= dEFSApply(c->efs,cjinv,1,ue,du,dAPPLY_GRAD,INSERT_VALUES);
    for (int q=0; q<nq; q++) {
      const double *st = c->stash + ((long)e*nq+q)*5,*Du = du+3*q,w = jw[q];
      const double dotw = (st[1]*Du[0] + st[2]*Du[1] + st[3]*Du[2])*w,etaw = st[0]*w;
      v[q] = -w*st[4]*u[q];
      for (int i=0; i<3; i++) dv[3*q+i] = etaw*Du[i] - dotw*st[1+i];
    }
    err |= dEFSApply(c->efs,cjinv,1,v,ye,dAPPLY_INTERP_TRANSPOSE,INSERT_VALUES);
    err |= dEFSApply(c->efs,cjinv,1,dv,ye,dAPPLY_GRAD_TRANSPOSE,ADD_VALUES);
    for (int i=0; i<np; i++) y[idx[i]] += ye[i];
    *errout |= err;
  }
  dFree5(xq,dxq,cjinv,jw,ue);
  dFree5(u,du,v,dv,ye);
}
static pthread_mutex_t mm_lock = PTHREAD_MUTEX_INITIALIZER;
static void mm_body_locked(int lo,int hi,void *vc,int *errout)
{
  mm_ctx *c = vc;
  if (c->nthreads > 1 && lo > 0) {
    /* allocate the private y under a lock, then run */
    double *y = calloc((size_t)c->nloc,sizeof(double));
    pthread_mutex_lock(&mm_lock);
    for (int t=0; t<c->nthreads; t++) if (!c->ypriv[t]) {c->ypriv[t] = y; break;}
    pthread_mutex_unlock(&mm_lock);
    mm_ctx priv = *c;
    priv.yloc = y; priv.nthreads = 1;
    mm_body(lo,hi,&priv,errout);
  } else {
    mm_ctx priv = *c;
    priv.nthreads = 1;
    mm_body(lo,hi,&priv,errout);
  }
}
int RefEllipMatMult(void *vcefs,void *vefs,void *vrule,int nelem,int nloc,const int *e2l,const double *coords,const double *stash,
                    const double *xloc,double *yloc,int nthreads)
{
  int P[3],Q[3],err;
  RefEFSGetPQ(vefs,P,Q);
  if (nthreads < 1) nthreads = 1;
  double *ypriv[nthreads];
  memset(ypriv,0,sizeof(ypriv));
  mm_ctx c = {(dEFS)vcefs,(dEFS)vefs,(dRule)vrule,e2l,coords,stash,xloc,yloc,P[0]*P[1]*P[2],Q[0]*Q[1]*Q[2],nloc,nthreads,ypriv};
  err = par_for(nthreads,nelem,mm_body_locked,&c);
  for (int t=0; t<nthreads; t++) if (ypriv[t]) {
    for (int i=0; i<nloc; i++) yloc[i] += ypriv[t][i];
    free(ypriv[t]);
  }
  return err;
}

int RefMaxThreads(void) { return (int)sysconf(_SC_NPROCESSORS_ONLN); }
