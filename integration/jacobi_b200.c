/* Reference-side binding of the dJacobi seam: a dJacobi type "b200" for jedbrown/dohp whose element
 * function spaces apply through libdohp_b200.so (dB200EFSApplyBatch) instead of the CPU sum factorisation.
 *
 * Written against the REFERENCE'S OWN headers (include/dohpjacimpl.h, src/jacobi/impls/tensor/tensorimpl.h); a maintainer would
 * drop it at src/jacobi/impls/b200/b200.c and add one line to dJacobiRegisterAll (src/jacobi/interface/jacobi.c:182-183):
 *     err = dJacobiRegisterDynamic(dJACOBI_B200,path,"dJacobiCreate_B200",dJacobiCreate_B200);dCHK(err);
 * (here a constructor registers the type, because the reference tree is read-only).  In this repo it is compiled together with the
 * reference's unmodified sources by oracle/Makefile (`make refbind`) and driven by the reference's unmodified test program
 * src/jacobi/tests/ex1.c with -djac_type b200 (tests/test_gpu_zz_refbinding.py): the golden jacobi-1*.refout must come out.
 *
 * Design: "b200" IS the tensor type (same 1-D tables, node counts, constraint matrices: dJacobiCreate_Tensor runs first) with
 * dJacobiGetEFS wrapped: every dEFS the tensor type hands out is shadowed by a dEFS_B200 = the dEFS_Tensor (copied, so that every
 * other op -- getSizes, getTensorNodes, getExplicitSparse, ... -- works unchanged on it) + the library handles, with ops.apply
 * replaced.  Like upstream there is ONE shared object per (topology, degree, rule) (tensor.c:511-545); the GPU context is created
 * with the first dEFS.  Ownership: handles die with the dJacobi (dJacobiDestroy -> Destroy_B200).  Not thread-safe, as upstream. */
#include <tensorimpl.h>
#include "dohp_b200.h"

#define dJACOBI_B200 "b200"

typedef struct {
  dEFS_Tensor t;                /* MUST be first: the tensor ops read this object through a dEFS_Tensor* */
  dEFS        shadowed;         /* the tensor type's own object (cache key) */
  dB200Rule   rule;
  dB200EFS    h;
} dEFS_B200;

typedef struct s_B200Node {struct s_B200Node *next; dEFS_B200 *efs;} B200Node;
typedef struct s_B200Jac {
  struct s_B200Jac *next;
  dJacobi   jac;
  dB200Ctx  ctx;
  B200Node *efs;
  dErr (*GetEFS_Tensor)(dJacobi,dInt,const dEntTopology[],const dPolynomialOrder[],const dRule[],dEFS[]);
  dErr (*Destroy_Tensor)(dJacobi);
} B200Jac;
static B200Jac *b200_jacs;      /* jac->data belongs to the tensor type, so the b200 part is kept beside it */

static B200Jac *B200Find(dJacobi jac)
{
  for (B200Jac *b=b200_jacs; b; b=b->next) if (b->jac == jac) return b;
  return NULL;
}

/* dEFSApply (include/dohpjacobi.h:177; src/jacobi/interface/jacobi.c:489-506): same modes, InsertMode, layouts and error cases */
static dErr dEFSApply_B200(dEFS efs,const dReal jinv[],dInt D,const dScalar in[],dScalar out[],dApplyMode amode,InsertMode imode)
{
  dEFS_B200 *b = (dEFS_B200*)efs;

  dFunctionBegin;
  if (dB200EFSApplyBatch(b->h,1,jinv,D,in,out,(int)amode,(int)imode)) dERROR(PETSC_COMM_SELF,PETSC_ERR_LIB,"%s",dB200GetLastError());
  dFunctionReturn(0);
}

static dErr B200ShadowCreate(B200Jac *bj,dEFS shadowed,dEFS_B200 **out)
{
  const dEFS_Tensor *t = (const dEFS_Tensor*)shadowed;
  const dReal *rcoord[3],*rweight[3],*interp[3],*deriv[3];
  dInt rdim,rsize[3];
  int  P[3],Q[3],dim;
  dEFS_B200 *b;
  dErr err;

  dFunctionBegin;
  dim = t->topo == dTOPO_LINE ? 1 : t->topo == dTOPO_QUAD ? 2 : t->topo == dTOPO_HEX ? 3 : 0;
  if (!dim) dERROR(PETSC_COMM_SELF,PETSC_ERR_SUP,"no b200 dEFS for topology %s",dMeshEntTopologyName(t->topo));
  if (!bj->ctx && dB200Init(0,&bj->ctx)) dERROR(PETSC_COMM_SELF,PETSC_ERR_LIB,"%s",dB200GetLastError());
  err = dRuleGetTensorNodeWeight(t->rule,&rdim,rsize,rcoord,rweight);dCHK(err);
  if (rdim != dim) dERROR(PETSC_COMM_SELF,1,"Incompatible Rule dim %d, expected %d",rdim,dim);
  for (int d=0; d<dim; d++) {
    P[d] = t->basis[d]->P; Q[d] = t->basis[d]->Q;
    if (Q[d] != rsize[d]) dERROR(PETSC_COMM_SELF,PETSC_ERR_PLIB,"basis tabulated at %d points, rule has %d",Q[d],rsize[d]);
    interp[d] = t->basis[d]->interp; deriv[d] = t->basis[d]->deriv;      /* [Q][P] row-major, tensorimpl.h:17-24 */
  }
  err = dNew(dEFS_B200,&b);dCHK(err);
  b->t        = *t;                                                       /* ops table (by value), rule, topo, basis pointers */
  b->shadowed = shadowed;
  if ((dim == 3 ? dB200RuleCreate(bj->ctx,Q,rcoord,rweight,&b->rule) : dB200RuleCreateDim(bj->ctx,dim,Q,rcoord,rweight,&b->rule))
      || (dim == 3 ? dB200EFSCreate(bj->ctx,b->rule,P,interp,deriv,&b->h) : dB200EFSCreateDim(bj->ctx,b->rule,dim,P,interp,deriv,&b->h))) {
    dB200RuleDestroy(&b->rule);
    err = dFree(b);dCHK(err);
    dERROR(PETSC_COMM_SELF,PETSC_ERR_LIB,"%s",dB200GetLastError());
  }
  b->t.ops.apply = dEFSApply_B200;
  *out = b;
  dFunctionReturn(0);
}

static dErr dJacobiGetEFS_B200(dJacobi jac,dInt n,const dEntTopology topo[],const dPolynomialOrder order[],const dRule rules[],dEFS efs[])
{
  B200Jac *bj = B200Find(jac);
  dErr err;

  dFunctionBegin;
  if (!bj) dERROR(PETSC_COMM_SELF,PETSC_ERR_PLIB,"not a b200 dJacobi");
  err = bj->GetEFS_Tensor(jac,n,topo,order,rules,efs);dCHK(err);         /* tables, hashing and sharing are the tensor type's */
  for (dInt i=0; i<n; i++) {
    B200Node *node;
    for (node=bj->efs; node; node=node->next) if (node->efs->shadowed == efs[i]) break;
    if (!node) {
      dEFS_B200 *shadow = NULL;
      err = B200ShadowCreate(bj,efs[i],&shadow);dCHK(err);                /* first: nothing to undo if the library refuses (no GPU, bad tables) */
      err = dNew(B200Node,&node);dCHK(err);
      node->efs  = shadow;
      node->next = bj->efs;
      bj->efs    = node;
    }
    efs[i] = (dEFS)node->efs;
  }
  dFunctionReturn(0);
}

static dErr dJacobiDestroy_B200(dJacobi jac)
{
  B200Jac *bj = B200Find(jac),**link;
  dErr err;

  dFunctionBegin;
  if (!bj) dERROR(PETSC_COMM_SELF,PETSC_ERR_PLIB,"not a b200 dJacobi");
  for (B200Node *node=bj->efs,*next; node; node=next) {
    next = node->next;
    dB200EFSDestroy(&node->efs->h);
    dB200RuleDestroy(&node->efs->rule);
    err = dFree(node->efs);dCHK(err);
    err = dFree(node);dCHK(err);
  }
  if (bj->ctx) dB200Destroy(&bj->ctx);
  for (link=&b200_jacs; *link; link=&(*link)->next) if (*link == bj) {*link = bj->next; break;}
  err = bj->Destroy_Tensor(jac);dCHK(err);
  err = dFree(bj);dCHK(err);
  dFunctionReturn(0);
}

dErr dJacobiCreate_B200(dJacobi jac)
{
  B200Jac *bj;
  dErr err;

  dFunctionBegin;
  err = dJacobiCreate_Tensor(jac);dCHK(err);                              /* src/jacobi/impls/tensor/tensor.c:578-609 */
  err = dNew(B200Jac,&bj);dCHK(err);
  bj->jac            = jac;
  bj->ctx            = NULL;
  bj->efs            = NULL;
  bj->GetEFS_Tensor  = jac->ops->GetEFS;
  bj->Destroy_Tensor = jac->ops->Destroy;
  jac->ops->GetEFS   = dJacobiGetEFS_B200;
  jac->ops->Destroy  = dJacobiDestroy_B200;
  bj->next  = b200_jacs;
  b200_jacs = bj;
  dFunctionReturn(0);
}

/* stands in for the line in dJacobiRegisterAll (the list is a plain linked list: registering before dInitialize is fine) */
__attribute__((constructor)) static void dJacobiRegister_B200(void)
{
  (void)dJacobiRegister(dJACOBI_B200,"","dJacobiCreate_B200",dJacobiCreate_B200);
}
