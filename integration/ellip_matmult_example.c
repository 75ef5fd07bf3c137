/* The operator seam from plain C99 (no Python, no PETSc): what EllipB200Setup + EllipShellMatMult_B200 of INTEGRATION.md section 1 do,
 * on a genblock box built by the library's own setup helper.  One matrix-free product y = J x of the linear Poisson operator
 * (src/fs/tests/ellip.c:442-472 with the defaults p = 2, eps = 1, lambda = 0 of ellip.c:156-158) with host vectors.
 *
 *   gcc -std=c99 -Wall -Wextra -pedantic -Iinclude integration/ellip_matmult_example.c -Ldohp_b200 -ldohp_b200 -Wl,-rpath,$PWD/dohp_b200 -lm
 *
 * tests/test_abi.py compiles it exactly like that (the header must be valid, warning-free C99) and runs it: on a machine without a GPU
 * it prints the host-side setup line and then has to stop at dB200Init with the library's message -- there is no CPU path behind the ABI.
 * (The device part is the call sequence dohp_b200/lib.py drives in every GPU test; this file itself has only been run up to dB200Init.) */
#include "dohp_b200.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#define CHK(call) do { if ((call)) { fprintf(stderr,"%s:%d %s\n  -> %s\n",__FILE__,__LINE__,#call,dB200GetLastError()); return 1; } } while (0)

int main(int argc,char *argv[])
{
  const int ne = argc > 1 ? atoi(argv[1]) : 8,deg = argc > 2 ? atoi(argv[2]) : 4,P = deg+1;
  dB200Ctx ctx;
  dB200BoxMesh *mesh;
  dB200BoxSpec spec = {{0,0,0},0,{1,1,1},0,7,{-1,1,-1,1,-1,1},0.0,0.0,0};     /* Dirichlet on sets 100, 200, 300 = the whole boundary */
  dB200Rule rule;
  dB200EFS efs;
  dB200Space space;
  dB200Op op;
  dB200EllipParam prm = {2.0,1.0,0.0};
  double qcoord[32],qweight[32],interp[32*16],deriv[32*16],*x,*y,nrm = 0;      /* dB200Rule1D wants room for order+2 = 22 entries at degree 10 */
  const double *coord3[3],*weight3[3],*interp3[3],*deriv3[3],*gx[1];
  double *gy[1];
  int Q,Q3[3],P3[3];

  if (ne < 1 || deg < 1 || deg > 10) { fprintf(stderr,"usage: %s [elements per direction] [degree 1..10]\n",argv[0]); return 2; }
  /* host-side setup: needs no device */
  spec.ne[0] = spec.ne[1] = spec.ne[2] = ne; spec.p = deg;
  CHK(dB200BoxMeshCreate(&spec,&mesh));                                         /* genblock + dFSBuildSpace_Cont, natural ordering */
  CHK(dB200Rule1D(dB200_QMETHOD_FAST,dB200_GAUSS_LEGENDRE,2*deg,&Q,qcoord,qweight));      /* rule order 2*degree, fsruleset.c:32-37 */
  CHK(dB200Basis1D(P,Q,qcoord,interp,deriv,NULL,NULL));                         /* TensorBasisCreate, tensor.c:22-109 */
  for (int d=0; d<3; d++) { Q3[d] = Q; P3[d] = P; coord3[d] = qcoord; weight3[d] = qweight; interp3[d] = interp; deriv3[d] = deriv; }
  printf("host setup: %d elements, %d dofs (+ %d Dirichlet nodes), %d^3 quadrature points per element\n",mesh->nelem,mesh->n,mesh->nc-mesh->n,Q);
  fflush(stdout);
  /* everything below runs on the GPU */
  CHK(dB200Init(0,&ctx));
  CHK(dB200RuleCreate(ctx,Q3,coord3,weight3,&rule));
  CHK(dB200EFSCreate(ctx,rule,P3,interp3,deriv3,&efs));
  CHK(dB200SpaceCreate(ctx,mesh->nelem,P,1,mesh->n,mesh->nc,mesh->ngh,mesh->e2l,&space));
  CHK(dB200OpCreate(ctx,dB200_OP_ELLIP,&prm,rule,1,&space,&efs,mesh->coords,&op));
  x = malloc((size_t)mesh->n*sizeof *x); y = malloc((size_t)mesh->n*sizeof *y);
  if (!x || !y) return 3;
  for (int i=0; i<mesh->n; i++) x[i] = sin(0.37*i);                             /* x_i = sin(0.37 i) */
  gx[0] = x; gy[0] = y;
  CHK(dB200OpMult(op,gx,gy));                                                   /* = EllipShellMatMult: HOMOGENEOUS, y overwritten */
  for (int i=0; i<mesh->n; i++) nrm += y[i]*y[i];
  printf("ellip %d^3 elements, degree %d: %d dofs, |J x|_2 = %.12e\n",ne,deg,mesh->n,sqrt(nrm));
  free(x); free(y);
  CHK(dB200OpDestroy(&op)); CHK(dB200SpaceDestroy(&space)); CHK(dB200EFSDestroy(&efs)); CHK(dB200RuleDestroy(&rule));
  CHK(dB200BoxMeshDestroy(&mesh)); CHK(dB200Destroy(&ctx));
  return 0;
}
