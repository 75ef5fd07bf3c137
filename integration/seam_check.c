/* Drop-in check of the dJacobi seam, written against the reference's public API only (include/dohpjacobi.h): the same program
 * creates a dJacobi of type "tensor" (the reference's CPU implementation, compiled unmodified) and one of type "b200"
 * (integration/jacobi_b200.c -> libdohp_b200.so), asks both for the same element function spaces and compares dEFSApply in every
 * mode the reference implements, INSERT and ADD, with and without jinv, on hexes (isotropic, anisotropic, P > Q, SPARSE rule), quads
 * and lines.  Exit status 0 iff every case agrees to 1e-12 relative (the repo's FP64 bar) and the error cases error on both sides.
 * Built by oracle/Makefile (`make refbind`) into oracle/_ref/seam_check_b200; run by tests/test_gpu_zz_refbinding.py on the GPU box. */
#include <dohpjacobi.h>
#include <dohpsys.h>
#include <dohp.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

static unsigned long long lcg = 88172645463325252ULL;
static double rnd(void) { lcg = lcg*6364136223846793005ULL + 1442695040888963407ULL; return (double)(lcg>>11) / 9007199254740992.0 - 0.5; }

typedef struct {dEntTopology topo; int dim; dInt b[3],r[3]; dQuadratureMethod qm;} Case;

static double reldiff(dInt n,const dScalar *a,const dScalar *b)
{
  double d = 0,m = 1e-300;
  for (dInt i=0; i<n; i++) { if (fabs(a[i]-b[i]) > d) d = fabs(a[i]-b[i]); if (fabs(a[i]) > m) m = fabs(a[i]); }
  return d/m;
}

static dErr GetOne(dJacobi jac,const Case *c,dRule *rule,dEFS *efs)
{
  dPolynomialOrder rdeg = dPolynomialOrderCreate(0,c->r[0],c->r[1],c->r[2]),bdeg = dPolynomialOrderCreate(0,c->b[0],c->b[1],c->b[2]);
  dQuadrature quad;
  dRule *rules;
  dEFS  *efss;
  dErr err;

  dFunctionBegin;
  err = dJacobiGetQuadrature(jac,c->qm,&quad);dCHK(err);
  err = dQuadratureGetRules(quad,1,&c->topo,&rdeg,&rules);dCHK(err);
  err = dJacobiGetEFS(jac,1,&c->topo,&bdeg,rules,&efss);dCHK(err);
  *rule = rules[0]; *efs = efss[0];
  err = dFree(rules);dCHK(err);
  err = dFree(efss);dCHK(err);
  dFunctionReturn(0);
}

int main(int argc,char *argv[])
{
  static const Case cases[] = {
    {dTOPO_HEX ,3,{1,1,1},{2,2,2},dQUADRATURE_METHOD_FAST},
    {dTOPO_HEX ,3,{3,3,3},{6,6,6},dQUADRATURE_METHOD_FAST},
    {dTOPO_HEX ,3,{4,4,4},{8,8,8},dQUADRATURE_METHOD_FAST},
    {dTOPO_HEX ,3,{7,7,7},{14,14,14},dQUADRATURE_METHOD_FAST},
    {dTOPO_HEX ,3,{10,10,10},{20,20,20},dQUADRATURE_METHOD_FAST},
    {dTOPO_HEX ,3,{2,4,3},{6,8,4},dQUADRATURE_METHOD_FAST},          /* anisotropic, one direction with P > Q */
    {dTOPO_HEX ,3,{5,5,5},{22,22,22},dQUADRATURE_METHOD_FAST},        /* Q = 12 > the largest fused kernel */
    {dTOPO_HEX ,3,{3,3,3},{6,6,6},dQUADRATURE_METHOD_SPARSE},
    {dTOPO_HEX ,3,{4,4,4},{8,8,8},dQUADRATURE_METHOD_SPARSE},
    {dTOPO_QUAD,2,{3,4,0},{6,8,0},dQUADRATURE_METHOD_FAST},
    {dTOPO_LINE,1,{5,0,0},{10,0,0},dQUADRATURE_METHOD_FAST},
  };
  const int ncases = (int)(sizeof cases/sizeof cases[0]);
  dJacobi jt,jb;
  double worst = 0;
  int nchecked = 0,nfail = 0;
  dErr err;

  err = dInitialize(&argc,&argv,0,"dJacobi seam check: type tensor vs type b200\n");dCHK(err);
  err = dJacobiCreate(PETSC_COMM_WORLD,&jt);dCHK(err);
  err = dJacobiSetType(jt,dJACOBI_TENSOR);dCHK(err);
  err = dJacobiCreate(PETSC_COMM_WORLD,&jb);dCHK(err);
  err = dJacobiSetType(jb,getenv("DOHP_SEAM_OTHER") ? getenv("DOHP_SEAM_OTHER") : "b200");dCHK(err);   /* "tensor": self-check of this driver on a CPU-only machine */
  for (int ic=0; ic<ncases; ic++) {
    const Case *c = &cases[ic];
    dRule rt,rb;
    dEFS et,eb;
    dInt dim,P,Q,Pb,Qb;
    err = GetOne(jt,c,&rt,&et);dCHK(err);
    err = GetOne(jb,c,&rb,&eb);dCHK(err);
    err = dEFSGetSizes(et,&dim,NULL,&P);dCHK(err);
    err = dRuleGetSize(rt,NULL,&Q);dCHK(err);
    err = dEFSGetSizes(eb,NULL,NULL,&Pb);dCHK(err);
    err = dRuleGetSize(rb,NULL,&Qb);dCHK(err);
    if (dim != c->dim || P != Pb || Q != Qb) { printf("case %d: sizes differ\n",ic); nfail++; continue; }
    for (dInt D=1; D<=3; D+=2) {
      const dInt nmax = (P > Q ? P : Q)*D*3;
      dScalar *in,*ot,*ob;
      dReal *jinv;
      err = dMallocA4(nmax,&in,nmax,&ot,nmax,&ob,Q*9,&jinv);dCHK(err);
      for (dInt q=0; q<Q; q++) for (int k=0; k<9; k++) jinv[q*9+k] = (k%4 == 0 ? 1.0 : 0.0) + 0.3*rnd();
      for (int usej=0; usej<2; usej++) {
        if (usej && dim < 3) continue;                                   /* upstream: "Not implemented" (efstopo.c:638,653) */
        for (int m=0; m<4; m++) {
          const dApplyMode amode = (dApplyMode)(dAPPLY_INTERP+m);
          const int fwd = amode == dAPPLY_INTERP || amode == dAPPLY_GRAD,grad = amode == dAPPLY_GRAD || amode == dAPPLY_GRAD_TRANSPOSE;
          const dInt nin = D*(fwd ? P : Q*(grad ? 3 : 1)),nout = D*(fwd ? Q*(grad ? 3 : 1) : P);
          if (amode == dAPPLY_GRAD_TRANSPOSE && !usej) continue;         /* error case, checked below */
          for (int add=0; add<2; add++) {
            double d;
            for (dInt i=0; i<nin; i++) in[i] = rnd();
            for (dInt i=0; i<nout; i++) ot[i] = ob[i] = rnd();           /* ADD accumulates onto this, INSERT must overwrite it */
            err = dEFSApply(et,usej ? jinv : NULL,D,in,ot,amode,add ? ADD_VALUES : INSERT_VALUES);dCHK(err);
            err = dEFSApply(eb,usej ? jinv : NULL,D,in,ob,amode,add ? ADD_VALUES : INSERT_VALUES);dCHK(err);
            if (amode == dAPPLY_GRAD && dim < 3) {                       /* directions >= dim: upstream leaves garbage, b200 writes zeros */
              for (dInt i=0; i<Q*D; i++) for (int e=dim; e<3; e++) ot[i*3+e] = ob[i*3+e] = 0;
            }
            d = reldiff(nout,ot,ob);
            nchecked++;
            if (d > worst) worst = d;
            if (!(d <= 1e-12)) { printf("case %d D=%d mode %d jinv %d add %d: rel diff %.3e\n",ic,(int)D,(int)amode,usej,add,d); nfail++; }
          }
        }
      }
      err = dFree4(in,ot,ob,jinv);dCHK(err);
    }
    printf("case %2d: %s basis (%d,%d,%d) rule order (%d,%d,%d) %s: P=%d Q=%d ok\n",ic,c->dim == 3 ? "hex " : c->dim == 2 ? "quad" : "line",
           (int)c->b[0],(int)c->b[1],(int)c->b[2],(int)c->r[0],(int)c->r[1],(int)c->r[2],c->qm == dQUADRATURE_METHOD_SPARSE ? "sparse" : "fast",(int)P,(int)Q);
  }
  {                                                                      /* error behaviour: both sides refuse, neither crashes */
    dRule r; dEFS e[2]; dScalar in[64*3],out[64*3*3];
    for (int i=0; i<64*3; i++) in[i] = 1;
    err = GetOne(jt,&cases[1],&r,&e[0]);dCHK(err);
    err = GetOne(jb,&cases[1],&r,&e[1]);dCHK(err);
    fprintf(stderr,"---- four error tracebacks are expected below (dofs < 1; GRAD_TRANSPOSE without jinv), two per implementation ----\n");
    for (int k=0; k<2; k++) {
      if (!dEFSApply(e[k],NULL,0,in,out,dAPPLY_INTERP,INSERT_VALUES)) { printf("%s accepted dofs = 0\n",k ? "b200" : "tensor"); nfail++; }
      if (!dEFSApply(e[k],NULL,1,in,out,dAPPLY_GRAD_TRANSPOSE,INSERT_VALUES)) { printf("%s accepted GRAD_TRANSPOSE without jinv\n",k ? "b200" : "tensor"); nfail++; }
    }
  }
  err = dJacobiDestroy(&jt);dCHK(err);
  err = dJacobiDestroy(&jb);dCHK(err);
  printf("seam check: %d applies compared, max rel diff %.2e, %d failures: %s\n",nchecked,worst,nfail,nfail ? "FAIL" : "PASS");
  err = dFinalize();dCHK(err);
  return nfail ? 1 : 0;
}
