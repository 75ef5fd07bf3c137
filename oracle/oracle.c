/* TEST INFRASTRUCTURE — CPU oracle for dohp-b200 (see oracle.h for scope and the rule about who may load it).
 *
 * Plain-C restatement of the reference's matrix-free hot path.  Each function cites the reference file:line whose
 * arithmetic (including operation ORDER, which fixes the rounding) it follows.  Parity: PINNED against oracle/_ref
 * and the reference goldens by tests/test_oracle_*.py.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>
#include <pthread.h>

#define ORCHK(e) do {int _e = (e); if (_e) {fprintf(stderr,"[oracle] error %d at %s:%d\n",_e,__FILE__,__LINE__); return _e;}} while (0)
#define ORERR(...) do {fprintf(stderr,"[oracle] " __VA_ARGS__); fputc('\n',stderr); return 1;} while (0)

static int or_nthreads = 1;
int or_num_threads(void) { return or_nthreads; }
void or_set_num_threads(int n) { or_nthreads = n > 0 ? n : 1; }

/* static-range parallel-for on pthreads (no libgomp in this image) */
typedef struct {int lo,hi,err; int (*body)(int,int,void*); void *ctx;} or_task;
static void *or_tramp(void *v) { or_task *t = v; t->err = t->body(t->lo,t->hi,t->ctx); return NULL; }
static int or_parfor(int n,int (*body)(int,int,void*),void *ctx)
{
  int nt = or_nthreads;
  if (nt > n) nt = n > 0 ? n : 1;
  if (nt <= 1) return body(0,n,ctx);
  pthread_t th[nt];
  or_task task[nt];
  int err = 0;
  for (int t=0; t<nt; t++) {
    task[t] = (or_task){(int)((long)n*t/nt),(int)((long)n*(t+1)/nt),0,body,ctx};
    if (t) pthread_create(&th[t],NULL,or_tramp,&task[t]);
  }
  or_tramp(&task[0]);
  for (int t=1; t<nt; t++) pthread_join(th[t],NULL);
  for (int t=0; t<nt; t++) err |= task[t].err;
  return err;
}

/* ================================================================================================================= */
/* Jacobi polynomials, Gauss and Gauss-Lobatto points: polylib.c (alpha,beta general in the recurrences; the tensor
 * basis always passes alpha=beta=0, tensor.c:560-561).                                                              */
/* ================================================================================================================= */

/* P^{a,b}_n and its derivative at z: three-term recurrence and derivative relation of polylib.c:794-878 */
static void jacobi_pd(double z,int n,double alpha,double beta,double *poly,double *polyd)
{
  const double one = 1.0,two = 2.0,apb = alpha + beta;
  double pn2 = one,pn1 = 0.5*(alpha - beta + (apb + two)*z),p = 0;
  if (n == 0) {if (poly) *poly = one; if (polyd) *polyd = 0.0; return;}
  if (n == 1) {if (poly) *poly = pn1; if (polyd) *polyd = 0.5*(apb + two); return;}
  for (int k=2; k<=n; k++) {
    double a1 =  two*k*(k + apb)*(two*k + apb - two);
    double a2 = (two*k + apb - one)*(alpha*alpha - beta*beta);
    double a3 = (two*k + apb - two)*(two*k + apb - one)*(two*k + apb);
    double a4 =  two*(k + alpha - one)*(k + beta - one)*(two*k + apb);
    a2 /= a1; a3 /= a1; a4 /= a1;
    p   = (a2 + a3*z)*pn1 - a4*pn2;
    pn2 = pn1;
    pn1 = p;
  }
  if (poly) *poly = p;
  if (polyd) {                    /* pn2 holds P_{n-1} here; formula valid for -1<z<1 */
    double a1 = n*(alpha - beta),a2 = n*(two*n + apb),a3 = two*(n + alpha)*(n + beta),a4 = two*n + apb;
    a1 /= a4; a2 /= a4; a3 /= a4;
    *polyd = ((a1 - a2*z)*p + a3*pn2)/(one - z*z);
  }
}

/* d/dz P^{a,b}_n = (a+b+n+1)/2 P^{a+1,b+1}_{n-1}, valid on the closed interval: polylib.c:893-904 */
static double jacobi_d(double z,int n,double alpha,double beta)
{
  double p;
  if (n == 0) return 0.0;
  jacobi_pd(z,n-1,alpha+1.0,beta+1.0,&p,NULL);
  return p*0.5*(alpha + beta + (double)n + 1.0);
}

/* Zeros of P^{a,b}_n.  The reference finds them as eigenvalues of the recurrence's tridiagonal matrix
 * (polylib.c:1008-1066, JacZeros+TriQL); polylib's own alternative (polylib.c:955-985, Jacobz) is Newton iteration
 * with deflation from Chebyshev guesses.  We use the latter with a final undeflated Newton polish, which converges to
 * the same correctly rounded roots to within an ulp or two (checked against _ref in tests/test_oracle_jacobi.py). */
static void jacobi_zeros(int n,double *z,double alpha,double beta)
{
  const double dth = M_PI/(2.0*(double)n);
  double rlast = 0.0;
  for (int k=0; k<n; k++) {
    double r = -cos((2.0*(double)k + 1.0)*dth);
    if (k) r = 0.5*(r + rlast);
    for (int it=0; it<100; it++) {
      double poly,pder,sum = 0.0,delr;
      jacobi_pd(r,n,alpha,beta,&poly,&pder);
      for (int i=0; i<k; i++) sum += 1.0/(r - z[i]);
      delr = -poly/(pder - sum*poly);
      r += delr;
      if (fabs(delr) < 1e-15) break;
    }
    for (int it=0; it<3; it++) {  /* polish on the undeflated polynomial */
      double poly,pder;
      jacobi_pd(r,n,alpha,beta,&poly,&pder);
      r -= poly/pder;
    }
    z[k] = r;
    rlast = r;
  }
  /* enforce the exact antisymmetry that the symmetric eigenproblem delivers for alpha==beta */
  if (alpha == beta) for (int k=0; k<n/2; k++) {double m = 0.5*(z[n-1-k] - z[k]); z[k] = -m; z[n-1-k] = m;}
  if (alpha == beta && (n & 1)) z[n/2] = 0.0;
}

static double gamma_int(int n) { double g = 1.0; for (int i=2; i<n; i++) g *= i; return g; } /* Gamma(n), polylib.c:918-947 */

int or_gauss_legendre(int np,double *z,double *w)            /* zwgj, polylib.c:182-196 with alpha=beta=0 */
{
  if (np < 1) ORERR("np=%d",np);
  jacobi_zeros(np,z,0.0,0.0);
  double fac = 2.0*gamma_int(np+1)*gamma_int(np+1);
  fac /= gamma_int(np+1)*gamma_int(np+1);
  for (int i=0; i<np; i++) {
    double pd = jacobi_d(z[i],np,0.0,0.0);
    w[i] = fac/(pd*pd*(1.0 - z[i]*z[i]));
  }
  return 0;
}

int or_gauss_lobatto(int np,double *z,double *w)             /* zwglj, polylib.c:272-295 with alpha=beta=0 */
{
  if (np < 1) ORERR("np=%d",np);
  if (np == 1) {z[0] = 0.0; w[0] = 2.0; return 0;}
  z[0] = -1.0; z[np-1] = 1.0;
  if (np > 2) jacobi_zeros(np-2,z+1,1.0,1.0);
  double fac = 2.0*gamma_int(np)*gamma_int(np);
  fac /= (np-1)*gamma_int(np)*gamma_int(np+1);
  for (int i=0; i<np; i++) {
    double p;
    jacobi_pd(z[i],np-1,0.0,0.0,&p,NULL);
    w[i] = fac/(p*p);
  }
  /* w[0] *= (beta+1); w[np-1] *= (alpha+1): both factors are 1 */
  return 0;
}

/* i-th Lagrange interpolant through the np Gauss-Lobatto points, evaluated at z: hglj, polylib.c:632-651 */
static double lagrange_gll(int i,double z,const double *zglj,int np)
{
  const double zi = zglj[i],dz = z - zi;
  double p,pd,h;
  if (fabs(dz) < 1e-12) return 1.0;                                   /* EPS in polylib.c is 100*DBL_EPSILON; see note */
  jacobi_pd(zi,np-2,1.0,1.0,&p,NULL);
  pd = jacobi_d(zi,np-2,1.0,1.0);
  h = (1.0 - zi*zi)*pd - 2.0*zi*p;
  jacobi_pd(z,np-2,1.0,1.0,&p,NULL);
  return (1.0 - z*z)*p/(h*dz);
}

/* collocation derivative matrix on the Gauss-Lobatto points: Dglj, polylib.c:452-495 with alpha=beta=0 */
static void deriv_gll(double *D,const double *z,int np)
{
  double pd[np];
  pd[0] = 2.0*pow(-1.0,np)*gamma_int(np);
  pd[0] /= gamma_int(np-1)*gamma_int(2);
  for (int i=1; i<np-1; i++) pd[i] = jacobi_d(z[i],np-2,1.0,1.0)*(1.0 - z[i]*z[i]);
  pd[np-1] = -2.0*gamma_int(np);
  pd[np-1] /= gamma_int(np-1)*gamma_int(2);
  for (int i=0; i<np; i++) {
    for (int j=0; j<np; j++) {
      if (i != j) D[i*np+j] = pd[i]/(pd[j]*(z[i]-z[j]));
      else if (i == 0)    D[i*np+j] = (0.0 - (np-1)*(np + 0.0))/(2.0*2.0);
      else if (i == np-1) D[i*np+j] = -(0.0 - (np-1)*(np + 0.0))/(2.0*2.0);
      else                D[i*np+j] = (0.0*z[i])/(2.0*(1.0 - z[i]*z[i]));
    }
  }
}

/* TensorBasisCreate, tensor.c:22-72: nodes = GLL(P); interp = Lagrange at rule points; deriv = interp * Dglj */
int or_basis_1d(int P,int Q,const double *qcoord,double *interp,double *deriv,double *node,double *weight)
{
  if (P < 2 || P > 14) ORERR("TensorBasis size P=%d unsupported (tensor.c:75-96)",P);
  double cD[P*P];
  ORCHK(or_gauss_lobatto(P,node,weight));
  for (int i=0; i<Q; i++) for (int j=0; j<P; j++) interp[i*P+j] = lagrange_gll(j,qcoord[i],node,P);
  deriv_gll(cD,node,P);
  for (int i=0; i<Q; i++) {
    for (int j=0; j<P; j++) {
      double zz = 0;
      for (int k=0; k<P; k++) zz += interp[i*P+k]*cD[k*P+j];
      deriv[i*P+j] = zz;
    }
  }
  return 0;
}

/* TensorGetRule, tensorquad.c:51-118 (+ two_point_private :32-47) */
int or_rule_1d(int method,int family,int order,int *size,double *coord,double *weight)
{
  int n;
  if (!(0 <= order && order < 100)) ORERR("rule order %d out of bounds",order);
  switch (method) {
  case OR_QMETHOD_SPARSE: {
    n = order > 0 ? order : 1;
    if (n == 1) break;
    const int nelems = n/2;
    double intervals[nelems+1],unused[nelems+1];
    const double x0 = family == OR_GAUSS_LEGENDRE ? -0.57735026918962573 : -1.0,x1 = -x0;
    ORCHK(or_gauss_lobatto(nelems+1,intervals,unused));
    for (int i=0; i<nelems; i++) {
      double h = intervals[i+1] - intervals[i];
      coord[2*i+0] = intervals[i] + (x0 - (-1))*h/2.;
      coord[2*i+1] = intervals[i] + (x1 - (-1))*h/2.;
      weight[2*i+0] = h/2.;
      weight[2*i+1] = h/2.;
    }
    *size = n;
    return 0;
  }
  case OR_QMETHOD_FAST:
    n = family == OR_GAUSS_LEGENDRE ? 1 + order/2 : 2 + order/2;
    if (n > 1) {
      if (family == OR_GAUSS_LEGENDRE) ORCHK(or_gauss_legendre(n,coord,weight));
      else ORCHK(or_gauss_lobatto(n,coord,weight));
    }
    break;
  case OR_QMETHOD_SELF:
    if (family != OR_GAUSS_LOBATTO) ORERR("SELF needs the Lobatto family (tensorquad.c:94-101)");
    n = 1 + order/2;
    if (n > 1) ORCHK(or_gauss_lobatto(n,coord,weight));
    break;
  default: ORERR("quadrature method %d",method);
  }
  if (n == 1) {weight[0] = 2.0; coord[0] = 0.0;}
  *size = n;
  return 0;
}

/* ================================================================================================================= */
/* EFS objects                                                                                                        */
/* ================================================================================================================= */
typedef struct {int P,Q; double *interp,*deriv,*interpT,*derivT,*node,*weight;} OrBasis;
typedef struct {int size; double *coord,*weight;} OrRule1;
struct OrEFS_s {
  OrBasis b[3];
  OrRule1 r[3];
  int method;
  int npatches,patchsize,*patchind;
  double *patchweight;
};

int or_efs_destroy(OrEFS *pefs)
{
  OrEFS e = *pefs;
  if (!e) return 0;
  for (int d=0; d<3; d++) {
    free(e->b[d].interp); free(e->b[d].deriv); free(e->b[d].interpT); free(e->b[d].derivT); free(e->b[d].node); free(e->b[d].weight);
    free(e->r[d].coord); free(e->r[d].weight);
  }
  free(e->patchind); free(e->patchweight);
  free(e);
  *pefs = NULL;
  return 0;
}

int or_efs_create(const int bdeg[3],const int rorder[3],int method,int family,OrEFS *pefs)
{
  OrEFS e = calloc(1,sizeof(*e));
  int Q[3],N[3],Pp[3];
  e->method = method;
  for (int d=0; d<3; d++) {
    OrRule1 *r = &e->r[d];
    OrBasis *b = &e->b[d];
    r->coord = calloc(128,sizeof(double)); r->weight = calloc(128,sizeof(double));
    ORCHK(or_rule_1d(method,family,rorder[d],&r->size,r->coord,r->weight));
    const int P = bdeg[d]+1;            /* dJacobiGetEFS_Tensor: basis size = degree+1, tensor.c:521-530 */
    b->P = P; b->Q = Q[d] = r->size;
    b->interp = calloc(P*b->Q,sizeof(double)); b->deriv = calloc(P*b->Q,sizeof(double));
    b->interpT = calloc(P*b->Q,sizeof(double)); b->derivT = calloc(P*b->Q,sizeof(double));
    b->node = calloc(P,sizeof(double)); b->weight = calloc(P,sizeof(double));
    ORCHK(or_basis_1d(P,b->Q,r->coord,b->interp,b->deriv,b->node,b->weight));
    for (int i=0; i<b->Q; i++) for (int j=0; j<P; j++) {   /* explicit transposes, tensor.c:66-71 */
      b->interpT[j*b->Q+i] = b->interp[i*P+j];
      b->derivT[j*b->Q+i]  = b->deriv[i*P+j];
    }
  }
  /* dRulePatchSetup_Tensor, tensorquad.c:120-169 */
  if (method == OR_QMETHOD_SPARSE) {
    for (int d=0; d<3; d++) {
      Pp[d] = Q[d] > 1 ? 2 : 1;
      N[d] = Q[d]/Pp[d];
      if (N[d]*Pp[d] != Q[d]) ORERR("sparse quadrature needs an even number of points, got %d",Q[d]);
    }
  } else {
    for (int d=0; d<3; d++) {N[d] = 1; Pp[d] = Q[d];}
  }
  const int NN = N[0]*N[1]*N[2],PP = Pp[0]*Pp[1]*Pp[2];
  e->npatches = NN; e->patchsize = PP;
  e->patchind = calloc((size_t)NN*PP,sizeof(int));
  e->patchweight = calloc((size_t)NN*PP,sizeof(double));
  for (int i=0; i<N[0]; i++) for (int j=0; j<N[1]; j++) for (int k=0; k<N[2]; k++)
    for (int ii=0; ii<Pp[0]; ii++) for (int jj=0; jj<Pp[1]; jj++) for (int kk=0; kk<Pp[2]; kk++) {
      const int patch = (i*N[1]+j)*N[2]+k,patchidx = (ii*Pp[1]+jj)*Pp[2]+kk;
      const int ielem = i*Pp[0]+ii,jelem = j*Pp[1]+jj,kelem = k*Pp[2]+kk;
      e->patchind[patch*PP+patchidx] = (ielem*Q[1] + jelem)*Q[2] + kelem;
      /* reference quirk (iv): the z weight is indexed with jelem (tensorquad.c:161); replicated */
      e->patchweight[patch*PP+patchidx] = e->r[0].weight[ielem]*e->r[1].weight[jelem]*e->r[2].weight[jelem];
    }
  *pefs = e;
  return 0;
}

int or_efs_get_sizes(OrEFS e,int P[3],int Q[3])
{ for (int d=0; d<3; d++) {if (P) P[d] = e->b[d].P; if (Q) Q[d] = e->b[d].Q;} return 0; }
int or_efs_get_tables(OrEFS e,int dir,double *interp,double *deriv,double *node,double *weight)
{
  const OrBasis *b = &e->b[dir];
  if (interp) memcpy(interp,b->interp,sizeof(double)*b->P*b->Q);
  if (deriv)  memcpy(deriv,b->deriv,sizeof(double)*b->P*b->Q);
  if (node)   memcpy(node,b->node,sizeof(double)*b->P);
  if (weight) memcpy(weight,b->weight,sizeof(double)*b->P);
  return 0;
}
int or_efs_get_rule(OrEFS e,int dir,double *coord,double *weight)
{
  if (coord)  memcpy(coord,e->r[dir].coord,sizeof(double)*e->r[dir].size);
  if (weight) memcpy(weight,e->r[dir].weight,sizeof(double)*e->r[dir].size);
  return 0;
}
int or_efs_get_patches(OrEFS e,int *npatches,int *patchsize,const int **patchind,const double **patchweight)
{
  if (npatches) *npatches = e->npatches;
  if (patchsize) *patchsize = e->patchsize;
  if (patchind) *patchind = e->patchind;
  if (patchweight) *patchweight = e->patchweight;
  return 0;
}

/* ================================================================================================================= */
/* A3: TensorMult_Hex_nounroll, inlinetmulthex.h:19-100 — same loop nests, same accumulation order.                   */
/* ================================================================================================================= */
int or_tensor_mult_hex(int D,const int P[3],const int Q[3],const double *const A[3],const double *in,double *out,int imode)
{
  const int na = Q[0]*P[1]*P[2]*D,nb = Q[0]*Q[1]*P[2]*D;
  double amem[na],bmem[nb];
  const double *Ax = A[0],*Ay = A[1],*Az = A[2];
  memset(amem,0,sizeof(amem));
  memset(bmem,0,sizeof(bmem));
  if (imode == OR_INSERT_VALUES) memset(out,0,sizeof(double)*Q[0]*Q[1]*Q[2]*D);
  else if (imode != OR_ADD_VALUES) ORERR("InsertMode %d not supported",imode);
  {
    const int JKD = P[1]*P[2]*D;
    for (int l=0; l<Q[0]; l++) {
      double *aa = &amem[l*JKD];
      for (int i=0; i<P[0]; i++) {
        const double *ff = &in[i*JKD],aax = Ax[l*P[0]+i];
        for (int jkd=0; jkd<JKD; jkd++) aa[jkd] += aax*ff[jkd];
      }
    }
  }
  {
    const int JKD = P[1]*P[2]*D,KD = P[2]*D;
    for (int i=0; i<Q[0]; i++) {
      const double *aa = &amem[i*JKD];
      for (int l=0; l<Q[1]; l++) {
        double *bb = &bmem[(i*Q[1]+l)*KD];
        for (int j=0; j<P[1]; j++) {
          const double *aaa = &aa[j*KD],aay = Ay[l*P[1]+j];
          for (int kd=0; kd<KD; kd++) bb[kd] += aay*aaa[kd];
        }
      }
    }
  }
  if (D == 1) {
    for (int ij=0; ij<Q[0]*Q[1]; ij++) {
      const double *bb = &bmem[ij*P[2]];
      double *cc = &out[ij*Q[2]];
      for (int l=0; l<Q[2]; l++) {
        const double *aaz = &Az[l*P[2]];
        double s = 0;
        for (int k=0; k<P[2]; k++) s += aaz[k]*bb[k];
        cc[l] += s;
      }
    }
  } else {
    for (int ij=0; ij<Q[0]*Q[1]; ij++) {
      const double *bb = &bmem[ij*P[2]*D];
      for (int l=0; l<Q[2]; l++) {
        double *cc = &out[(ij*Q[2]+l)*D];
        const double *aa = &Az[l*P[2]];
        int kd = 0;
        for (int k=0; k<P[2]; k++) {
          const double aaz = aa[k];
          for (int d=0; d<D; d++,kd++) cc[d] += aaz*bb[kd];
        }
      }
    }
  }
  return 0;
}

/* A5: TensorRuleMapping / NoMapping, efstopo.c:518-591 */
static int rule_mapping(int Q,const double *jinv,int D,const double *in,double *out,int imode)
{
  const double (*u)[Q][D] = (const double(*)[Q][D])in;
  double (*v)[D][3] = (double(*)[D][3])out;
  if (imode != OR_INSERT_VALUES && imode != OR_ADD_VALUES) ORERR("InsertMode %d",imode);
  const int add = imode == OR_ADD_VALUES;
  if (!jinv) {
    for (int i=0; i<Q; i++) for (int d=0; d<D; d++) for (int e=0; e<3; e++) {
      if (add) v[i][d][e] += u[e][i][d]; else v[i][d][e] = u[e][i][d];
    }
    return 0;
  }
  const double (*J)[3][3] = (const double(*)[3][3])jinv;
  for (int i=0; i<Q; i++) for (int d=0; d<D; d++) for (int e=0; e<3; e++) {
    const double t = u[0][i][d]*J[i][0][e] + u[1][i][d]*J[i][1][e] + u[2][i][d]*J[i][2][e];
    if (add) v[i][d][e] += t; else v[i][d][e] = t;
  }
  return 0;
}

/* A6: TensorRuleMappingTranspose, efstopo.c:593-623 */
static int rule_mapping_transpose(int Q,const double *jinv,int D,const double *in,double *out,int imode)
{
  const double (*u)[D][3] = (const double(*)[D][3])in;
  double (*v)[Q][D] = (double(*)[Q][D])out;
  if (!jinv) ORERR("No Jinv, need a mapping");
  const double (*J)[3][3] = (const double(*)[3][3])jinv;
  const int add = imode == OR_ADD_VALUES;
  for (int i=0; i<Q; i++) for (int d=0; d<D; d++) for (int e=0; e<3; e++) {
    const double t = u[i][d][0]*J[i][e][0] + u[i][d][1]*J[i][e][1] + u[i][d][2]*J[i][e][2];
    if (add) v[e][i][d] += t; else v[e][i][d] = t;
  }
  return 0;
}

/* A4: dEFSApply_Tensor_Hex, efstopo.c:368-423 (dofs<1 is an error: jacobi.c:495) */
int or_efs_apply(OrEFS e,const double *jinv,int D,const double *in,double *out,int amode,int imode)
{
  int P[3],Q[3];
  const double *A[3];
  if (D < 1) ORERR("dofs=%d, must be at least 1",D);
  or_efs_get_sizes(e,P,Q);
  const int nq = Q[0]*Q[1]*Q[2];
  switch (amode) {
  case OR_APPLY_INTERP:
    A[0] = e->b[0].interp; A[1] = e->b[1].interp; A[2] = e->b[2].interp;
    return or_tensor_mult_hex(D,P,Q,A,in,out,imode);
  case OR_APPLY_INTERP_TRANSPOSE:
    A[0] = e->b[0].interpT; A[1] = e->b[1].interpT; A[2] = e->b[2].interpT;
    return or_tensor_mult_hex(D,Q,P,A,in,out,imode);
  case OR_APPLY_GRAD: {
    double *df = malloc(sizeof(double)*3*nq*D);
    A[0] = e->b[0].deriv; A[1] = e->b[1].interp; A[2] = e->b[2].interp;
    or_tensor_mult_hex(D,P,Q,A,in,df,OR_INSERT_VALUES);
    A[0] = e->b[0].interp; A[1] = e->b[1].deriv; A[2] = e->b[2].interp;
    or_tensor_mult_hex(D,P,Q,A,in,df+nq*D,OR_INSERT_VALUES);
    A[0] = e->b[0].interp; A[1] = e->b[1].interp; A[2] = e->b[2].deriv;
    or_tensor_mult_hex(D,P,Q,A,in,df+2*nq*D,OR_INSERT_VALUES);
    int err = rule_mapping(nq,jinv,D,df,out,imode);
    free(df);
    return err;
  }
  case OR_APPLY_GRAD_TRANSPOSE: {
    double *df = malloc(sizeof(double)*3*nq*D);
    int err = rule_mapping_transpose(nq,jinv,D,in,df,OR_INSERT_VALUES);
    if (err) {free(df); return err;}
    if (imode == OR_INSERT_VALUES) memset(out,0,sizeof(double)*P[0]*P[1]*P[2]*D);
    else if (imode != OR_ADD_VALUES) {free(df); ORERR("InsertMode %d invalid",imode);}
    A[0] = e->b[0].derivT; A[1] = e->b[1].interpT; A[2] = e->b[2].interpT;
    or_tensor_mult_hex(D,Q,P,A,df,out,OR_ADD_VALUES);
    A[0] = e->b[0].interpT; A[1] = e->b[1].derivT; A[2] = e->b[2].interpT;
    or_tensor_mult_hex(D,Q,P,A,df+nq*D,out,OR_ADD_VALUES);
    A[0] = e->b[0].interpT; A[1] = e->b[1].interpT; A[2] = e->b[2].derivT;
    or_tensor_mult_hex(D,Q,P,A,df+2*nq*D,out,OR_ADD_VALUES);
    free(df);
    return 0;
  }
  default: ORERR("invalid/unimplemented dApplyMode %d",amode);
  }
}

typedef struct {OrEFS e; const double *jinv,*in; double *out; int D,amode,imode,nq; long is,os;} apply_ctx;
static int apply_body(int lo,int hi,void *vc)
{
  apply_ctx *c = vc;
  int err = 0;
  for (int el=lo; el<hi; el++)
    err |= or_efs_apply(c->e,c->jinv ? c->jinv+(long)el*c->nq*9 : NULL,c->D,c->in+el*c->is,c->out+el*c->os,c->amode,c->imode);
  return err;
}
int or_efs_apply_batch(OrEFS e,int nelem,const double *jinv,int D,const double *in,double *out,int amode,int imode)
{
  int P[3],Q[3];
  or_efs_get_sizes(e,P,Q);
  const long np = P[0]*P[1]*P[2],nq = Q[0]*Q[1]*Q[2];
  apply_ctx c = {e,jinv,in,out,D,amode,imode,(int)nq,
                 (amode == OR_APPLY_INTERP || amode == OR_APPLY_GRAD) ? np*D : (amode == OR_APPLY_INTERP_TRANSPOSE ? nq*D : nq*D*3),
                 amode == OR_APPLY_INTERP ? nq*D : (amode == OR_APPLY_GRAD ? nq*D*3 : np*D)};
  return or_parfor(nelem,apply_body,&c);
}

/* dGeomInvert3, include/dohpgeom.h:57-75 */
static inline void invert3(const double a[9],double b[9],double *det)
{
  const double b0 =  (a[4]*a[8] - a[7]*a[5]);
  const double b3 = -(a[3]*a[8] - a[6]*a[5]);
  const double b6 =  (a[3]*a[7] - a[6]*a[4]);
  const double ldet = a[0]*b0 + a[1]*b3 + a[2]*b6;
  const double idet = 1.0/ldet;
  b[0] =  idet*b0;
  b[1] = -idet*(a[1]*a[8] - a[7]*a[2]);
  b[2] =  idet*(a[1]*a[5] - a[4]*a[2]);
  b[3] =  idet*b3;
  b[4] =  idet*(a[0]*a[8] - a[6]*a[2]);
  b[5] = -idet*(a[0]*a[5] - a[3]*a[2]);
  b[6] =  idet*b6;
  b[7] = -idet*(a[0]*a[7] - a[6]*a[1]);
  b[8] =  idet*(a[0]*a[4] - a[3]*a[1]);
  *det = ldet;
}

/* A7: dRuleComputePhysical, jacobi.c:383-414 with dRuleGetNodeWeight, jacobi.c:334-369 */
int or_rule_compute_physical(OrEFS e,const double *jac,double *jinv,double *jw)
{
  const int Q0 = e->r[0].size,Q1 = e->r[1].size,Q2 = e->r[2].size,n = Q0*Q1*Q2;
  for (int i=0; i<Q0; i++) for (int j=0; j<Q1; j++) for (int k=0; k<Q2; k++)
    jw[(i*Q1+j)*Q2+k] = e->r[0].weight[i]*e->r[1].weight[j]*e->r[2].weight[k];
  if (jac) {
    for (int i=0; i<n; i++) {
      double det;
      invert3(jac+i*9,jinv+i*9,&det);
      if (det < 0.) ORERR("Negative Jacobian at node %d",i);
      jw[i] *= det;
    }
  } else {
    for (int i=0; i<n; i++) for (int j=0; j<9; j++) jinv[9*i+j] = (j % 4 == 0) ? 1. : 0.;
  }
  return 0;
}
int or_rule_compute_physical_batch(OrEFS e,int nelem,const double *jac,double *jinv,double *jw)
{
  const long n = (long)e->r[0].size*e->r[1].size*e->r[2].size;
  for (int el=0; el<nelem; el++) ORCHK(or_rule_compute_physical(e,jac+el*n*9,jinv+el*n*9,jw+el*n));
  return 0;
}

/* Q1-b: dEFSGetExplicit, jacobi.c:508-559: apply INTERP/GRAD to each unit vector */
int or_efs_get_explicit(OrEFS e,const double *jinv,double *interp,double *deriv)
{
  int P3[3],Q3[3];
  or_efs_get_sizes(e,P3,Q3);
  const int P = P3[0]*P3[1]*P3[2],Q = Q3[0]*Q3[1]*Q3[2];
  double *v = calloc(P,sizeof(double)),*u = malloc(sizeof(double)*Q),*du = malloc(sizeof(double)*Q*3);
  for (int i=0; i<P; i++) {
    v[i] = 1;
    or_efs_apply(e,jinv,1,v,u,OR_APPLY_INTERP,OR_INSERT_VALUES);
    or_efs_apply(e,jinv,1,v,du,OR_APPLY_GRAD,OR_INSERT_VALUES);
    for (int j=0; j<Q; j++) {
      interp[j*P+i] = u[j];
      for (int k=0; k<3; k++) deriv[(j*P+i)*3+k] = du[j*3+k];
    }
    v[i] = 0;
  }
  free(v); free(u); free(du);
  return 0;
}

/* Q1-b: dEFSGetExplicitSparse_Basic, jacobi.c:592-643, including quirk (iii): cjinv is permuted into element order
 * with  cjinv_elem[qidx[i]] <- cjinv[i]  i.e. the incoming array is treated as patch-ordered. */
static int explicit_sparse_basic(OrEFS e,int npatches,int Q,const int *qidx,const double *cjinv,int eoffset,int *P,int *eidx,double *interp,double *deriv)
{
  int P3[3],Q3[3];
  or_efs_get_sizes(e,P3,Q3);
  const int PP = P3[0]*P3[1]*P3[2],QQ = Q3[0]*Q3[1]*Q3[2];
  if (npatches*Q != QQ) ORERR("patch specification does not cover element: %d*%d != %d",npatches,Q,QQ);
  double *cj = NULL,*ei = malloc(sizeof(double)*QQ*PP),*ed = malloc(sizeof(double)*QQ*PP*3);
  if (cjinv) {
    cj = malloc(sizeof(double)*QQ*9);
    /* jacobi.c:619-625, literally: ecjinv[i] = cjinv[qidx[i]] -- quirk (iii): the caller passes the ELEMENT-natural array (fsrulesetit.c:677-683),
     * so the gradient at element point i ends up mapped with the inverse Jacobian of point qidx[i]; identity on a single-patch (FAST) rule */
    for (int i=0; i<QQ; i++) for (int j=0; j<9; j++) cj[i*9+j] = cjinv[qidx[i]*9+j];
  }
  or_efs_get_explicit(e,cj,ei,ed);
  *P = PP;
  for (int i=0; i<npatches; i++) {
    for (int j=0; j<PP; j++) eidx[i*PP+j] = eoffset + j;
    for (int q=0; q<Q; q++) {
      const int src = qidx[i*Q+q];
      memcpy(&interp[(i*Q+q)*PP],&ei[src*PP],sizeof(double)*PP);
      memcpy(&deriv[(i*Q+q)*PP*3],&ed[src*PP*3],sizeof(double)*PP*3);
    }
  }
  free(cj); free(ei); free(ed);
  return 0;
}

/* Q1-a: dEFSGetExplicitSparse_Tensor_Hex, efstopo.c:660-717.  Quirk (i): deriv1 = +-1 regardless of the
 * sub-interval width; quirk (ii): cjinv is indexed patch-natural, (pidx*8+pqidx), whatever ordering was passed. */
int or_efs_get_explicit_sparse(OrEFS e,int npatches,int Q,const int *qidx,const double *cjinv,int eoffset,int *P,int *eidx,double *interp,double *deriv)
{
  const int px = e->b[0].P-1,py = e->b[1].P-1,pz = e->b[2].P-1,PP = 8;
  const double interp1[2][2] = {{0.78867513459481287,0.21132486540518713},{0.21132486540518713,0.78867513459481287}};
  const double deriv1[2][2] = {{-1.,1.},{-1.,1.}};
  static const double eye3[3][3] = {{1,0,0},{0,1,0},{0,0,1}};
  if (Q != 8 || px*py*pz != npatches) return explicit_sparse_basic(e,npatches,Q,qidx,cjinv,eoffset,P,eidx,interp,deriv);
  *P = PP;
  for (int i=0; i<px; i++) for (int j=0; j<py; j++) for (int k=0; k<pz; k++) {
    const int pidx = (i*py+j)*pz+k;
    for (int iii=0; iii<2; iii++) for (int jjj=0; jjj<2; jjj++) for (int kkk=0; kkk<2; kkk++)
      eidx[pidx*PP+(iii*2+jjj)*2+kkk] = eoffset + ((i+iii)*(py+1) + j+jjj)*(pz+1) + k+kkk;
    for (int ii=0; ii<2; ii++) for (int jj=0; jj<2; jj++) for (int kk=0; kk<2; kk++) {
      const int pqidx = (ii*2+jj)*2+kk;
      const double (*dXdx)[3] = cjinv ? (const double(*)[3])&cjinv[(pidx*8+pqidx)*9] : eye3;
      if (qidx[pidx*Q+pqidx] != ((i*2+ii)*py*2 + (j*2+jj))*pz*2 + (k*2+kk)) ORERR("Unexpected qidx");
      for (int iii=0; iii<2; iii++) for (int jjj=0; jjj<2; jjj++) for (int kkk=0; kkk<2; kkk++) {
        const int bidx = (iii*2+jjj)*2+kkk,qboff = (pidx*Q + pqidx)*PP + bidx;
        const double dX =  deriv1[ii][iii]*interp1[jj][jjj]*interp1[kk][kkk],
                     dY = interp1[ii][iii]* deriv1[jj][jjj]*interp1[kk][kkk],
                     dZ = interp1[ii][iii]*interp1[jj][jjj]* deriv1[kk][kkk];
        interp[qboff] = interp1[ii][iii]*interp1[jj][jjj]*interp1[kk][kkk];
        deriv[qboff*3+0] = dX*dXdx[0][0] + dY*dXdx[1][0] + dZ*dXdx[2][0];
        deriv[qboff*3+1] = dX*dXdx[0][1] + dY*dXdx[1][1] + dZ*dXdx[2][1];
        deriv[qboff*3+2] = dX*dXdx[0][2] + dY*dXdx[1][2] + dZ*dXdx[2][2];
      }
    }
  }
  return 0;
}

/* ================================================================================================================= */
/* A11: global <-> local <-> expanded                                                                                 */
/* ================================================================================================================= */

/* dFSGlobalToExpanded up to the ghost update, fs.c:543-579: closure = [global | dirichlet]; HOMOGENEOUS zeroes the
 * Dirichlet part, INHOMOGENEOUS restores it from the cache.  Ghost entries (local[nc..nc+ngh)) belong to the caller's
 * halo exchange (VecGhostUpdate INSERT FORWARD, fs.c:578-579); they are zeroed here. */
int or_fs_global_to_local(const OrSpace *s,const double *g,const double *dcache,int hmode,double *local)
{
  const long bs = s->bs;
  memcpy(local,g,sizeof(double)*s->n*bs);
  if (hmode == OR_HOMOGENEOUS) memset(local+s->n*bs,0,sizeof(double)*(s->nc-s->n)*bs);
  else if (hmode == OR_INHOMOGENEOUS) {
    if (s->nc > s->n && !dcache) ORERR("inhomogeneous mode needs the Dirichlet cache");
    if (s->nc > s->n) memcpy(local+s->n*bs,dcache,sizeof(double)*(s->nc-s->n)*bs);
  } else ORERR("hmode %d unsupported",hmode);
  memset(local+s->nc*bs,0,sizeof(double)*s->ngh*bs);
  return 0;
}

/* dFSLocalToExpanded = MatMult(E), fs.c:502-521, E boolean with one entry per row (tensor.c:283-302) */
int or_fs_local_to_expanded(const OrSpace *s,int np,const double *local,double *x)
{
  const long N = (long)s->nelem*np;
  const int bs = s->bs;
  for (long i=0; i<N; i++) for (int d=0; d<bs; d++) x[i*bs+d] = local[(long)s->e2l[i]*bs+d];
  return 0;
}

/* dFSExpandedToLocal = MatMultTranspose(Add)(E), fs.c:481-500 */
int or_fs_expanded_to_local(const OrSpace *s,int np,const double *x,double *local,int imode)
{
  const long N = (long)s->nelem*np;
  const int bs = s->bs;
  if (imode == OR_INSERT_VALUES) memset(local,0,sizeof(double)*(s->nc+s->ngh)*bs);
  for (long i=0; i<N; i++) for (int d=0; d<bs; d++) local[(long)s->e2l[i]*bs+d] += x[i*bs+d];
  return 0;
}

/* ================================================================================================================= */
/* Traversal (A10): one "link" per function space, link 0 = coordinates.  For each element                           */
/*   coords: INTERP -> x, GRAD(NULL jinv) -> dx, dRuleComputePhysical -> cjinv, jw     (fsrulesetit.c:545-560)        */
/*   field : INTERP -> u, GRAD(cjinv) -> du                                               (fsrulesetit.c:549-552)        */
/*   user pointwise loop                                                                                              */
/*   field : INTERP_TRANSPOSE(v) INSERT then GRAD_TRANSPOSE(dv) ADD into the expanded y   (fsrulesetit.c:631-637)        */
/* then y_expanded is scattered into the local vector (fsrulesetit.c:309-322 -> fs.c:592-640).                        */
/* ================================================================================================================= */
typedef struct {
  double *x,*dx,*cjinv,*jw;
  int nq;
} ElemGeom;

static int geom_alloc(ElemGeom *g,int nq)
{
  g->nq = nq;
  g->x = malloc(sizeof(double)*nq*3); g->dx = malloc(sizeof(double)*nq*9);
  g->cjinv = malloc(sizeof(double)*nq*9); g->jw = malloc(sizeof(double)*nq);
  return !(g->x && g->dx && g->cjinv && g->jw);
}
static void geom_free(ElemGeom *g) { free(g->x); free(g->dx); free(g->cjinv); free(g->jw); }
static int geom_eval(OrEFS cefs,const double *ecoords,ElemGeom *g)
{
  ORCHK(or_efs_apply(cefs,NULL,3,ecoords,g->x,OR_APPLY_INTERP,OR_INSERT_VALUES));
  ORCHK(or_efs_apply(cefs,NULL,3,ecoords,g->dx,OR_APPLY_GRAD,OR_INSERT_VALUES));
  ORCHK(or_rule_compute_physical(cefs,g->dx,g->cjinv,g->jw));
  return 0;
}

static void gather_elem(const OrSpace *s,int np,int e,const double *loc,double *ue)
{
  const int bs = s->bs,*idx = s->e2l + (long)e*np;
  for (int i=0; i<np; i++) for (int d=0; d<bs; d++) ue[i*bs+d] = loc[(long)idx[i]*bs+d];
}
static pthread_mutex_t scatter_lock = PTHREAD_MUTEX_INITIALIZER;
static void scatter_elem(const OrSpace *s,int np,int e,const double *ye,double *loc)
{
  const int bs = s->bs,*idx = s->e2l + (long)e*np;
  if (or_nthreads > 1) pthread_mutex_lock(&scatter_lock);
  for (int i=0; i<np; i++) for (int d=0; d<bs; d++) loc[(long)idx[i]*bs+d] += ye[i*bs+d];
  if (or_nthreads > 1) pthread_mutex_unlock(&scatter_lock);
}

/* ---- ellip --------------------------------------------------------------------------------------------------------- */

/* u, grad u, Hessian of the manufactured solutions of ellip.c:50-117 */
static void ellip_exact_uh(const OrEllipParam *prm,const double xyz[3],double *u,double du[3],double H[6])
{
  const double a = prm->a,b = prm->b,c = prm->c,x = xyz[0],y = xyz[1],z = xyz[2];
  switch (prm->exact) {
  case 0: {                        /* u = cos(ax) exp(by) sin(cz), ellip.c:50-57 */
    const double ca = cos(a*x),sa = sin(a*x),eb = exp(b*y),sc = sin(c*z),cc = cos(c*z);
    *u = ca*eb*sc;
    du[0] = -a*sa*eb*sc; du[1] = ca*b*eb*sc; du[2] = ca*eb*c*cc;
    H[0] = -a*a*ca*eb*sc; H[1] = b*b*ca*eb*sc; H[2] = -c*c*ca*eb*sc;      /* xx yy zz */
    H[3] = -a*b*sa*eb*sc; H[4] = -a*c*sa*eb*cc; H[5] = b*c*ca*eb*cc;       /* xy xz yz */
  } break;
  case 1:                          /* ellip.c:64-71 */
    *u = a*x*x + b*y*y + c*(1-z*z);
    du[0] = 2*a*x; du[1] = 2*b*y; du[2] = -2*c*z;
    H[0] = 2*a; H[1] = 2*b; H[2] = -2*c; H[3] = H[4] = H[5] = 0;
    break;
  case 2:                          /* ellip.c:78-85 */
    *u = a*x + b*y + c*z;
    du[0] = a; du[1] = b; du[2] = c;
    for (int i=0; i<6; i++) H[i] = 0;
    break;
  case 3:                          /* ellip.c:91-98 */
    *u = a*x*x*x + b*y*y*z + c*(1-z*z)*x;
    du[0] = 3*a*x*x + c*(1-z*z); du[1] = 2*b*y*z; du[2] = b*y*y - 2*c*z*x;
    H[0] = 6*a*x; H[1] = 2*b*z; H[2] = -2*c*x; H[3] = 0; H[4] = -2*c*z; H[5] = 2*b*y;
    break;
  default: {                       /* 4: sin(pi x) sin(pi y) sin(pi z), ellip.c:104-111 */
    const double pi = M_PI,sx = sin(pi*x),sy = sin(pi*y),sz = sin(pi*z),cx = cos(pi*x),cy = cos(pi*y),cz = cos(pi*z);
    *u = sx*sy*sz;
    du[0] = pi*cx*sy*sz; du[1] = pi*sx*cy*sz; du[2] = pi*sx*sy*cz;
    H[0] = H[1] = H[2] = -pi*pi*sx*sy*sz;
    H[3] = pi*pi*cx*cy*sz; H[4] = pi*pi*cx*sy*cz; H[5] = pi*pi*sx*cy*cz;
  }
  }
}

void or_ellip_exact_solution(const OrEllipParam *prm,const double x[3],double u[1],double du[3])
{ double H[6]; ellip_exact_uh(prm,x,u,du,H); }

/* f = -div(eta grad u), eta = (eps^2+gamma)^((p-2)/2), gamma = |grad u|^2/2  (the sympy recipe in the comment at
 * ellip.c:42-49).  Expanded by the product rule instead of pasting generated code:
 *   f = -eta*lap(u) - ((p-2)/2) * eta/(eps^2+gamma) * (grad u . H grad u).
 * Exact 4 is "not actually an exact solution" upstream: its forcing is the Laplacian one whatever p (ellip.c:112-116). */
void or_ellip_exact_forcing(const OrEllipParam *prm,const double x[3],double f[1])
{
  double u,du[3],H[6];
  ellip_exact_uh(prm,x,&u,du,H);
  if (prm->exact == 4) {f[0] = 3*M_PI*M_PI*u; return;}
  const double gamma = 0.5*(du[0]*du[0] + du[1]*du[1] + du[2]*du[2]),espg = prm->epsilon*prm->epsilon + gamma;
  const double eta = pow(espg,(prm->p-2)/2);
  const double Hg0 = H[0]*du[0] + H[3]*du[1] + H[4]*du[2],Hg1 = H[3]*du[0] + H[1]*du[1] + H[5]*du[2],Hg2 = H[4]*du[0] + H[5]*du[1] + H[2]*du[2];
  f[0] = -eta*(H[0]+H[1]+H[2]) - ((prm->p-2)/2)*eta/espg*(du[0]*Hg0 + du[1]*Hg1 + du[2]*Hg2);
}

/* EllipPointwiseComputeStore, ellip.c:333-342 */
static inline void ellip_store(const OrEllipParam *prm,const double u[1],const double Du[3],double st[OR_ELLIP_STORE])
{
  const double p = prm->p;
  const double gamma = 0.5*(Du[0]*Du[0] + Du[1]*Du[1] + Du[2]*Du[2]);
  const double espg = prm->epsilon*prm->epsilon + gamma;
  st[0] = pow(espg,(p-2)/2);
  const double sqrt_mdeta = sqrt(-((p-2)/2)*st[0]/espg);
  for (int i=0; i<3; i++) st[1+i] = sqrt_mdeta*Du[i];
  st[4] = prm->lambda != 0 ? prm->lambda*exp(u[0]) : 0;
}
/* EllipPointwiseFunction, ellip.c:355-367 */
static inline void ellip_pw_function(const OrEllipParam *prm,const double x[3],double w,const double u[1],const double Du[3],double st[],double v[1],double Dv[3])
{
  double f[1];
  ellip_store(prm,u,Du,st);
  or_ellip_exact_forcing(prm,x,f);
  v[0] = -w*f[0];
  v[0] += -w*st[4];
  for (int i=0; i<3; i++) Dv[i] = w*st[0]*Du[i];
}
/* EllipPointwiseJacobian, ellip.c:369-377 */
static inline void ellip_pw_jacobian(const double st[],double w,const double u[1],const double Du[3],double v[1],double Dv[3])
{
  const double dotw = (st[1]*Du[0] + st[2]*Du[1] + st[3]*Du[2])*w;
  const double etaw = st[0]*w;
  v[0] = -w*st[4]*u[0];
  for (int i=0; i<3; i++) Dv[i] = etaw*Du[i] - dotw*st[1+i];
}

typedef struct {
  OrEFS cefs,efs; const OrSpace *s; const double *coords; const OrEllipParam *prm;
  const double *xloc; double *yloc; double *stash; const double *cstash; int function;
} ellip_ctx;

static int ellip_body(int lo,int hi,void *vc)
{
  ellip_ctx *c = vc;
  int P[3],Q[3];
  or_efs_get_sizes(c->efs,P,Q);
  const int np = P[0]*P[1]*P[2],nq = Q[0]*Q[1]*Q[2];
  ElemGeom g;
  if (geom_alloc(&g,nq)) return 1;
  double *ue = malloc(sizeof(double)*np),*ye = malloc(sizeof(double)*np);
  double *u = malloc(sizeof(double)*nq),*du = malloc(sizeof(double)*nq*3),*v = malloc(sizeof(double)*nq),*dv = malloc(sizeof(double)*nq*3);
  int err = 0;
  for (int e=lo; e<hi && !err; e++) {
    err |= geom_eval(c->cefs,c->coords+(long)e*24,&g);
    gather_elem(c->s,np,e,c->xloc,ue);
    err |= or_efs_apply(c->efs,g.cjinv,1,ue,u,OR_APPLY_INTERP,OR_INSERT_VALUES);
    err |= or_efs_apply(c->efs,g.cjinv,1,ue,du,OR_APPLY_GRAD,OR_INSERT_VALUES);
    for (int q=0; q<nq; q++) {
      if (c->function) ellip_pw_function(c->prm,&g.x[3*q],g.jw[q],&u[q],&du[3*q],&c->stash[((long)e*nq+q)*OR_ELLIP_STORE],&v[q],&dv[3*q]);
      else ellip_pw_jacobian(&c->cstash[((long)e*nq+q)*OR_ELLIP_STORE],g.jw[q],&u[q],&du[3*q],&v[q],&dv[3*q]);
    }
    err |= or_efs_apply(c->efs,g.cjinv,1,v,ye,OR_APPLY_INTERP_TRANSPOSE,OR_INSERT_VALUES);
    err |= or_efs_apply(c->efs,g.cjinv,1,dv,ye,OR_APPLY_GRAD_TRANSPOSE,OR_ADD_VALUES);
    scatter_elem(c->s,np,e,ye,c->yloc);
  }
  geom_free(&g); free(ue); free(ye); free(u); free(du); free(v); free(dv);
  return err;
}

int or_ellip_function(OrEFS cefs,OrEFS efs,const OrSpace *s,const double *coords,const OrEllipParam *prm,const double *xloc,double *yloc,double *stash)
{
  ellip_ctx c = {cefs,efs,s,coords,prm,xloc,yloc,stash,NULL,1};
  return or_parfor(s->nelem,ellip_body,&c);
}
int or_ellip_mult(OrEFS cefs,OrEFS efs,const OrSpace *s,const double *coords,const OrEllipParam *prm,const double *stash,const double *xloc,double *yloc)
{
  ellip_ctx c = {cefs,efs,s,coords,prm,xloc,yloc,NULL,stash,0};
  return or_parfor(s->nelem,ellip_body,&c);
}

/* ---- Q1 assembly plumbing ---------------------------------------------------------------------------------------------- */

/* dFSMatSetValuesBlockedExpanded, fs.c:826-930, for boolean E: expanded row/col -> local node -> global block row
 * (negative = Dirichlet = dropped, fsbuild.c:31-34); values ADDed into the CSR at the position of the column. */
static int csr_add_block(int bs,const int *rowptr,const int *colidx,double *vals,int grow,int gcol,const double *K,int ldk)
{
  for (int fi=0; fi<bs; fi++) {
    const int r = grow*bs+fi,lo = rowptr[r],hi = rowptr[r+1];
    for (int fj=0; fj<bs; fj++) {
      const int c = gcol*bs+fj;
      int a = lo,b = hi-1,pos = -1;
      while (a <= b) {int m = (a+b)/2; if (colidx[m] == c) {pos = m; break;} else if (colidx[m] < c) a = m+1; else b = m-1;}
      if (pos < 0) ORERR("CSR pattern has no entry (%d,%d)",r,c);
      vals[pos] += K[fi*ldk+fj];
    }
  }
  return 0;
}

/* State evaluation on the SPARSE rule and its patch-ordered view: GetPatchApplied with npatches>1,
 * fsrulesetit.c:530-570,875-901,964-980. */
int or_ellip_jacobian_q1(OrEFS cefs_s,OrEFS efs_s,const OrSpace *s,const double *coords,const OrEllipParam *prm,
                         const double *xloc,const int *l2g,const int *rowptr,const int *colidx,double *vals)
{
  int P3[3],Q3[3],npatches,Q,P = 0;
  const int *patchind;
  or_efs_get_sizes(efs_s,P3,Q3);
  or_efs_get_patches(efs_s,&npatches,&Q,&patchind,NULL);
  const int np = P3[0]*P3[1]*P3[2],nq = Q3[0]*Q3[1]*Q3[2];
  ElemGeom g;
  if (geom_alloc(&g,nq)) return 1;
  double *ue = malloc(sizeof(double)*np),*u = malloc(sizeof(double)*nq),*du = malloc(sizeof(double)*nq*3);
  const int maxP = (Q == 8 && (P3[0]-1)*(P3[1]-1)*(P3[2]-1) == npatches) ? 8 : np;
  int *eidx = malloc(sizeof(int)*(size_t)npatches*maxP);
  double *interp = malloc(sizeof(double)*(size_t)npatches*Q*maxP),*deriv = malloc(sizeof(double)*(size_t)npatches*Q*maxP*3);
  double *K = malloc(sizeof(double)*maxP*maxP);
  int err = 0;
  for (int e=0; e<s->nelem && !err; e++) {
    err |= geom_eval(cefs_s,coords+(long)e*24,&g);
    gather_elem(s,np,e,xloc,ue);
    err |= or_efs_apply(efs_s,g.cjinv,1,ue,u,OR_APPLY_INTERP,OR_INSERT_VALUES);
    err |= or_efs_apply(efs_s,g.cjinv,1,ue,du,OR_APPLY_GRAD,OR_INSERT_VALUES);
    /* link 1 receives cjinv_elem (element-natural) — fsrulesetit.c:677,683 — quirks (ii)/(iii) live in the callee */
    err |= or_efs_get_explicit_sparse(efs_s,npatches,Q,patchind,g.cjinv,e*np,&P,eidx,interp,deriv);
    for (int pt=0; pt<npatches && !err; pt++) {
      const double *pin = interp + (size_t)pt*Q*P,*pde = deriv + (size_t)pt*Q*P*3;
      const int *rowcol = eidx + (size_t)pt*P;
      memset(K,0,sizeof(double)*P*P);
      for (int q=0; q<Q; q++) {                    /* ellip.c:506-520 */
        const int eq = patchind[pt*Q+q];           /* ValueCacheExtract: patch-ordered view of element-ordered data */
        double st[OR_ELLIP_STORE];
        ellip_store(prm,&u[eq],&du[3*eq],st);
        for (int j=0; j<P; j++) {
          double v[1],dv[3];
          ellip_pw_jacobian(st,g.jw[eq],&pin[q*P+j],&pde[(q*P+j)*3],v,dv);
          for (int i=0; i<P; i++)
            K[i*P+j] += (pin[q*P+i]*v[0] + pde[(q*P+i)*3+0]*dv[0] + pde[(q*P+i)*3+1]*dv[1] + pde[(q*P+i)*3+2]*dv[2]);
        }
      }
      for (int i=0; i<P && !err; i++) {
        const int gi = l2g[s->e2l[rowcol[i]]];
        if (gi < 0) continue;
        for (int j=0; j<P && !err; j++) {
          const int gj = l2g[s->e2l[rowcol[j]]];
          if (gj < 0) continue;
          err |= csr_add_block(1,rowptr,colidx,vals,gi,gj,&K[i*P+j],P);
        }
      }
    }
  }
  geom_free(&g); free(ue); free(u); free(du); free(eidx); free(interp); free(deriv); free(K);
  return err;
}

/* dNormsUpdate / dNormsFinish, src/util/util.c:119-152: integral error norms accumulated point by point */
static void norms_update(double uerr[3],double gerr[3],double jw,int bs,const double *uu,const double *u,const double *duu,const double *du)
{
  for (int i=0; i<bs; i++) {
    if (uerr) {
      const double r = uu[i] - u[i];
      uerr[0] += fabs(r)*jw;
      uerr[1] += r*r*jw;
      if (fabs(r) > uerr[2]) uerr[2] = fabs(r);
    }
    if (gerr) {
      const double dr[3] = {duu[i*3+0] - du[i*3+0],duu[i*3+1] - du[i*3+1],duu[i*3+2] - du[i*3+2]};
      const double gr2 = dr[0]*dr[0] + dr[1]*dr[1] + dr[2]*dr[2],grabs = sqrt(gr2);
      gerr[0] += grabs*jw;
      gerr[1] += gr2*jw;
      if (grabs > gerr[2]) gerr[2] = grabs;
    }
  }
}

/* EllipErrorNorms, ellip.c:538-564: the "Pointwise solution error" / "Pointwise gradient error" lines of the example's output */
int or_ellip_error_norms(OrEFS cefs,OrEFS efs,const OrSpace *s,const double *coords,const OrEllipParam *prm,const double *xloc,double enorm[3],double gnorm[3])
{
  int P[3],Q[3];
  or_efs_get_sizes(efs,P,Q);
  const int np = P[0]*P[1]*P[2],nq = Q[0]*Q[1]*Q[2];
  ElemGeom g;
  if (geom_alloc(&g,nq)) return 1;
  double *ue = malloc(sizeof(double)*np),*u = malloc(sizeof(double)*nq),*du = malloc(sizeof(double)*nq*3);
  int err = 0;
  for (int i=0; i<3; i++) enorm[i] = gnorm[i] = 0;
  for (int e=0; e<s->nelem && !err; e++) {
    err |= geom_eval(cefs,coords+(long)e*24,&g);
    gather_elem(s,np,e,xloc,ue);
    err |= or_efs_apply(efs,g.cjinv,1,ue,u,OR_APPLY_INTERP,OR_INSERT_VALUES);
    err |= or_efs_apply(efs,g.cjinv,1,ue,du,OR_APPLY_GRAD,OR_INSERT_VALUES);
    for (int q=0; q<nq; q++) {
      double uu[1],duu[3];
      or_ellip_exact_solution(prm,&g.x[3*q],uu,duu);
      norms_update(enorm,gnorm,g.jw[q],1,uu,&u[q],duu,&du[3*q]);
    }
  }
  enorm[1] = sqrt(enorm[1]); gnorm[1] = sqrt(gnorm[1]);
  geom_free(&g); free(ue); free(u); free(du);
  return err;
}

/* ================================================================================================================= */
/* L2 projection: the dFS integration test src/fs/tests/ex1.c (three API styles of one residual, ex1.c:294-430)        */
/* ================================================================================================================= */
/* exact_0 / exact_1, ex1.c:33-51 */
void or_proj_exact(const OrProjParam *prm,const double x[3],double f[1],double df[3])
{
  const double a = prm->a,b = prm->b,c = prm->c;
  if (prm->exact == 0) {
    f[0] = sin(x[0])*cosh(b*x[1]) + cos(a*x[0])*exp(x[1]) + sinh(x[1])*tanh(c*x[2]);
    df[0] = cos(x[0])*cosh(b*x[1]) - a*sin(a*x[0])*exp(x[1]);
    df[1] = sin(x[0])*b*sinh(b*x[1]) + cos(a*x[0])*exp(x[1]) + cosh(x[1])*tanh(c*x[2]);
    df[2] = sinh(x[1]) * c*(1.0 - tanh(c*x[2])*tanh(c*x[2]));
  } else {
    f[0] = 10*x[0] + 1*x[1] + 0.1*x[2];
    df[0] = 10; df[1] = 1; df[2] = 0.1;
  }
}

typedef struct {
  OrEFS cefs,efs; const OrSpace *s; const double *coords; const OrProjParam *prm;
  const double *xloc; double *yloc;
} proj_ctx;

/* prm != NULL: ProjResidual, v = (u - f) jw (ex1.c:331-335, :419-424 with dv = 0); prm == NULL: its derivative v = u jw, the mass
 * operator that -snes_mf differences */
static int proj_body(int lo,int hi,void *vc)
{
  proj_ctx *c = vc;
  int P[3],Q[3];
  or_efs_get_sizes(c->efs,P,Q);
  const int np = P[0]*P[1]*P[2],nq = Q[0]*Q[1]*Q[2];
  ElemGeom g;
  if (geom_alloc(&g,nq)) return 1;
  double *ue = malloc(sizeof(double)*np),*ye = malloc(sizeof(double)*np),*u = malloc(sizeof(double)*nq),*v = malloc(sizeof(double)*nq);
  int err = 0;
  for (int e=lo; e<hi && !err; e++) {
    err |= geom_eval(c->cefs,c->coords+(long)e*24,&g);
    gather_elem(c->s,np,e,c->xloc,ue);
    err |= or_efs_apply(c->efs,g.cjinv,1,ue,u,OR_APPLY_INTERP,OR_INSERT_VALUES);
    for (int q=0; q<nq; q++) {
      double f[1] = {0},df[3];
      if (c->prm) or_proj_exact(c->prm,&g.x[3*q],f,df);
      v[q] = (u[q] - f[0]) * g.jw[q];
    }
    err |= or_efs_apply(c->efs,g.cjinv,1,v,ye,OR_APPLY_INTERP_TRANSPOSE,OR_INSERT_VALUES);
    scatter_elem(c->s,np,e,ye,c->yloc);
  }
  geom_free(&g); free(ue); free(ye); free(u); free(v);
  return err;
}
int or_proj_function(OrEFS cefs,OrEFS efs,const OrSpace *s,const double *coords,const OrProjParam *prm,const double *xloc,double *yloc)
{
  proj_ctx c = {cefs,efs,s,coords,prm,xloc,yloc};
  return or_parfor(s->nelem,proj_body,&c);
}
int or_proj_mult(OrEFS cefs,OrEFS efs,const OrSpace *s,const double *coords,const double *xloc,double *yloc)
{
  proj_ctx c = {cefs,efs,s,coords,NULL,xloc,yloc};
  return or_parfor(s->nelem,proj_body,&c);
}

/* ProjResidualNorms, ex1.c:484-532 (the gradient "norms" are of |grad r|^2, as upstream computes them) */
int or_proj_residual_norms(OrEFS cefs,OrEFS efs,const OrSpace *s,const double *coords,const OrProjParam *prm,const double *xloc,double res[3],double gres[3])
{
  int P[3],Q[3];
  or_efs_get_sizes(efs,P,Q);
  const int np = P[0]*P[1]*P[2],nq = Q[0]*Q[1]*Q[2];
  ElemGeom g;
  if (geom_alloc(&g,nq)) return 1;
  double *ue = malloc(sizeof(double)*np),*u = malloc(sizeof(double)*nq),*du = malloc(sizeof(double)*nq*3);
  int err = 0;
  for (int i=0; i<3; i++) res[i] = gres[i] = 0;
  for (int e=0; e<s->nelem && !err; e++) {
    err |= geom_eval(cefs,coords+(long)e*24,&g);
    gather_elem(s,np,e,xloc,ue);
    err |= or_efs_apply(efs,g.cjinv,1,ue,u,OR_APPLY_INTERP,OR_INSERT_VALUES);
    err |= or_efs_apply(efs,g.cjinv,1,ue,du,OR_APPLY_GRAD,OR_INSERT_VALUES);
    for (int q=0; q<nq; q++) {
      double f[1],df[3],gr[3];
      or_proj_exact(prm,&g.x[3*q],f,df);
      const double r = u[q] - f[0];
      for (int d=0; d<3; d++) gr[d] = du[3*q+d] - df[d];
      const double grsum = gr[0]*gr[0] + gr[1]*gr[1] + gr[2]*gr[2];
      res[0] += fabs(r) * g.jw[q];
      res[1] += r*r * g.jw[q];
      if (fabs(r) > res[2]) res[2] = fabs(r);
      gres[0] += grsum * g.jw[q];
      gres[1] += grsum*grsum * g.jw[q];
      if (grsum > gres[2]) gres[2] = grsum;
    }
  }
  res[1] = sqrt(res[1]);
  gres[1] = sqrt(gres[1]);
  geom_free(&g); free(ue); free(u); free(du);
  return err;
}

/* ProjJacobian1, ex1.c:432-482: the Q1 mass matrix on the patches of the SPARSE rule, K[i][j] += interp[q][i] jw[q] interp[q][j] */
int or_proj_jacobian_q1(OrEFS cefs_s,OrEFS efs_s,const OrSpace *s,const double *coords,const int *l2g,const int *rowptr,const int *colidx,double *vals)
{
  int P3[3],Q3[3],npatches,Q,P = 0;
  const int *patchind;
  or_efs_get_sizes(efs_s,P3,Q3);
  or_efs_get_patches(efs_s,&npatches,&Q,&patchind,NULL);
  const int np = P3[0]*P3[1]*P3[2],nq = Q3[0]*Q3[1]*Q3[2];
  ElemGeom g;
  if (geom_alloc(&g,nq)) return 1;
  const int maxP = (Q == 8 && (P3[0]-1)*(P3[1]-1)*(P3[2]-1) == npatches) ? 8 : np;
  int *eidx = malloc(sizeof(int)*(size_t)npatches*maxP);
  double *interp = malloc(sizeof(double)*(size_t)npatches*Q*maxP),*deriv = malloc(sizeof(double)*(size_t)npatches*Q*maxP*3);
  double *K = malloc(sizeof(double)*maxP*maxP);
  int err = 0;
  for (int e=0; e<s->nelem && !err; e++) {
    err |= geom_eval(cefs_s,coords+(long)e*24,&g);
    err |= or_efs_get_explicit_sparse(efs_s,npatches,Q,patchind,g.cjinv,e*np,&P,eidx,interp,deriv);
    for (int pt=0; pt<npatches && !err; pt++) {
      const double *pin = interp + (size_t)pt*Q*P;
      const int *rowcol = eidx + (size_t)pt*P;
      memset(K,0,sizeof(double)*P*P);
      for (int i=0; i<P; i++)
        for (int q=0; q<Q; q++)
          for (int j=0; j<P; j++) K[i*P+j] += pin[q*P+i] * g.jw[patchind[pt*Q+q]] * pin[q*P+j];
      for (int i=0; i<P && !err; i++) {
        const int gi = l2g[s->e2l[rowcol[i]]];
        if (gi < 0) continue;
        for (int j=0; j<P && !err; j++) {
          const int gj = l2g[s->e2l[rowcol[j]]];
          if (gj < 0) continue;
          err |= csr_add_block(1,rowptr,colidx,vals,gi,gj,&K[i*P+j],P);
        }
      }
    }
  }
  geom_free(&g); free(eidx); free(interp); free(deriv); free(K);
  return err;
}

/* ================================================================================================================= */
/* Elasticity: St-Venant-Kirchhoff, src/fs/tests/elast.c:220-407                                                      */
/* ================================================================================================================= */
#include "exact_gen.h"

void or_elast_exact_solution(const OrElastParam *prm,const double x[3],double u[3],double du[9])
{
  if (prm->exact == 0) gen_elast0_solution(prm->a,prm->b,prm->c,prm->scale,x,u,du);
  else {for (int i=0; i<3; i++) u[i] = 0; for (int i=0; i<9; i++) du[i] = 0;}
}
void or_elast_exact_forcing(const OrElastParam *prm,const double x[3],double f[3])
{
  if (prm->exact == 0) gen_elast0_forcing(prm->a,prm->b,prm->c,prm->scale,prm->lambda,prm->mu,x,f);
  else f[0] = f[1] = f[2] = 0;
}

static inline double dot_column(const double a[9],const double b[9],int i,int j)       /* DotColumn, elast.c:220-221 */
{ return a[0*3+i]*b[0*3+j] + a[1*3+i]*b[1*3+j] + a[2*3+i]*b[2*3+j]; }
static inline void mult_ge_sy3(double C[9],const double A[9],const double S[6],int add)  /* dTensorMult(Add)GESY3, dohp.h:83-102 */
{
  for (int i=0; i<3; i++) {
    const double c0 = A[i*3+0]*S[0] + A[i*3+1]*S[3] + A[i*3+2]*S[4];
    const double c1 = A[i*3+0]*S[3] + A[i*3+1]*S[1] + A[i*3+2]*S[5];
    const double c2 = A[i*3+0]*S[4] + A[i*3+1]*S[5] + A[i*3+2]*S[2];
    if (add) {C[i*3+0] += c0; C[i*3+1] += c1; C[i*3+2] += c2;}
    else     {C[i*3+0]  = c0; C[i*3+1]  = c1; C[i*3+2]  = c2;}
  }
}
static inline void elast_pk2(const OrElastParam *prm,const double F[9],double S[6])       /* ElastPiolaKirchoff2, elast.c:227-241 */
{
  double E[6];
  E[0] = 0.5*(dot_column(F,F,0,0) - 1);
  E[1] = 0.5*(dot_column(F,F,1,1) - 1);
  E[2] = 0.5*(dot_column(F,F,2,2) - 1);
  E[3] = 0.5*(dot_column(F,F,0,1));
  E[4] = 0.5*(dot_column(F,F,0,2));
  E[5] = 0.5*(dot_column(F,F,1,2));
  const double trace = E[0] + E[1] + E[2];
  for (int i=0; i<6; i++) S[i] = prm->lambda*trace*(i<3) + 2*prm->mu*E[i];
}
static inline void elast_pk2_jacobian(const OrElastParam *prm,const double F[9],const double dF[9],double dS[6])  /* elast.c:242-256 */
{
  double dE[6];
  dE[0] = 0.5*(dot_column(F,dF,0,0) + dot_column(dF,F,0,0));
  dE[1] = 0.5*(dot_column(F,dF,1,1) + dot_column(dF,F,1,1));
  dE[2] = 0.5*(dot_column(F,dF,2,2) + dot_column(dF,F,2,2));
  dE[3] = 0.5*(dot_column(F,dF,0,1) + dot_column(dF,F,0,1));
  dE[4] = 0.5*(dot_column(F,dF,0,2) + dot_column(dF,F,0,2));
  dE[5] = 0.5*(dot_column(F,dF,1,2) + dot_column(dF,F,1,2));
  const double dtrace = dE[0] + dE[1] + dE[2];
  for (int i=0; i<6; i++) dS[i] = prm->lambda*dtrace*(i<3) + 2*prm->mu*dE[i];
}
static inline void elast_store(const double Du[9],double st[OR_ELAST_STORE])            /* DeformationGradient, elast.c:222-225 */
{ for (int i=0; i<3; i++) for (int j=0; j<3; j++) st[i*3+j] = (i==j) + Du[i*3+j]; }
static inline void elast_pw_function(const OrElastParam *prm,const double x[3],double w,const double Du[9],double st[],double v[3],double Dv[9])
{                                                                                         /* ElastPointwiseFunction, elast.c:257-268 */
  double f[3],S[6];
  or_elast_exact_forcing(prm,x,f);
  for (int i=0; i<3; i++) v[i] = -w*f[i];
  elast_store(Du,st);
  elast_pk2(prm,st,S);
  mult_ge_sy3(Dv,st,S,0);
  for (int i=0; i<9; i++) Dv[i] *= w;
}
static inline void elast_pw_jacobian(const OrElastParam *prm,const double st[],double w,const double Du[9],double v[3],double Dv[9])
{                                                                                         /* ElastPointwiseJacobian, elast.c:270-280 */
  double S[6],dS[6];
  v[0] = v[1] = v[2] = 0;
  elast_pk2(prm,st,S);
  elast_pk2_jacobian(prm,st,Du,dS);
  mult_ge_sy3(Dv,st,dS,0);
  mult_ge_sy3(Dv,Du,S,1);
  for (int i=0; i<9; i++) Dv[i] *= w;
}

typedef struct {
  OrEFS cefs,efs; const OrSpace *s; const double *coords; const OrElastParam *prm;
  const double *xloc; double *yloc; double *stash; const double *cstash; int function;
} elast_ctx;

static int elast_body(int lo,int hi,void *vc)
{
  elast_ctx *c = vc;
  int P[3],Q[3];
  or_efs_get_sizes(c->efs,P,Q);
  const int np = P[0]*P[1]*P[2],nq = Q[0]*Q[1]*Q[2];
  ElemGeom g;
  if (geom_alloc(&g,nq)) return 1;
  double *ue = malloc(sizeof(double)*np*3),*ye = malloc(sizeof(double)*np*3);
  double *u = malloc(sizeof(double)*nq*3),*du = malloc(sizeof(double)*nq*9),*v = malloc(sizeof(double)*nq*3),*dv = malloc(sizeof(double)*nq*9);
  int err = 0;
  for (int e=lo; e<hi && !err; e++) {
    err |= geom_eval(c->cefs,c->coords+(long)e*24,&g);
    gather_elem(c->s,np,e,c->xloc,ue);
    err |= or_efs_apply(c->efs,g.cjinv,3,ue,u,OR_APPLY_INTERP,OR_INSERT_VALUES);
    err |= or_efs_apply(c->efs,g.cjinv,3,ue,du,OR_APPLY_GRAD,OR_INSERT_VALUES);
    for (int q=0; q<nq; q++) {
      if (c->function) elast_pw_function(c->prm,&g.x[3*q],g.jw[q],&du[9*q],&c->stash[((long)e*nq+q)*OR_ELAST_STORE],&v[3*q],&dv[9*q]);
      else elast_pw_jacobian(c->prm,&c->cstash[((long)e*nq+q)*OR_ELAST_STORE],g.jw[q],&du[9*q],&v[3*q],&dv[9*q]);
    }
    err |= or_efs_apply(c->efs,g.cjinv,3,v,ye,OR_APPLY_INTERP_TRANSPOSE,OR_INSERT_VALUES);
    err |= or_efs_apply(c->efs,g.cjinv,3,dv,ye,OR_APPLY_GRAD_TRANSPOSE,OR_ADD_VALUES);
    scatter_elem(c->s,np,e,ye,c->yloc);
  }
  geom_free(&g); free(ue); free(ye); free(u); free(du); free(v); free(dv);
  return err;
}
int or_elast_function(OrEFS cefs,OrEFS efs,const OrSpace *s,const double *coords,const OrElastParam *prm,const double *xloc,double *yloc,double *stash)
{
  elast_ctx c = {cefs,efs,s,coords,prm,xloc,yloc,stash,NULL,1};
  return or_parfor(s->nelem,elast_body,&c);
}
int or_elast_mult(OrEFS cefs,OrEFS efs,const OrSpace *s,const double *coords,const OrElastParam *prm,const double *stash,const double *xloc,double *yloc)
{
  elast_ctx c = {cefs,efs,s,coords,prm,xloc,yloc,NULL,stash,0};
  return or_parfor(s->nelem,elast_body,&c);
}

/* ElastJacobian, elast.c:343-407: K[i][fi][j][fj] from unit trial functions; rows/cols are node blocks of 3 */
/* ElastErrorNorms, elast.c:409-438 */
int or_elast_error_norms(OrEFS cefs,OrEFS efs,const OrSpace *s,const double *coords,const OrElastParam *prm,const double *xloc,double enorm[3],double gnorm[3])
{
  int P[3],Q[3];
  or_efs_get_sizes(efs,P,Q);
  const int np = P[0]*P[1]*P[2],nq = Q[0]*Q[1]*Q[2];
  ElemGeom g;
  if (geom_alloc(&g,nq)) return 1;
  double *ue = malloc(sizeof(double)*np*3),*u = malloc(sizeof(double)*nq*3),*du = malloc(sizeof(double)*nq*9);
  int err = 0;
  for (int i=0; i<3; i++) enorm[i] = gnorm[i] = 0;
  for (int e=0; e<s->nelem && !err; e++) {
    err |= geom_eval(cefs,coords+(long)e*24,&g);
    gather_elem(s,np,e,xloc,ue);
    err |= or_efs_apply(efs,g.cjinv,3,ue,u,OR_APPLY_INTERP,OR_INSERT_VALUES);
    err |= or_efs_apply(efs,g.cjinv,3,ue,du,OR_APPLY_GRAD,OR_INSERT_VALUES);
    for (int q=0; q<nq; q++) {
      double uu[3],duu[9];
      or_elast_exact_solution(prm,&g.x[3*q],uu,duu);
      norms_update(enorm,gnorm,g.jw[q],3,uu,&u[3*q],duu,&du[9*q]);
    }
  }
  enorm[1] = sqrt(enorm[1]); gnorm[1] = sqrt(gnorm[1]);
  geom_free(&g); free(ue); free(u); free(du);
  return err;
}

int or_elast_jacobian_q1(OrEFS cefs_s,OrEFS efs_s,const OrSpace *s,const double *coords,const OrElastParam *prm,
                         const double *xloc,const int *l2g,const int *rowptr,const int *colidx,double *vals)
{
  int P3[3],Q3[3],npatches,Q,P = 0;
  const int *patchind;
  or_efs_get_sizes(efs_s,P3,Q3);
  or_efs_get_patches(efs_s,&npatches,&Q,&patchind,NULL);
  const int np = P3[0]*P3[1]*P3[2],nq = Q3[0]*Q3[1]*Q3[2];
  ElemGeom g;
  if (geom_alloc(&g,nq)) return 1;
  double *ue = malloc(sizeof(double)*np*3),*du = malloc(sizeof(double)*nq*9);
  const int maxP = (Q == 8 && (P3[0]-1)*(P3[1]-1)*(P3[2]-1) == npatches) ? 8 : np;
  int *eidx = malloc(sizeof(int)*(size_t)npatches*maxP);
  double *interp = malloc(sizeof(double)*(size_t)npatches*Q*maxP),*deriv = malloc(sizeof(double)*(size_t)npatches*Q*maxP*3);
  double *K = malloc(sizeof(double)*maxP*3*maxP*3);
  int err = 0;
  for (int e=0; e<s->nelem && !err; e++) {
    err |= geom_eval(cefs_s,coords+(long)e*24,&g);
    gather_elem(s,np,e,xloc,ue);
    err |= or_efs_apply(efs_s,g.cjinv,3,ue,du,OR_APPLY_GRAD,OR_INSERT_VALUES);
    err |= or_efs_get_explicit_sparse(efs_s,npatches,Q,patchind,g.cjinv,e*np,&P,eidx,interp,deriv);
    for (int pt=0; pt<npatches && !err; pt++) {
      const double *pin = interp + (size_t)pt*Q*P,*pde = deriv + (size_t)pt*Q*P*3;
      const int *rowcol = eidx + (size_t)pt*P;
      const int ld = P*3;
      memset(K,0,sizeof(double)*ld*ld);
      for (int q=0; q<Q; q++) {
        const int eq = patchind[pt*Q+q];
        double st[OR_ELAST_STORE];
        elast_store(&du[9*eq],st);
        for (int j=0; j<P; j++) for (int fj=0; fj<3; fj++) {
          double duu[9] = {0},v[3],dv[9];
          for (int d=0; d<3; d++) duu[fj*3+d] = pde[(q*P+j)*3+d];
          elast_pw_jacobian(prm,st,g.jw[eq],duu,v,dv);
          for (int i=0; i<P; i++) for (int fi=0; fi<3; fi++)
            K[(i*3+fi)*ld + j*3+fj] += pin[q*P+i]*v[fi] + pde[(q*P+i)*3+0]*dv[fi*3+0] + pde[(q*P+i)*3+1]*dv[fi*3+1] + pde[(q*P+i)*3+2]*dv[fi*3+2];
        }
      }
      for (int i=0; i<P && !err; i++) {
        const int gi = l2g[s->e2l[rowcol[i]]];
        if (gi < 0) continue;
        for (int j=0; j<P && !err; j++) {
          const int gj = l2g[s->e2l[rowcol[j]]];
          if (gj < 0) continue;
          err |= csr_add_block(3,rowptr,colidx,vals,gi,gj,&K[(i*3)*ld + j*3],ld);
        }
      }
    }
  }
  geom_free(&g); free(ue); free(du); free(eidx); free(interp); free(deriv); free(K);
  return err;
}

/* ================================================================================================================= */
/* Stokes: power-law, src/fs/tests/stokes.c:447-673                                                                   */
/* ================================================================================================================= */
void or_stokes_exact_solution(const OrStokesParam *prm,const double x[3],double u[3],double du[9],double p[1],double dp[3])
{
  if (prm->exact == 0) gen_stokes0_solution(prm->a,prm->b,prm->c,prm->scale,x,u,du,p,dp);
  else if (prm->exact == 1) gen_stokes1_solution(prm->a,prm->b,prm->c,prm->scale,x,u,du,p,dp);
  else {for (int i=0; i<3; i++) u[i] = dp[i] = 0; for (int i=0; i<9; i++) du[i] = 0; p[0] = 0;}
}
void or_stokes_exact_forcing(const OrStokesParam *prm,const double x[3],double fu[3],double fp[1])
{
  if (prm->exact == 0) gen_stokes0_forcing(prm->a,prm->b,prm->c,prm->scale,prm->B,prm->eps,prm->p,x,fu,fp);
  else if (prm->exact == 1) gen_stokes1_forcing(prm->a,prm->b,prm->c,prm->scale,prm->B,prm->eps,prm->p,x,fu,fp);
  else fu[0] = fu[1] = fu[2] = fp[0] = 0;
}
static inline void sym_compress3(const double Du[9],double Dv[6])     /* dTensorSymCompress3, dohp.h:79-80 */
{Dv[0] = Du[0]; Dv[1] = Du[4]; Dv[2] = Du[8]; Dv[3] = (Du[1]+Du[3])/2; Dv[4] = (Du[2]+Du[6])/2; Dv[5] = (Du[5]+Du[7])/2;}
static inline void sym_uncompress3(const double Du[6],double Dv[9])   /* dTensorSymUncompress3, dohp.h:81-82 */
{Dv[0] = Du[0]; Dv[4] = Du[1]; Dv[8] = Du[2]; Dv[1] = Dv[3] = Du[3]; Dv[2] = Dv[6] = Du[4]; Dv[5] = Dv[7] = Du[5];}
static inline double colon_sym3(const double a[6],const double b[6])  /* dColonSymScalar3, dohp.h:77-78 */
{ return a[0]*b[0] + a[1]*b[1] + a[2]*b[2] + 2*a[3]*b[3] + 2*a[4]*b[4] + 2*a[5]*b[5]; }
static inline void stokes_store(const OrStokesParam *r,const double Du[6],double st[OR_STOKES_STORE])   /* stokes.c:447-453 */
{
  const double gamma_reg = 0.5*r->eps*r->eps + 0.5*colon_sym3(Du,Du);
  st[0] = r->B * pow(gamma_reg,0.5*(r->p-2));
  st[1] = 0.5*(r->p-2) * st[0] / gamma_reg;
  for (int i=0; i<6; i++) st[2+i] = Du[i];
}
static inline double stokes_lambda(const OrStokesParam *r,const double st[]) { return ((1 + 0*st[0]) * r->nu) / (1 - 2*r->nu); }   /* stokes.c:455-458 */
static inline void stokes_pw_function(const OrStokesParam *r,const double x[3],double w,const double Du[6],double p,double st[],double v[3],double Dv[6],double *q)
{                                                                       /* StokesPointwiseFunction, stokes.c:460-475 */
  double fu[3],fp;
  const double trDu = Du[0]+Du[1]+Du[2];
  stokes_store(r,Du,st);
  const double lambda = 0*stokes_lambda(r,st);
  or_stokes_exact_forcing(r,x,fu,&fp);
  for (int i=0; i<3; i++) v[i] = -w*fu[i];
  *q = -w*(Du[0]+Du[1]+Du[2] + fp);
  for (int i=0; i<3; i++) Dv[i] = w*(st[0]*Du[i] + lambda*trDu - p);
  for (int i=3; i<6; i++) Dv[i] = w*st[0]*Du[i];
}
static inline void stokes_pw_jacobian(const OrStokesParam *r,const double st[],double w,const double Du[6],double p,double Dv[6],double *q)
{                                                                       /* StokesPointwiseJacobian, stokes.c:477-489 */
  const double lambda = 0*stokes_lambda(r,st);
  const double deta_colon = st[1]*colon_sym3(&st[2],Du);
  for (int i=0; i<6; i++) Dv[i] = w*(st[0]*Du[i] + deta_colon*st[2+i] + lambda*(i<3)*(Du[0]+Du[1]+Du[2]) - p*(i<3));
  *q = -w*(Du[0]+Du[1]+Du[2]);
}
static inline void stokes_pw_jacobian_A(const OrStokesParam *r,const double st[],double w,const double Du[6],double Dv[6])
{                                                                       /* stokes.c:491-496 */
  const double deta_colon = st[1]*colon_sym3(&st[2],Du);
  const double lambda = stokes_lambda(r,st);
  for (int i=0; i<6; i++) Dv[i] = w*(st[0]*Du[i] + deta_colon*st[2+i] + lambda*(i<3)*(Du[0]+Du[1]+Du[2]));
}

/* StokesJacobianAssemble_Velocity, stokes.c:675-737.  K[i][fi][j][fj] += deriv_i . dv with dv = StokesPointwiseJacobian_A applied to the
 * symmetrised gradient of the unit trial function e_fj phi_j (stokes.c:700-719), store recomputed at the SPARSE points from the
 * current velocity (stokes.c:703-706).  Mdiag (stokes.c:721-728): Mentry_i = sum_q interp[q][i]^2 jw[q] is added to v[i][0..2] where
 * v is the patch's array of test-function coefficients AT THE QUADRATURE POINTS (dRulesetIteratorGetPatchApplied's `v`), i.e. the
 * basis-function index i is used as a point index; CommitPatchApplied(INSERT) of the last patch then applies INTERP_TRANSPOSE of the
 * high-order basis on the SPARSE points to the whole element (fsrulesetit.c:595-643) and Finish adds the expanded vector into Mdiag
 * (HOMOGENEOUS).  Restated literally. */
int or_stokes_jacobian_q1_velocity(OrEFS cefs_s,OrEFS efsu_s,const OrSpace *su,const double *coords,const OrStokesParam *prm,
                                   const double *xuloc,const int *l2g,const int *rowptr,const int *colidx,double *vals,double *mdiag_loc)
{
  int P3[3],Q3[3],npatches,Q,P = 0;
  const int *patchind;
  or_efs_get_sizes(efsu_s,P3,Q3);
  or_efs_get_patches(efsu_s,&npatches,&Q,&patchind,NULL);
  const int np = P3[0]*P3[1]*P3[2],nq = Q3[0]*Q3[1]*Q3[2];
  if (!(Q == 8 && (P3[0]-1)*(P3[1]-1)*(P3[2]-1) == npatches)) return 1;      /* SPARSE rule only */
  ElemGeom g;
  if (geom_alloc(&g,nq)) return 1;
  double *ue = malloc(sizeof(double)*np*3),*ye = malloc(sizeof(double)*np*3),*du = malloc(sizeof(double)*nq*9),*vq = malloc(sizeof(double)*nq*3);
  int *eidx = malloc(sizeof(int)*(size_t)npatches*8);
  double *interp = malloc(sizeof(double)*(size_t)npatches*Q*8),*deriv = malloc(sizeof(double)*(size_t)npatches*Q*8*3);
  double *K = malloc(sizeof(double)*24*24);
  int err = 0;
  for (int e=0; e<su->nelem && !err; e++) {
    err |= geom_eval(cefs_s,coords+(long)e*24,&g);
    gather_elem(su,np,e,xuloc,ue);
    err |= or_efs_apply(efsu_s,g.cjinv,3,ue,du,OR_APPLY_GRAD,OR_INSERT_VALUES);
    err |= or_efs_get_explicit_sparse(efsu_s,npatches,Q,patchind,g.cjinv,e*np,&P,eidx,interp,deriv);
    memset(vq,0,sizeof(double)*nq*3);                      /* ValueCacheReset at the start of the element, fsrulesetit.c:545 */
    for (int pt=0; pt<npatches && !err; pt++) {
      const double *pin = interp + (size_t)pt*Q*P,*pde = deriv + (size_t)pt*Q*P*3;
      const int *rowcol = eidx + (size_t)pt*P;
      const int ld = P*3;
      memset(K,0,sizeof(double)*ld*ld);
      for (int q=0; q<Q; q++) {
        const int eq = patchind[pt*Q+q];
        double st[OR_STOKES_STORE],Dusym[6];
        sym_compress3(&du[9*eq],Dusym);
        stokes_store(prm,Dusym,st);
        for (int j=0; j<P; j++) for (int fj=0; fj<3; fj++) {
          double duu[9] = {0},dv[9],Duusym[6],Dvsym[6];
          for (int d=0; d<3; d++) duu[fj*3+d] = pde[(q*P+j)*3+d];
          sym_compress3(duu,Duusym);
          stokes_pw_jacobian_A(prm,st,g.jw[eq],Duusym,Dvsym);
          sym_uncompress3(Dvsym,dv);
          for (int i=0; i<P; i++) for (int fi=0; fi<3; fi++)
            K[(i*3+fi)*ld + j*3+fj] += pde[(q*P+i)*3+0]*dv[fi*3+0] + pde[(q*P+i)*3+1]*dv[fi*3+1] + pde[(q*P+i)*3+2]*dv[fi*3+2];
        }
      }
      for (int i=0; i<P && !err; i++) {
        const int gi = l2g[su->e2l[rowcol[i]]];
        if (gi < 0) continue;
        for (int j=0; j<P && !err; j++) {
          const int gj = l2g[su->e2l[rowcol[j]]];
          if (gj < 0) continue;
          err |= csr_add_block(3,rowptr,colidx,vals,gi,gj,&K[(i*3)*ld + j*3],ld);
        }
      }
      for (int i=0; i<P; i++) {                            /* stokes.c:721-728 */
        double Mentry = 0;
        for (int q=0; q<Q; q++) Mentry += pin[q*P+i] * g.jw[patchind[pt*Q+q]] * pin[q*P+i];
        for (int f=0; f<3; f++) vq[patchind[pt*Q+i]*3+f] += Mentry;
      }
    }
    if (mdiag_loc && !err) {
      err |= or_efs_apply(efsu_s,g.cjinv,3,vq,ye,OR_APPLY_INTERP_TRANSPOSE,OR_INSERT_VALUES);
      scatter_elem(su,np,e,ye,mdiag_loc);
    }
  }
  geom_free(&g); free(ue); free(ye); free(du); free(vq); free(eidx); free(interp); free(deriv); free(K);
  return err;
}

/* StokesJacobianAssemble_Pressure, stokes.c:739-794: D = scaled mass matrix sum_q interp_i jw/eta interp_j, Daux = Neumann Laplacian
 * sum_q deriv_i . deriv_j jw, both on the pressure space with the explicit basis of dEFSGetExplicitSparse_Basic (the pressure basis on
 * the velocity's SPARSE rule does not line up with Q1 sub-elements: every patch carries all pressure basis functions, jacobi.c:611-643,
 * incl. quirk (iii)); eta recomputed from the velocity gradient at the SPARSE points.  Scalar CSR over pressure rows/columns. */
int or_stokes_jacobian_pressure(OrEFS cefs_s,OrEFS efsu_s,OrEFS efsp_s,const OrSpace *su,const OrSpace *sp,const double *coords,const OrStokesParam *prm,
                                const double *xuloc,const int *l2gp,const int *rowptr,const int *colidx,double *valsD,double *valsDaux)
{
  int P3[3],Pp3[3],Q3[3],npatches,Q,P = 0;
  const int *patchind;
  or_efs_get_sizes(efsu_s,P3,Q3);
  or_efs_get_sizes(efsp_s,Pp3,Q3);
  or_efs_get_patches(efsu_s,&npatches,&Q,&patchind,NULL);
  const int np = P3[0]*P3[1]*P3[2],npp = Pp3[0]*Pp3[1]*Pp3[2],nq = Q3[0]*Q3[1]*Q3[2];
  ElemGeom g;
  if (geom_alloc(&g,nq)) return 1;
  double *ue = malloc(sizeof(double)*np*3),*du = malloc(sizeof(double)*nq*9);
  int *eidx = malloc(sizeof(int)*(size_t)npatches*npp);
  double *interp = malloc(sizeof(double)*(size_t)npatches*Q*npp),*deriv = malloc(sizeof(double)*(size_t)npatches*Q*npp*3);
  double *K = malloc(sizeof(double)*npp*npp),*Ka = malloc(sizeof(double)*npp*npp);
  int err = 0;
  for (int e=0; e<su->nelem && !err; e++) {
    err |= geom_eval(cefs_s,coords+(long)e*24,&g);
    gather_elem(su,np,e,xuloc,ue);
    err |= or_efs_apply(efsu_s,g.cjinv,3,ue,du,OR_APPLY_GRAD,OR_INSERT_VALUES);
    err |= or_efs_get_explicit_sparse(efsp_s,npatches,Q,patchind,g.cjinv,e*npp,&P,eidx,interp,deriv);
    for (int pt=0; pt<npatches && !err; pt++) {
      const double *pin = interp + (size_t)pt*Q*P,*pde = deriv + (size_t)pt*Q*P*3;
      const int *rowcol = eidx + (size_t)pt*P;
      memset(K,0,sizeof(double)*P*P); memset(Ka,0,sizeof(double)*P*P);
      for (int q=0; q<Q; q++) {
        const int eq = patchind[pt*Q+q];
        double st[OR_STOKES_STORE],Dusym[6];
        sym_compress3(&du[9*eq],Dusym);
        stokes_store(prm,Dusym,st);
        for (int j=0; j<P; j++) for (int i=0; i<P; i++) {
          K[i*P+j] += pin[q*P+i] * g.jw[eq] * (1./st[0]) * pin[q*P+j];
          Ka[i*P+j] += pde[(q*P+i)*3+0]*g.jw[eq]*pde[(q*P+j)*3+0] + pde[(q*P+i)*3+1]*g.jw[eq]*pde[(q*P+j)*3+1] + pde[(q*P+i)*3+2]*g.jw[eq]*pde[(q*P+j)*3+2];
        }
      }
      for (int i=0; i<P && !err; i++) {
        const int gi = l2gp[sp->e2l[rowcol[i]]];
        if (gi < 0) continue;
        for (int j=0; j<P && !err; j++) {
          const int gj = l2gp[sp->e2l[rowcol[j]]];
          if (gj < 0) continue;
          if (valsD) err |= csr_add_block(1,rowptr,colidx,valsD,gi,gj,&K[i*P+j],P);
          if (valsDaux) err |= csr_add_block(1,rowptr,colidx,valsDaux,gi,gj,&Ka[i*P+j],P);
        }
      }
    }
  }
  geom_free(&g); free(ue); free(du); free(eidx); free(interp); free(deriv); free(K); free(Ka);
  return err;
}

typedef struct {
  OrEFS cefs,efsu,efsp; const OrSpace *su,*sp; const double *coords; const OrStokesParam *prm;
  const double *xu,*xp; double *yu,*yp; double *stash; const double *cstash; int function,block;
} stokes_ctx;

/* function: StokesFunction stokes.c:507-539.  block 0: MatMult_Nest_StokesCoupled stokes.c:541-577;
 * 1/2/3: StokesShellMatMult_All_IorA A / B / Bt stokes.c:579-673 */
static int stokes_body(int lo,int hi,void *vc)
{
  stokes_ctx *c = vc;
  int P[3],Q[3],Pp[3];
  or_efs_get_sizes(c->efsu,P,Q);
  or_efs_get_sizes(c->efsp,Pp,Q);
  const int np = P[0]*P[1]*P[2],npp = Pp[0]*Pp[1]*Pp[2],nq = Q[0]*Q[1]*Q[2];
  ElemGeom g;
  if (geom_alloc(&g,nq)) return 1;
  double *ue = malloc(sizeof(double)*np*3),*ye = malloc(sizeof(double)*np*3),*pe = malloc(sizeof(double)*npp),*qe = malloc(sizeof(double)*npp);
  double *du = calloc(nq*9,sizeof(double)),*v = calloc(nq*3,sizeof(double)),*dv = calloc(nq*9,sizeof(double)),*p = calloc(nq,sizeof(double)),*q = calloc(nq,sizeof(double));
  const int useu = c->function || c->block != 3,usep = c->function || c->block == 0 || c->block == 3;
  const int outu = c->function || c->block != 2,outp = c->function || c->block == 0 || c->block == 2;
  int err = 0;
  for (int e=lo; e<hi && !err; e++) {
    err |= geom_eval(c->cefs,c->coords+(long)e*24,&g);
    if (useu) {
      gather_elem(c->su,np,e,c->xu,ue);
      err |= or_efs_apply(c->efsu,g.cjinv,3,ue,du,OR_APPLY_GRAD,OR_INSERT_VALUES);
    }
    if (usep) {
      gather_elem(c->sp,npp,e,c->xp,pe);
      err |= or_efs_apply(c->efsp,g.cjinv,1,pe,p,OR_APPLY_INTERP,OR_INSERT_VALUES);
    }
    for (int i=0; i<nq; i++) {
      double Du[6],Dv[6] = {0};
      double *st = c->function ? &c->stash[((long)e*nq+i)*OR_STOKES_STORE] : NULL;
      const double *cst = c->function ? NULL : &c->cstash[((long)e*nq+i)*OR_STOKES_STORE];
      sym_compress3(&du[9*i],Du);
      if (c->function) stokes_pw_function(c->prm,&g.x[3*i],g.jw[i],Du,p[i],st,&v[3*i],Dv,&q[i]);
      else if (c->block == 0) stokes_pw_jacobian(c->prm,cst,g.jw[i],Du,p[i],Dv,&q[i]);
      else if (c->block == 1) stokes_pw_jacobian_A(c->prm,cst,g.jw[i],Du,Dv);
      else if (c->block == 2) q[i] = -g.jw[i]*(Du[0]+Du[1]+Du[2]);                               /* stokes.c:498-501 */
      else {for (int k=0; k<3; k++) Dv[k] = -g.jw[i]*p[i]; for (int k=3; k<6; k++) Dv[k] = 0;}  /* stokes.c:503-507 */
      sym_uncompress3(Dv,&dv[9*i]);
    }
    if (outu) {
      if (c->function) {
        err |= or_efs_apply(c->efsu,g.cjinv,3,v,ye,OR_APPLY_INTERP_TRANSPOSE,OR_INSERT_VALUES);
        err |= or_efs_apply(c->efsu,g.cjinv,3,dv,ye,OR_APPLY_GRAD_TRANSPOSE,OR_ADD_VALUES);
      } else err |= or_efs_apply(c->efsu,g.cjinv,3,dv,ye,OR_APPLY_GRAD_TRANSPOSE,OR_INSERT_VALUES);
      scatter_elem(c->su,np,e,ye,c->yu);
    }
    if (outp) {
      err |= or_efs_apply(c->efsp,g.cjinv,1,q,qe,OR_APPLY_INTERP_TRANSPOSE,OR_INSERT_VALUES);
      scatter_elem(c->sp,npp,e,qe,c->yp);
    }
  }
  geom_free(&g); free(ue); free(ye); free(pe); free(qe); free(du); free(v); free(dv); free(p); free(q);
  return err;
}
int or_stokes_function(OrEFS cefs,OrEFS efsu,OrEFS efsp,const OrSpace *su,const OrSpace *sp,const double *coords,const OrStokesParam *prm,
                       const double *xu,const double *xp,double *yu,double *yp,double *stash)
{
  stokes_ctx c = {cefs,efsu,efsp,su,sp,coords,prm,xu,xp,yu,yp,stash,NULL,1,0};
  return or_parfor(su->nelem,stokes_body,&c);
}
int or_stokes_mult(OrEFS cefs,OrEFS efsu,OrEFS efsp,const OrSpace *su,const OrSpace *sp,const double *coords,const OrStokesParam *prm,
                   const double *stash,int block,const double *xu,const double *xp,double *yu,double *yp)
{
  stokes_ctx c = {cefs,efsu,efsp,su,sp,coords,prm,xu,xp,yu,yp,NULL,stash,0,block};
  return or_parfor(su->nelem,stokes_body,&c);
}

/* StokesErrorNorms, stokes.c:835-885 (alldirichlet: the constant pressure mode is removed by a volume integral of the exact pressure first) */
int or_stokes_error_norms(OrEFS cefs,OrEFS efsu,OrEFS efsp,const OrSpace *su,const OrSpace *sp,const double *coords,const OrStokesParam *prm,
                          const double *xu,const double *xp,int alldirichlet,double enorm[3],double gnorm[3],double pnorm[3])
{
  int P[3],Q[3],Pp[3];
  or_efs_get_sizes(efsu,P,Q);
  or_efs_get_sizes(efsp,Pp,Q);
  const int np = P[0]*P[1]*P[2],npp = Pp[0]*Pp[1]*Pp[2],nq = Q[0]*Q[1]*Q[2];
  ElemGeom g;
  if (geom_alloc(&g,nq)) return 1;
  double *ue = malloc(sizeof(double)*np*3),*pe = malloc(sizeof(double)*npp);
  double *u = malloc(sizeof(double)*nq*3),*du = malloc(sizeof(double)*nq*9),*p = malloc(sizeof(double)*nq);
  double volume = 0,pressureshift = 0;
  int err = 0;
  for (int i=0; i<3; i++) enorm[i] = gnorm[i] = pnorm[i] = 0;
  if (alldirichlet) {
    for (int e=0; e<su->nelem && !err; e++) {
      err |= geom_eval(cefs,coords+(long)e*24,&g);
      gather_elem(sp,npp,e,xp,pe);
      err |= or_efs_apply(efsp,g.cjinv,1,pe,p,OR_APPLY_INTERP,OR_INSERT_VALUES);
      for (int i=0; i<nq; i++) {
        double uu[3],duu[9],pp[1],dpp[3];
        or_stokes_exact_solution(prm,&g.x[3*i],uu,duu,pp,dpp);
        volume += g.jw[i];
        pressureshift += (pp[0] - p[i]) * g.jw[i];
      }
    }
    pressureshift /= volume;
  }
  for (int e=0; e<su->nelem && !err; e++) {
    err |= geom_eval(cefs,coords+(long)e*24,&g);
    gather_elem(su,np,e,xu,ue);
    gather_elem(sp,npp,e,xp,pe);
    err |= or_efs_apply(efsu,g.cjinv,3,ue,u,OR_APPLY_INTERP,OR_INSERT_VALUES);
    err |= or_efs_apply(efsu,g.cjinv,3,ue,du,OR_APPLY_GRAD,OR_INSERT_VALUES);
    err |= or_efs_apply(efsp,g.cjinv,1,pe,p,OR_APPLY_INTERP,OR_INSERT_VALUES);
    for (int i=0; i<nq; i++) {
      double uu[3],duu[9],pp[1],dpp[3];
      or_stokes_exact_solution(prm,&g.x[3*i],uu,duu,pp,dpp);
      pp[0] -= pressureshift;
      norms_update(enorm,gnorm,g.jw[i],3,uu,&u[3*i],duu,&du[9*i]);
      norms_update(pnorm,NULL,g.jw[i],1,pp,&p[i],NULL,NULL);
    }
  }
  enorm[1] = sqrt(enorm[1]); gnorm[1] = sqrt(gnorm[1]); pnorm[1] = sqrt(pnorm[1]);
  geom_free(&g); free(ue); free(pe); free(u); free(du); free(p);
  return err;
}
