// Host-side launch descriptors: runtime sizes in, template instantiation selected inside the per-N translation units
// (apply_inst.cu / ellip_inst.cu / ... are compiled once per quadrature size N with -DDOHP_N=<N>).
#pragma once
#include <cuda_runtime.h>

struct dB200Ctx_s;

namespace dohp {

struct Tables {                  // host pointers to the isotropic 1-D tables
  int NP, NQ;
  const double *B;               // interp [NQ][NP]
  const double *Dt;              // collocation derivative on the rule points [NQ][NQ]
  const double *w, *q;           // rule weights / points [NQ]
};

struct ApplyLaunch {
  Tables tab;
  int nelem, D, amode, add;
  const double *jinv, *in;
  double *out;
};

struct EllipLaunch {
  Tables tab;
  dB200Ctx_s *ctx;               // owner of the cached shared-memory thread maps
  int affine;                    // every element is a parallelepiped: geometry once per element
  int mode;                      // ELLIP_LINEAR / ELLIP_JACOBIAN / ELLIP_FUNCTION
  int nelem, n, nc;
  const int *elist;              // optional element list of this launch (count entries)
  int count;
  const int *e2l;
  const double *coords, *x, *xaux;
  double *y, *yaux;
  double *stash;
  const double *forcing;
  double pm2h, eps2, lambda;
};

struct VecLaunch {
  Tables tab, tabp;              // velocity / displacement tables; pressure tables (stokes)
  dB200Ctx_s *ctx;
  int affine;
  int kind, mode, flags;         // VEC_ELAST / VEC_STOKES ; VEC_FUNCTION / VEC_JACOBIAN ; USE_* | OUT_* | BLOCK_A
  int nelem, n_u, nc_u, n_p, nc_p;
  const int *elist;
  int count;
  const int *e2l_u, *e2l_p;
  const double *coords, *xu, *xuaux, *xp, *xpaux;
  double *yu, *yuaux, *yp, *ypaux;
  double *stash;
  const double *forcing;
  double c0, c1, c2, c3;
};

// return 0 on success, -1 if (NP,NQ) is not instantiated, else a cudaError_t
typedef int (*ApplyFn)(const ApplyLaunch &, cudaStream_t);
typedef int (*EllipFn)(const EllipLaunch &, cudaStream_t);
typedef int (*VecFn)(const VecLaunch &, cudaStream_t);

#define DOHP_NMIN 2
#define DOHP_NMAX 11
#define DOHP_DECL(N) \
  int launch_apply_##N(const ApplyLaunch &, cudaStream_t); \
  int launch_ellip_##N(const EllipLaunch &, cudaStream_t); \
  int launch_vec_##N(const VecLaunch &, cudaStream_t);
DOHP_DECL(2) DOHP_DECL(3) DOHP_DECL(4) DOHP_DECL(5) DOHP_DECL(6) DOHP_DECL(7) DOHP_DECL(8) DOHP_DECL(9) DOHP_DECL(10) DOHP_DECL(11)
#undef DOHP_DECL

}  // namespace dohp
struct dB200EFS_s;
namespace dohp {
// any (P[3], Q[3]) with explicit tables (apply_generic.cu); -2: does not fit the shared-memory staging
int launch_apply_generic(dB200EFS_s *efs, int nelem, int D, int amode, int add, const double *jinv, const double *in, double *out, cudaStream_t s);

template <class TabT> inline void fill_basis(TabT &t, const Tables &s, int np, int nq) { for (int i = 0; i < np * nq; i++) t.B[i] = s.B[i]; }
template <class TabT> inline void fill_rule(TabT &t, const Tables &s, int nq)
{
  for (int i = 0; i < nq * nq; i++) t.Dt[i] = s.Dt[i];
  for (int i = 0; i < nq; i++) { t.w[i] = s.w[i]; t.q[i] = s.q[i]; }
}

}  // namespace dohp
