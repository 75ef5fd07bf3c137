// Fused scalar elliptic (p-Laplacian / Bratu) element kernel: the whole dRulesetIterator traversal of EllipFunction /
// EllipShellMatMult (src/fs/tests/ellip.c:407-472) for a batch of elements in one thread block:
//   E gather (dFSGlobalToExpanded, src/fs/interface/fs.c:543-586) -> quadrature geometry -> u, grad u at the points ->
//   pointwise physics (ellip.c:333-377) -> transposes -> E^T scatter-add (dFSExpandedToGlobal, fs.c:592-640).
// Nothing but x, the index array, 24 vertex coordinates per element, the stash and y touches HBM.
#pragma once
#include "tensor_stages.cuh"

namespace dohp {

enum { ELLIP_LINEAR = 0, ELLIP_JACOBIAN = 1, ELLIP_FUNCTION = 2 };
#define ELLIP_NSTASH 5      // struct EllipStore {eta, sqrt_mdeta_Du[3], lambda_exp_u}, ellip.c:119-123

template <int NP, int NQ> struct EllipArgs {
  BasisTab<NP, NQ> bt;
  RuleTab<NQ> rt;
  int nelem, n, nc;              // elements; owned explicit nodes (indices >= n live in the aux arrays); closure size
  const int *e2l;                // [nelem][NP^3]
  const double *coords;          // [nelem][8][3]
  const double *x, *xaux;        // global vector [n]; [dirichlet | ghost] values [nc+ngh-n]
  double *y, *yaux;
  double *stash;                 // SoA: [NSTASH][nelem][NQ (c)][NQ*NQ (a,b)]
  const double *forcing;         // [nelem][NQ (c)][NQ*NQ] or NULL
  double pm2h, eps2, lambda;     // (p-2)/2, epsilon^2, lambda
};

template <int NQ> struct EllipCfg {
  static constexpr int T = NQ * NQ;                                   // threads per element (one per z line)
  static constexpr int EPB = (T >= 128) ? 2 : (256 / T);               // elements per block
  static constexpr int THREADS = ((EPB * T + 31) / 32) * 32;
  static constexpr int SMEM_DOUBLES = EPB * (3 * Lay<NQ>::CS + 24);
#ifndef DOHP_MINB
  static constexpr int MINB = NQ <= 8 ? 3 : 2;                        // resident blocks per SM the register budget is capped for
#else
  static constexpr int MINB = DOHP_MINB;
#endif
  static constexpr int MINB_NL = 2;                                   // stash modes keep more values live per point (80 registers spill)
};

template <int NP, int NQ, int MODE>
__global__ void __launch_bounds__(EllipCfg<NQ>::THREADS, MODE == ELLIP_LINEAR ? EllipCfg<NQ>::MINB : EllipCfg<NQ>::MINB_NL) ellip_kernel(const __grid_constant__ EllipArgs<NP, NQ> P)
{
  using L = Lay<NQ>;
  using C = EllipCfg<NQ>;
  constexpr int T = C::T;
  extern __shared__ double smem[];
  const int le = threadIdx.x / T, t = threadIdx.x - le * T, t_ = t;
  const long e = (long)blockIdx.x * C::EPB + le;
  const bool live = le < C::EPB && e < P.nelem;
  double *A = smem + (live ? le : 0) * (3 * L::CS + 24), *Gx = A + L::CS, *Gy = Gx + L::CS, *X = Gy + L::CS;

  // ---- gather + x contraction -----------------------------------------------------------------------------------------
  int idx[NP];
  if (MODE == ELLIP_JACOBIAN && live && (t & 15) == 0) {
    // the stash planes are read only after the three contractions: pull their lines into L2 now (one lane per 128 B)
    const long splane = (long)P.nelem * NQ * T;
#pragma unroll
    for (int c = 0; c < NQ; c++)
#pragma unroll
      for (int k = 0; k < ELLIP_NSTASH; k++)
        if (k < 4 || P.lambda != 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.stash + k * splane + (e * NQ + c) * T + t));
  }
  if (live) {
    for (int s = t; s < 24; s += T) X[s] = P.coords[e * 24 + s];
    if (t < NP * NP) {
      const int j = t / NP, k = t - j * NP;
      double r[NP];
#pragma unroll
      for (int i = 0; i < NP; i++) idx[i] = __ldg(P.e2l + e * (NP * NP * NP) + i * NP * NP + t);
#pragma unroll
      for (int i = 0; i < NP; i++) r[i] = idx[i] < P.n ? __ldg(P.x + idx[i]) : __ldg(P.xaux + (idx[i] - P.n));
      fwd_x<NP, NQ>(P.bt.B, r, A, j, k);
    }
  }
  __syncthreads();
  if (live && t < NQ * NP) { const int a = t / NP, k = t - a * NP; fwd_y<NP, NQ>(P.bt.B, A, a, k); }
  __syncthreads();
  const int ta = t / NQ, tb = t - ta * NQ;          // (a,b) for z lines, (b,c) for x lines, (a,c) for y lines
  double u[NQ], gz[NQ];
  if (live) {
    fwd_z<NP, NQ>(P.bt.B, A, ta, tb, u);
    der_z<NQ>(P.rt.Dt, u, gz);
    put_z<NQ>(A, ta, tb, u);
  }
  __syncthreads();
  if (live) {
    der_x<NQ>(P.rt.Dt, A, Gx, ta, tb);
    der_y<NQ>(P.rt.Dt, A, Gy, ta, tb);
  }
  __syncthreads();

  // ---- pointwise ----------------------------------------------------------------------------------------------------------
  double h[NQ];
  if (live) {
    LineGeom lg;
    line_geom(X, P.rt.q[ta], P.rt.q[tb], lg, false);
    const double wab = P.rt.w[ta] * P.rt.w[tb];
    const bool affine = line_is_affine(lg);
    PointGeom pg;
    if (affine) point_geom_adj(lg, 0.0, pg);
#pragma unroll
    for (int c = 0; c < NQ; c++) {
      if (!affine) point_geom_adj(lg, P.rt.q[c], pg);
      const double *adj = pg.adj;
      const double g0 = Gx[L::at(ta, tb, c)], g1 = Gy[L::at(ta, tb, c)], g2 = gz[c];
      // TensorRuleMapping, efstopo.c:574, with jinv = adj/det: t = det * Du
      double t[3], Dv[3], v = 0;
#pragma unroll
      for (int d = 0; d < 3; d++) t[d] = g0 * adj[0 + d] + g1 * adj[3 + d] + g2 * adj[6 + d];
      const double wq = wab * P.rt.w[c];                                      // jw = wq * det (jacobi.c:383-414)
      const long sq = (e * NQ + c) * T + t_;                                  // point index in the stash planes
      const long splane = (long)P.nelem * NQ * T;
      double scale;                                                          // Dv_ref = scale * adj . Dv
      if (MODE == ELLIP_LINEAR) {                                            // p=2, lambda=0: eta=1, s=0 (ellip.c:369-377)
        scale = wq * pg.rdet;                                                // jw * (1/det)^2
#pragma unroll
        for (int d = 0; d < 3; d++) Dv[d] = t[d];
      } else {
        const double jw = wq * pg.det;
        double Du[3];
#pragma unroll
        for (int d = 0; d < 3; d++) Du[d] = t[d] * pg.rdet;
        scale = pg.rdet;
        if (MODE == ELLIP_JACOBIAN) {                                        // EllipPointwiseJacobian, ellip.c:369-377
          const double eta = P.stash[sq], s0 = P.stash[splane + sq], s1 = P.stash[2 * splane + sq], s2 = P.stash[3 * splane + sq];
          const double dotw = (s0 * Du[0] + s1 * Du[1] + s2 * Du[2]) * jw, etaw = eta * jw;
          Dv[0] = etaw * Du[0] - dotw * s0;
          Dv[1] = etaw * Du[1] - dotw * s1;
          Dv[2] = etaw * Du[2] - dotw * s2;
          if (P.lambda != 0) v = -jw * P.stash[4 * splane + sq] * u[c];
        } else {                                                             // EllipPointwiseFunction + ComputeStore, ellip.c:333-367
          const double gamma = 0.5 * (Du[0] * Du[0] + Du[1] * Du[1] + Du[2] * Du[2]);
          const double espg = P.eps2 + gamma;
          const double eta = pow(espg, P.pm2h);
          const double sqrt_mdeta = sqrt(-P.pm2h * eta / espg);
          const double lexp = P.lambda != 0 ? P.lambda * exp(u[c]) : 0;
          P.stash[sq] = eta;
          P.stash[splane + sq] = sqrt_mdeta * Du[0];
          P.stash[2 * splane + sq] = sqrt_mdeta * Du[1];
          P.stash[3 * splane + sq] = sqrt_mdeta * Du[2];
          P.stash[4 * splane + sq] = lexp;
          const double f = P.forcing ? P.forcing[sq] : 0;
          v = -jw * f;
          v += -jw * lexp;
#pragma unroll
          for (int d = 0; d < 3; d++) Dv[d] = jw * eta * Du[d];
        }
      }
      // TensorRuleMappingTranspose, efstopo.c:593-623: reference-direction k gets sum_e Dv[e] jinv[k][e]
      Gx[L::at(ta, tb, c)] = scale * (Dv[0] * adj[0] + Dv[1] * adj[1] + Dv[2] * adj[2]);
      Gy[L::at(ta, tb, c)] = scale * (Dv[0] * adj[3] + Dv[1] * adj[4] + Dv[2] * adj[5]);
      gz[c] = scale * (Dv[0] * adj[6] + Dv[1] * adj[7] + Dv[2] * adj[8]);
      h[c] = v;
    }
    der_zT_add<NQ>(P.rt.Dt, gz, h);
  }
  __syncthreads();

  // ---- transposes ---------------------------------------------------------------------------------------------------------
  if (live) {
    der_xT<NQ>(P.rt.Dt, Gx, ta, tb);
    der_yT<NQ>(P.rt.Dt, Gy, ta, tb);
  }
  __syncthreads();
  if (live) {
#pragma unroll
    for (int c = 0; c < NQ; c++) h[c] += Gx[L::at(ta, tb, c)] + Gy[L::at(ta, tb, c)];
    bwd_z<NP, NQ>(P.bt.B, h, A, ta, tb);
  }
  __syncthreads();
  if (live && t < NQ * NP) { const int a = t / NP, k = t - a * NP; bwd_y<NP, NQ>(P.bt.B, A, a, k); }
  __syncthreads();
  if (live && t < NP * NP) {
    const int j = t / NP, k = t - j * NP;
    double o[NP];
    bwd_x<NP, NQ>(P.bt.B, A, j, k, o);
#pragma unroll
    for (int i = 0; i < NP; i++) {
      if (idx[i] < P.n) atomicAdd(P.y + idx[i], o[i]);
      else if (idx[i] >= P.nc) atomicAdd(P.yaux + (idx[i] - P.n), o[i]);       // ghosts; Dirichlet rows are dropped
    }
  }
}

}  // namespace dohp
