// Fused element kernels for the vector-valued examples: St-Venant-Kirchhoff elasticity (src/fs/tests/elast.c:220-341,
// one bs=3 space) and power-law Stokes (src/fs/tests/stokes.c:447-673, velocity bs=3 of degree p plus pressure bs=1 of
// degree p-1, both tabulated on the velocity rule, stokes.c:292-298).  Same traversal as ellip_kernel: E gather ->
// geometry -> values/gradients at the points -> pointwise physics -> transposes -> E^T scatter-add, one thread per z line.
#pragma once
#include "eo_stages.cuh"

namespace dohp {

enum { VEC_ELAST = 0, VEC_STOKES = 1 };
enum { VEC_FUNCTION = 0, VEC_JACOBIAN = 1 };
enum { USE_U = 1, USE_P = 2, OUT_U = 4, OUT_P = 8, BLOCK_A = 16 };
#define ELAST_NSTASH 9       // struct ElastStore {F[9]}, elast.c:39-41
#define STOKES_NSTASH 8      // struct StokesStore {eta, deta, Du[6]}, stokesimpl.h:28-31

template <int NP, int NPP, int NQ> struct VecArgs {
  InterpMat<NP, NQ> bt;
  InterpMat<NPP, NQ> bp;         // pressure basis on the velocity rule (stokes only)
  RuleTab2<NQ> rt;
  int nelem, flags;
  int count;                     // elements of this launch: elist[0..count) (or 0..count when elist is NULL)
  const int *elist;
  int ES;                        // shared-memory stride between the elements of a block (doubles)
  const int *maps;               // [2][THREADS] lines of the x pass and of the y pass (smem_map.cpp)
  int n_u, nc_u, n_p, nc_p;      // owned / closure node counts of the two spaces
  const int *e2l_u, *e2l_p;
  const double *coords;
  const double *xu, *xuaux, *xp, *xpaux;
  double *yu, *yuaux, *yp, *ypaux;
  double *stash;                 // planes [NSTASH][nelem][NQ (c)][NQ*NQ (a,b)]
  const double *forcing;         // planes [3 or 4][nelem][NQ][NQ*NQ] or NULL
  double c0, c1, c2, c3;         // elast: lambda, mu ; stokes: B, eps^2/2, (p-2)/2, nu/(1-2nu)
};

template <int NQ, int KIND> struct VecCfg {
  static constexpr int T = NQ * NQ;
#ifdef DOHP_VTPB
  static constexpr int EPB = DOHP_VTPB / T < 1 ? 1 : DOHP_VTPB / T;
#else
  static constexpr int EPB = NQ == 7 ? 3 : (T >= 96) ? 1 : (128 / T);      // NQ = 7: 147 of 160 lanes instead of 98 of 128
#endif
  static constexpr int THREADS = ((EPB * T + 31) / 32) * 32;
  static constexpr int NBUF = 9 + (KIND == VEC_STOKES ? 1 : 0);          // 3 components x (A,Gx,Gy) + pressure
  static constexpr int ES_BASE = NBUF * Lay2<NQ>::CS + 24 + DOHP_NGEO;
  // Jacobian mode: per-thread ring of RING points of the stash staged through shared memory by cp.async (eo_stages.cuh StashRing)
#ifndef DOHP_VRING
#define DOHP_VRING 0
#endif
  static constexpr int NPL = KIND == VEC_STOKES ? 8 : 9;
  static constexpr int RING = DOHP_VRING < NQ ? DOHP_VRING : NQ;
  static constexpr int STAGE_DOUBLES = EPB * RING * NPL * T;
#ifndef DOHP_VTSM_S
#define DOHP_VTSM_S 512                   // measured: stokes p4/p3 30.6 -> 32.2 GDoF/s against 384 (three resident blocks)
#endif
  static constexpr int TSM = NQ <= 6 ? DOHP_VTSM_S : (NQ <= 8 ? 320 : 0);         // resident threads per SM the register budget is capped for
  static constexpr int MINB = TSM / THREADS < 1 ? 1 : TSM / THREADS;
};

DI double dot_column(const double (&a)[9], const double (&b)[9], int i, int j) { return a[i] * b[j] + a[3 + i] * b[3 + j] + a[6 + i] * b[6 + j]; }
// C (+)= A * S, S symmetric in compressed storage [00,11,22,01,02,12] (dTensorMult(Add)GESY3, include/dohp.h:83-102)
template <bool ADD> DI void mult_ge_sy3(double (&C)[9], const double (&A)[9], const double (&S)[6])
{
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const double c0 = A[i * 3] * S[0] + A[i * 3 + 1] * S[3] + A[i * 3 + 2] * S[4];
    const double c1 = A[i * 3] * S[3] + A[i * 3 + 1] * S[1] + A[i * 3 + 2] * S[5];
    const double c2 = A[i * 3] * S[4] + A[i * 3 + 1] * S[5] + A[i * 3 + 2] * S[2];
    if (ADD) { C[i * 3] += c0; C[i * 3 + 1] += c1; C[i * 3 + 2] += c2; }
    else { C[i * 3] = c0; C[i * 3 + 1] = c1; C[i * 3 + 2] = c2; }
  }
}
// ElastPiolaKirchoff2, elast.c:227-241
DI void elast_pk2(double lambda, double mu, const double (&F)[9], double (&S)[6])
{
  double E[6];
  E[0] = 0.5 * (dot_column(F, F, 0, 0) - 1);
  E[1] = 0.5 * (dot_column(F, F, 1, 1) - 1);
  E[2] = 0.5 * (dot_column(F, F, 2, 2) - 1);
  E[3] = 0.5 * dot_column(F, F, 0, 1);
  E[4] = 0.5 * dot_column(F, F, 0, 2);
  E[5] = 0.5 * dot_column(F, F, 1, 2);
  const double trace = E[0] + E[1] + E[2];
#pragma unroll
  for (int i = 0; i < 6; i++) S[i] = (i < 3 ? lambda * trace : 0.0) + 2 * mu * E[i];
}
// ElastPiolaKirchoff2Jacobian, elast.c:242-256
DI void elast_pk2_jacobian(double lambda, double mu, const double (&F)[9], const double (&dF)[9], double (&dS)[6])
{
  double dE[6];
  dE[0] = 0.5 * (dot_column(F, dF, 0, 0) + dot_column(dF, F, 0, 0));
  dE[1] = 0.5 * (dot_column(F, dF, 1, 1) + dot_column(dF, F, 1, 1));
  dE[2] = 0.5 * (dot_column(F, dF, 2, 2) + dot_column(dF, F, 2, 2));
  dE[3] = 0.5 * (dot_column(F, dF, 0, 1) + dot_column(dF, F, 0, 1));
  dE[4] = 0.5 * (dot_column(F, dF, 0, 2) + dot_column(dF, F, 0, 2));
  dE[5] = 0.5 * (dot_column(F, dF, 1, 2) + dot_column(dF, F, 1, 2));
  const double dtrace = dE[0] + dE[1] + dE[2];
#pragma unroll
  for (int i = 0; i < 6; i++) dS[i] = (i < 3 ? lambda * dtrace : 0.0) + 2 * mu * dE[i];
}
DI double colon_sym3(const double (&a)[6], const double (&b)[6])      // dColonSymScalar3, dohp.h:77-78
{ return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + 2 * a[3] * b[3] + 2 * a[4] * b[4] + 2 * a[5] * b[5]; }

template <int NP, int NPP, int NQ, int KIND, int MODE, bool AFFINE>
__global__ void __launch_bounds__(VecCfg<NQ, KIND>::THREADS, VecCfg<NQ, KIND>::MINB) vec_kernel(const __grid_constant__ VecArgs<NP, NPP, NQ> P)
{
  static_assert(NP == NQ, "the x and y passes are dealt one line per thread: NP == NQ");
  using L = Lay2<NQ>;
  using C = VecCfg<NQ, KIND>;
  constexpr int T = C::T, CS = L::CS;
  constexpr bool STOKES = KIND == VEC_STOKES;
  extern __shared__ double smem[];
  const int tid = threadIdx.x;
  // z lines: natural order; x lines (j,k)/(b,c) and y lines (a,k)/(a,c): dealt by the host (smem_map.cpp)
  const int le = tid / T, t = tid - le * T, ta = t / NQ, tb = t - ta * NQ;
  const long ie = (long)blockIdx.x * C::EPB + le;
  const bool live = le < C::EPB && ie < P.count;
  const long e = live && P.elist ? P.elist[ie] : ie;
  int lex, xb, xc, ley, ya, yc;
  map_decode(__ldg(P.maps + tid), lex, xb, xc);
  map_decode(__ldg(P.maps + C::THREADS + tid), ley, ya, yc);
  const long ix = (long)blockIdx.x * C::EPB + lex;
  const bool livex = lex < C::EPB && ix < P.count, livey = ley < C::EPB && (long)blockIdx.x * C::EPB + ley < P.count;
  const long ex = livex && P.elist ? P.elist[ix] : ix;
  double *S0 = smem + (live ? le : 0) * P.ES;                 // component d: A = S0+3d*CS, Gx = A+CS, Gy = A+2CS
  double *Ap = S0 + 9 * CS;                                     // pressure buffer (stokes)
  double *X = S0 + C::NBUF * CS, *geo = X + 24;
  double *S0x = smem + (livex ? lex : 0) * P.ES, *Apx = S0x + 9 * CS;
  double *S0y = smem + (livey ? ley : 0) * P.ES, *Apy = S0y + 9 * CS;
  const bool use_u = P.flags & USE_U, use_p = STOKES && (P.flags & USE_P), out_u = P.flags & OUT_U, out_p = STOKES && (P.flags & OUT_P);

  // ---- gather + x contraction ---------------------------------------------------------------------------------------------
  int idx[NP], idxp[STOKES ? NPP : 1];
  constexpr bool USE_RING = C::RING > 0;
  StashRing<NQ, T, C::NPL, USE_RING ? C::RING : 1> ring;
  const bool read_stash = MODE == VEC_JACOBIAN && (!STOKES || (out_u && use_u));      // the B and B^T blocks do not touch it
  if (read_stash && live) {
    if (USE_RING) {
      ring.init(smem + C::EPB * P.ES + le * (C::RING * C::NPL * T) + t, P.stash, e, t, (long)P.nelem * NQ * T, C::NPL);
      ring.prime();
    } else if ((t & 15) == 0) {
      // the stash planes are read only after the contractions: pull their lines into L2 now (one lane per 128 B)
      const long splane = (long)P.nelem * NQ * T;
#pragma unroll
      for (int c = 0; c < NQ; c++)
#pragma unroll
        for (int k = 0; k < C::NPL; k++) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.stash + k * splane + (e * NQ + c) * T + t));
    }
  }
  if (live) for (int s = t; s < 24; s += T) X[s] = P.coords[e * 24 + s];
  if (livex) {
    const int tx = xb * NP + xc;
#pragma unroll
    for (int i = 0; i < NP; i++) idx[i] = __ldg(P.e2l_u + ex * (NP * NP * NP) + i * NP * NP + tx);
    if (use_u) {
#pragma unroll
      for (int d = 0; d < 3; d++) {
        double r[NP];
#pragma unroll
        for (int i = 0; i < NP; i++) r[i] = idx[i] < P.n_u ? __ldg(P.xu + (long)idx[i] * 3 + d) : __ldg(P.xuaux + (long)(idx[i] - P.n_u) * 3 + d);
        fwd_x2<NP, NQ>(P.bt, r, S0x + 3 * d * CS, xb, xc);
      }
    }
    if (STOKES && xb < NPP && xc < NPP) {
      const int tp = xb * NPP + xc;
#pragma unroll
      for (int i = 0; i < NPP; i++) idxp[i] = __ldg(P.e2l_p + ex * (NPP * NPP * NPP) + i * NPP * NPP + tp);
      if (use_p) {
        double r[NPP];
#pragma unroll
        for (int i = 0; i < NPP; i++) r[i] = idxp[i] < P.n_p ? __ldg(P.xp + idxp[i]) : __ldg(P.xpaux + (idxp[i] - P.n_p));
        fwd_x2<NPP, NQ>(P.bp, r, Apx, xb, xc);
      }
    }
  }
  __syncthreads();
  if (livey && use_u) {
#pragma unroll
    for (int d = 0; d < 3; d++) fwd_y2<NP, NQ>(P.bt, S0y + 3 * d * CS, ya, yc);
  }
  if (livey && use_p && yc < NPP) fwd_y2<NPP, NQ>(P.bp, Apy, ya, yc);
  if (AFFINE && live && t == 0) affine_geometry(X, geo);
  __syncthreads();
  double gz[3][NQ], h[3][NQ], pq[NQ];
  if (live) {
    if (use_u) {
#pragma unroll
      for (int d = 0; d < 3; d++) {
        double u[NQ];
        fwd_z2<NP, NQ>(P.bt, S0 + 3 * d * CS, ta, tb, u);
        P.rt.D.fwd(u, gz[d]);
        put_z2<NQ>(S0 + 3 * d * CS, ta, tb, u);
      }
    }
    if (STOKES) {
      if (use_p) fwd_z2<NPP, NQ>(P.bp, Ap, ta, tb, pq);
      else {
#pragma unroll
        for (int c = 0; c < NQ; c++) pq[c] = 0;
      }
    }
  }
  __syncthreads();
  if (use_u) {
#pragma unroll
    for (int d = 0; d < 3; d++) {
      if (livex) der_x2<NQ>(P.rt.D, S0x + 3 * d * CS, S0x + 3 * d * CS + CS, xb, xc);
      if (livey) der_y2<NQ>(P.rt.D, S0y + 3 * d * CS, S0y + 3 * d * CS + 2 * CS, ya, yc);
    }
  }
  __syncthreads();

  // ---- pointwise ----------------------------------------------------------------------------------------------------------------
  if (live) {
    LineGeom lg;
    PointGeom pg;
    if (AFFINE) {
#pragma unroll
      for (int i = 0; i < 9; i++) pg.adj[i] = geo[i];
      pg.det = geo[9];
      pg.rdet = geo[10];
    } else line_geom(X, P.rt.q[ta], P.rt.q[tb], lg, false);
    const double wab = P.rt.w[ta] * P.rt.w[tb];
    const long splane = (long)P.nelem * NQ * T;
#pragma unroll
    for (int c = 0; c < NQ; c++) {
      if (!AFFINE) point_geom_adj(lg, P.rt.q[c], pg);
      const double *adj = pg.adj;
      const double jw = wab * P.rt.w[c] * pg.det;
      const long sq = (e * NQ + c) * T + t;
      double Du[9], Dv[9], v[3] = {0, 0, 0};
#pragma unroll
      for (int d = 0; d < 3; d++) {
        const double g0 = use_u ? S0[3 * d * CS + CS + L::at(ta, tb, c)] : 0.0, g1 = use_u ? S0[3 * d * CS + 2 * CS + L::at(ta, tb, c)] : 0.0;
        const double g2 = use_u ? gz[d][c] : 0.0;
#pragma unroll
        for (int x = 0; x < 3; x++) Du[d * 3 + x] = pg.rdet * (g0 * adj[x] + g1 * adj[3 + x] + g2 * adj[6 + x]);   // efstopo.c:574
      }
      if (KIND == VEC_ELAST) {
        const double lambda = P.c0, mu = P.c1;
        double F[9], S[6];
        if (MODE == VEC_FUNCTION) {                                          // ElastPointwiseFunction, elast.c:257-268
#pragma unroll
          for (int i = 0; i < 9; i++) { F[i] = ((i % 4 == 0) ? 1.0 : 0.0) + Du[i]; P.stash[i * splane + sq] = F[i]; }
          if (P.forcing) {
#pragma unroll
            for (int i = 0; i < 3; i++) v[i] = -jw * P.forcing[i * splane + sq];
          }
          elast_pk2(lambda, mu, F, S);
          mult_ge_sy3<false>(Dv, F, S);
        } else {                                                             // ElastPointwiseJacobian, elast.c:270-280
          double dS[6];
          if (USE_RING) {
            ring.wait(c);
#pragma unroll
            for (int i = 0; i < 9; i++) F[i] = ring.get(c, i);
            ring.refill(c);
          } else {
#pragma unroll
            for (int i = 0; i < 9; i++) F[i] = P.stash[i * splane + sq];
          }
          elast_pk2(lambda, mu, F, S);
          elast_pk2_jacobian(lambda, mu, F, Du, dS);
          mult_ge_sy3<false>(Dv, F, dS);
          mult_ge_sy3<true>(Dv, Du, S);
        }
#pragma unroll
        for (int i = 0; i < 9; i++) Dv[i] *= jw;
      } else {
        // dTensorSymCompress3, dohp.h:79-80
        const double D6[6] = {Du[0], Du[4], Du[8], (Du[1] + Du[3]) / 2, (Du[2] + Du[6]) / 2, (Du[5] + Du[7]) / 2};
        const double B = P.c0, eps2h = P.c1, pm2h = P.c2;
        const double pv = pq[c], tr = D6[0] + D6[1] + D6[2];
        double V6[6] = {0, 0, 0, 0, 0, 0}, q;
        if (MODE == VEC_FUNCTION) {                                          // StokesPointwiseFunction + ComputeStore, stokes.c:447-475
          const double gamma_reg = eps2h + 0.5 * colon_sym3(D6, D6);
          const double eta = B * pow(gamma_reg, pm2h);
          P.stash[sq] = eta;
          P.stash[splane + sq] = pm2h * eta / gamma_reg;
#pragma unroll
          for (int i = 0; i < 6; i++) P.stash[(2 + i) * splane + sq] = D6[i];
          double fp = 0;
          if (P.forcing) {
#pragma unroll
            for (int i = 0; i < 3; i++) v[i] = -jw * P.forcing[i * splane + sq];
            fp = P.forcing[3 * splane + sq];
          }
          q = -jw * (tr + fp);
#pragma unroll
          for (int i = 0; i < 3; i++) V6[i] = jw * (eta * D6[i] - pv);       // lambda = 0*StokesLambda, stokes.c:468
#pragma unroll
          for (int i = 3; i < 6; i++) V6[i] = jw * eta * D6[i];
        } else {                                                             // StokesPointwiseJacobian{,_A,_B,_Bt}, stokes.c:477-507
          q = -jw * tr;
          if (!out_u) {
          } else if (use_u) {
            double eta, deta, W6[6];
            if (USE_RING) {
              ring.wait(c);
              eta = ring.get(c, 0); deta = ring.get(c, 1);
#pragma unroll
              for (int i = 0; i < 6; i++) W6[i] = ring.get(c, 2 + i);
              ring.refill(c);
            } else {
              eta = P.stash[sq]; deta = P.stash[splane + sq];
#pragma unroll
              for (int i = 0; i < 6; i++) W6[i] = P.stash[(2 + i) * splane + sq];
            }
            const double deta_colon = deta * colon_sym3(W6, D6);
            const double lam = (P.flags & BLOCK_A) ? P.c3 : 0.0;             // the A block keeps the penalty, the coupled J multiplies it by 0
#pragma unroll
            for (int i = 0; i < 6; i++) V6[i] = jw * (eta * D6[i] + deta_colon * W6[i] + (i < 3 ? lam * tr - pv : 0.0));
          } else {
#pragma unroll
            for (int i = 0; i < 6; i++) V6[i] = i < 3 ? -jw * pv : 0.0;
          }
        }
        pq[c] = q;
        // dTensorSymUncompress3, dohp.h:81-82
        Dv[0] = V6[0]; Dv[4] = V6[1]; Dv[8] = V6[2]; Dv[1] = Dv[3] = V6[3]; Dv[2] = Dv[6] = V6[4]; Dv[5] = Dv[7] = V6[5];
      }
      if (out_u) {
#pragma unroll
        for (int d = 0; d < 3; d++) {                                        // TensorRuleMappingTranspose, efstopo.c:593-623
          const double *dv = Dv + 3 * d;
          S0[3 * d * CS + CS + L::at(ta, tb, c)] = pg.rdet * (dv[0] * adj[0] + dv[1] * adj[1] + dv[2] * adj[2]);
          S0[3 * d * CS + 2 * CS + L::at(ta, tb, c)] = pg.rdet * (dv[0] * adj[3] + dv[1] * adj[4] + dv[2] * adj[5]);
          gz[d][c] = pg.rdet * (dv[0] * adj[6] + dv[1] * adj[7] + dv[2] * adj[8]);
          h[d][c] = v[d];
        }
      }
    }
    if (out_u) {
#pragma unroll
      for (int d = 0; d < 3; d++) der_zT_add2<NQ>(P.rt.D, gz[d], h[d]);
    }
  }
  __syncthreads();

  // ---- transposes ------------------------------------------------------------------------------------------------------------------
  if (out_u) {
#pragma unroll
    for (int d = 0; d < 3; d++) {
      if (livex) der_xT2<NQ>(P.rt.D, S0x + 3 * d * CS + CS, xb, xc);
      if (livey) der_yT2<NQ>(P.rt.D, S0y + 3 * d * CS + 2 * CS, ya, yc);
    }
  }
  if (live && out_p) bwd_z2<NPP, NQ>(P.bp, pq, Ap, ta, tb);
  __syncthreads();
  if (live && out_u) {
#pragma unroll
    for (int d = 0; d < 3; d++) {
#pragma unroll
      for (int c = 0; c < NQ; c++) h[d][c] += S0[3 * d * CS + CS + L::at(ta, tb, c)] + S0[3 * d * CS + 2 * CS + L::at(ta, tb, c)];
      bwd_z2<NP, NQ>(P.bt, h[d], S0 + 3 * d * CS, ta, tb);
    }
  }
  if (livey && out_p && yc < NPP) bwd_y2<NPP, NQ>(P.bp, Apy, ya, yc);
  __syncthreads();
  if (livey && out_u) {
#pragma unroll
    for (int d = 0; d < 3; d++) bwd_y2<NP, NQ>(P.bt, S0y + 3 * d * CS, ya, yc);
  }
  __syncthreads();
  if (livex && out_u) {
#pragma unroll
    for (int d = 0; d < 3; d++) {
      double o[NP];
      bwd_x2<NP, NQ>(P.bt, S0x + 3 * d * CS, xb, xc, o);
#pragma unroll
      for (int i = 0; i < NP; i++) {
        if (idx[i] < P.n_u) atomicAdd(P.yu + (long)idx[i] * 3 + d, o[i]);
        else if (idx[i] >= P.nc_u) atomicAdd(P.yuaux + (long)(idx[i] - P.n_u) * 3 + d, o[i]);
      }
    }
  }
  if (livex && out_p && xb < NPP && xc < NPP) {
    double o[NPP];
    bwd_x2<NPP, NQ>(P.bp, Apx, xb, xc, o);
#pragma unroll
    for (int i = 0; i < NPP; i++) {
      if (idxp[i] < P.n_p) atomicAdd(P.yp + idxp[i], o[i]);
      else if (idxp[i] >= P.nc_p) atomicAdd(P.ypaux + (idxp[i] - P.n_p), o[i]);
    }
  }
}

}  // namespace dohp
