// Fused scalar elliptic (p-Laplacian / Bratu) element kernel: the whole dRulesetIterator traversal of EllipFunction /
// EllipShellMatMult (src/fs/tests/ellip.c:407-472) for a batch of elements in one thread block:
//   E gather (dFSGlobalToExpanded, src/fs/interface/fs.c:543-586) -> quadrature geometry -> u, grad u at the points ->
//   pointwise physics (ellip.c:333-377) -> transposes -> E^T scatter-add (dFSExpandedToGlobal, fs.c:592-640).
// Nothing but x, the index array, 24 vertex coordinates per element, the stash and y touches HBM.
// Second generation (eo_stages.cuh): even-odd 1-D contractions, conflict-free shared-memory scheme with host-built thread maps for
// the x and y passes, and a specialisation for meshes of AFFINE elements (geometry once per element instead of once per point; for
// the linear Poisson operator the pointwise step collapses to one symmetric 3x3 product).
#pragma once
#include "eo_stages.cuh"

namespace dohp {

enum { ELLIP_LINEAR = 0, ELLIP_JACOBIAN = 1, ELLIP_FUNCTION = 2,
       // the L2 projection of src/fs/tests/ex1.c through the same traversal: residual v = (u - f) jw, Dv = 0 (ProjResidual{,2,3},
       // ex1.c:294-430) and its derivative v = u jw (the mass operator); no derivative passes, no stash
       ELLIP_PROJ_FUNCTION = 3, ELLIP_PROJ_MULT = 4 };
#define ELLIP_NSTASH 5      // struct EllipStore {eta, sqrt_mdeta_Du[3], lambda_exp_u}, ellip.c:119-123

template <int NP, int NQ> struct EllipArgs {
  InterpMat<NP, NQ> bt;
  RuleTab2<NQ> rt;
  int nelem, n, nc;              // elements of the space; owned explicit nodes (indices >= n live in the aux arrays); closure size
  int count;                     // elements of this launch: elist[0..count) (or 0..count when elist is NULL)
  const int *elist;              // interior / ghost-touching split of a partitioned space (overlap of the ghost exchange, api.cu)
  int ES;                        // shared-memory stride between the elements of a block (doubles)
  const int *maps;               // [2][THREADS] lines of the x pass and of the y pass (smem_map.cpp)
  const int *e2l;                // [nelem][NP^3]
  const double *coords;          // [nelem][8][3]
  const double *x, *xaux;        // global vector [n]; [dirichlet | ghost] values [nc+ngh-n]
  double *y, *yaux;
  double *stash;                 // SoA: [NSTASH][nelem][NQ (c)][NQ*NQ (a,b)]
  const double *forcing;         // [nelem][NQ (c)][NQ*NQ] or NULL
  double pm2h, eps2, lambda;     // (p-2)/2, epsilon^2, lambda
  int nstage;                    // stash planes staged through shared memory by the Jacobian mode: 4, or 5 when lambda != 0
};

// threads per block the elements are batched for, per quadrature size and kernel class (measured, tools/sweep_bench.py; DOHP_TPB
// overrides it for such sweeps): small blocks decouple the memory phases of one block from the contractions of another
__host__ __device__ constexpr int ellip_tpb(int nq, bool stash, bool affine)
{
#ifdef DOHP_TPB
  return DOHP_TPB;
#else
  if (!affine && stash && (nq == 5 || nq == 6)) return 128;       // general geometry (profiles/r2o_block_size_sweep.txt, second table)
  if (!affine && !stash && nq >= 10) return 256;
  return stash ? (nq == 4 || nq == 7 ? 64 : nq == 5 || nq == 6 ? 160 : 128) : (nq == 7 || nq == 9 ? 256 : 128);      // profiles/r2o_block_size_sweep.txt
#endif
}

template <int NQ, int MODE, bool AFFINE> struct EllipCfg {
  static constexpr int T = NQ * NQ;                                   // threads per element (one per z line)
  static constexpr int TPB = ellip_tpb(NQ, MODE != 0, AFFINE);
  static constexpr int EPB = TPB / T < 1 ? 1 : TPB / T;                // elements per block
  static constexpr int THREADS = ((EPB * T + 31) / 32) * 32;
  static constexpr int ES_BASE = 3 * Lay2<NQ>::CS + 24 + DOHP_NGEO;    // A, Gx, Gy, the 8 vertices, the affine geometry
  // resident threads per SM the register budget is capped for (LINEAR ; stash modes, which keep more values live per point)
#ifndef DOHP_MINB_S
#define DOHP_MINB_S 4
#endif
#ifndef DOHP_MINB_M
#define DOHP_MINB_M 3
#endif
#ifndef DOHP_MINB_L
#define DOHP_MINB_L 2
#endif
#ifndef DOHP_MINB_NL_S
#define DOHP_MINB_NL_S 3
#endif
#ifndef DOHP_MINB_NL_M
#define DOHP_MINB_NL_M 2
#endif
#ifndef DOHP_MINB_NL_L
#define DOHP_MINB_NL_L 2
#endif
  // Jacobian mode: per-thread ring of RING points of the stash staged through shared memory by cp.async (eo_stages.cuh StashRing);
  // RING = 0: plain loads after an L2 prefetch at kernel start
#ifndef DOHP_RING
#define DOHP_RING 0
#endif
  static constexpr int RING = DOHP_RING < NQ ? DOHP_RING : NQ;
  static constexpr int stage_doubles(int nplanes) { return EPB * RING * nplanes * T; }
  static constexpr int MINB = NQ <= 6 ? DOHP_MINB_S : NQ <= 8 ? DOHP_MINB_M : DOHP_MINB_L;
  static constexpr int MINB_NL = NQ <= 6 ? DOHP_MINB_NL_S : NQ == 8 ? 3 : NQ <= 8 ? DOHP_MINB_NL_M : DOHP_MINB_NL_L;      // n = 8: 2.18 -> 2.00 ms with three (measured)
  // the general-geometry kernels keep the trilinear map of the thread's line live across the points (21 more doubles)
  // (the MINB numbers are resident 256-thread blocks)
  static constexpr int minb()
  {
    const bool affine = AFFINE;
    const int tsm = 256 * (affine ? (MODE == 0 ? MINB : MINB_NL) : (MODE == 0 ? (NQ <= 8 ? 3 : 2) : 2));
    return tsm / THREADS < 1 ? 1 : tsm / THREADS;
  }
};

template <int NP, int NQ, int MODE, bool AFFINE>
__global__ void __launch_bounds__(EllipCfg<NQ, MODE, AFFINE>::THREADS, EllipCfg<NQ, MODE, AFFINE>::minb()) ellip_kernel(const __grid_constant__ EllipArgs<NP, NQ> P)
{
  static_assert(NP == NQ, "the x and y passes are dealt one line per thread: NP == NQ");
  using L = Lay2<NQ>;
  using C = EllipCfg<NQ, MODE, AFFINE>;
  constexpr int T = C::T;
  constexpr bool PROJ = MODE == ELLIP_PROJ_FUNCTION || MODE == ELLIP_PROJ_MULT;
  extern __shared__ double smem[];
  const int tid = threadIdx.x;
  // z lines: natural order.  x lines (b,c) and y lines (a,c): whatever the host dealt to this thread
  const int lez = tid / T, tz = tid - lez * T, ta = tz / NQ, tb = tz - ta * NQ;
  const long iz = (long)blockIdx.x * C::EPB + lez;
  const bool livez = lez < C::EPB && iz < P.count;
  const long ez = livez && P.elist ? P.elist[iz] : iz;
  int lex, xb, xc, ley, ya, yc;
  map_decode(__ldg(P.maps + tid), lex, xb, xc);
  map_decode(__ldg(P.maps + C::THREADS + tid), ley, ya, yc);
  const long ix = (long)blockIdx.x * C::EPB + lex, iy = (long)blockIdx.x * C::EPB + ley;
  const bool livex = lex < C::EPB && ix < P.count, livey = ley < C::EPB && iy < P.count;
  const long ex = livex && P.elist ? P.elist[ix] : ix;
  double *Az = smem + (livez ? lez : 0) * P.ES, *Gxz = Az + L::CS, *Gyz = Gxz + L::CS, *X = Gyz + L::CS, *geo = X + 24;
  double *Ax = smem + (livex ? lex : 0) * P.ES, *Gxx = Ax + L::CS;
  double *Ay = smem + (livey ? ley : 0) * P.ES, *Gyy = Ay + 2 * L::CS;

  // ---- gather + x contraction -----------------------------------------------------------------------------------------
  int idx[NP];
  constexpr bool USE_RING = C::RING > 0;
  StashRing<NQ, T, ELLIP_NSTASH, USE_RING ? C::RING : 1> ring;
  if (MODE == ELLIP_JACOBIAN && livez) {
    if (USE_RING) {
      ring.init(smem + C::EPB * P.ES + lez * (C::RING * P.nstage * T) + tz, P.stash, ez, tz, (long)P.nelem * NQ * T, P.nstage);
      ring.prime();
    } else if ((tz & 15) == 0) {
      // the stash planes are read only after the three contractions: pull their lines into L2 now (one lane per 128 B)
      const long splane = (long)P.nelem * NQ * T;
#pragma unroll
      for (int c = 0; c < NQ; c++)
#pragma unroll
        for (int k = 0; k < ELLIP_NSTASH; k++)
          if (k < P.nstage) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.stash + k * splane + (ez * NQ + c) * T + tz));
    }
  }
#ifdef DOHP_PF_DIST
  {
    // the index lines of the elements a later wave of blocks will gather: pull them into L2 now (the gather is two dependent loads)
    const long ef = ((long)blockIdx.x + DOHP_PF_DIST) * C::EPB;
    constexpr int LINES = (C::EPB * NP * NP * NP * 4 + 127) / 128;
    if (!P.elist && ef + C::EPB <= P.count)
      for (int l = tid; l < LINES; l += C::THREADS) asm volatile("prefetch.global.L2 [%0];" ::"l"((const char *)(P.e2l + ef * (NP * NP * NP)) + l * 128));
  }
#endif
  if (livez) for (int s = tz; s < 24; s += T) X[s] = P.coords[ez * 24 + s];
  if (livex) {
    const int t = xb * NP + xc;
    double r[NP];
#pragma unroll
    for (int i = 0; i < NP; i++) idx[i] = __ldg(P.e2l + ex * (NP * NP * NP) + i * NP * NP + t);
#pragma unroll
    for (int i = 0; i < NP; i++) r[i] = idx[i] < P.n ? __ldg(P.x + idx[i]) : __ldg(P.xaux + (idx[i] - P.n));
    fwd_x2<NP, NQ>(P.bt, r, Ax, xb, xc);
  }
  __syncthreads();
  if (livey) fwd_y2<NP, NQ>(P.bt, Ay, ya, yc);
  if (AFFINE && livez && tz == 0) affine_geometry(X, geo);
  __syncthreads();
  double u[NQ], gz[NQ];
  if (livez) {
    fwd_z2<NP, NQ>(P.bt, Az, ta, tb, u);
    if (!PROJ) {
      P.rt.D.fwd(u, gz);
      put_z2<NQ>(Az, ta, tb, u);
    }
  }
  if (!PROJ) {
    __syncthreads();
    if (livex) der_x2<NQ>(P.rt.D, Ax, Gxx, xb, xc);
    if (livey) der_y2<NQ>(P.rt.D, Ay, Gyy, ya, yc);
  }
  __syncthreads();

  // ---- pointwise ----------------------------------------------------------------------------------------------------------
  double h[NQ];
  if (livez) {
    const double wab = P.rt.w[ta] * P.rt.w[tb];
    if (PROJ) {
      LineGeom lg;
      PointGeom pg;
      if (AFFINE) pg.det = geo[9];
      else line_geom(X, P.rt.q[ta], P.rt.q[tb], lg, false);
#pragma unroll
      for (int c = 0; c < NQ; c++) {
        if (!AFFINE) point_geom_adj(lg, P.rt.q[c], pg);
        const double jw = wab * P.rt.w[c] * pg.det;                               // jacobi.c:383-414
        const double f = MODE == ELLIP_PROJ_FUNCTION && P.forcing ? P.forcing[(ez * NQ + c) * T + tz] : 0.0;
        h[c] = (u[c] - f) * jw;                                                   // ex1.c:334, :421
      }
    } else if (AFFINE && MODE == ELLIP_LINEAR) {
      // p = 2, lambda = 0 on a parallelepiped: Dv_ref = w_q (adj adj^T / det) Du_ref
      const double m00 = geo[11] * wab, m01 = geo[12] * wab, m02 = geo[13] * wab, m11 = geo[14] * wab, m12 = geo[15] * wab, m22 = geo[16] * wab;
#pragma unroll
      for (int c = 0; c < NQ; c++) {
        const double wc = P.rt.w[c];
        const double g0 = Gxz[L::at(ta, tb, c)] * wc, g1 = Gyz[L::at(ta, tb, c)] * wc, g2 = gz[c] * wc;
        Gxz[L::at(ta, tb, c)] = m00 * g0 + m01 * g1 + m02 * g2;
        Gyz[L::at(ta, tb, c)] = m01 * g0 + m11 * g1 + m12 * g2;
        gz[c] = m02 * g0 + m12 * g1 + m22 * g2;
        h[c] = 0;
      }
    } else {
      LineGeom lg;
      PointGeom pg;
      if (AFFINE) {
#pragma unroll
        for (int i = 0; i < 9; i++) pg.adj[i] = geo[i];
        pg.det = geo[9];
        pg.rdet = geo[10];
      } else line_geom(X, P.rt.q[ta], P.rt.q[tb], lg, false);
      const long splane = (long)P.nelem * NQ * T;
#pragma unroll
      for (int c = 0; c < NQ; c++) {
        if (!AFFINE) point_geom_adj(lg, P.rt.q[c], pg);
        const double *adj = pg.adj;
        const double g0 = Gxz[L::at(ta, tb, c)], g1 = Gyz[L::at(ta, tb, c)], g2 = gz[c];
        // TensorRuleMapping, efstopo.c:574, with jinv = adj/det: t = det * Du
        double t[3], Dv[3], v = 0;
#pragma unroll
        for (int d = 0; d < 3; d++) t[d] = g0 * adj[0 + d] + g1 * adj[3 + d] + g2 * adj[6 + d];
        const double wq = wab * P.rt.w[c];                                      // jw = wq * det (jacobi.c:383-414)
        const long sq = (ez * NQ + c) * T + tz;                                 // point index in the stash planes
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
            double eta, s0, s1, s2, lexp = 0;
            if (USE_RING) {
              ring.wait(c);
              eta = ring.get(c, 0); s0 = ring.get(c, 1); s1 = ring.get(c, 2); s2 = ring.get(c, 3);
              if (P.lambda != 0) lexp = ring.get(c, 4);
              ring.refill(c);
            } else {
              eta = P.stash[sq]; s0 = P.stash[splane + sq]; s1 = P.stash[2 * splane + sq]; s2 = P.stash[3 * splane + sq];
              if (P.lambda != 0) lexp = P.stash[4 * splane + sq];
            }
            const double dotw = (s0 * Du[0] + s1 * Du[1] + s2 * Du[2]) * jw, etaw = eta * jw;
            Dv[0] = etaw * Du[0] - dotw * s0;
            Dv[1] = etaw * Du[1] - dotw * s1;
            Dv[2] = etaw * Du[2] - dotw * s2;
            if (P.lambda != 0) v = -jw * lexp * u[c];
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
        Gxz[L::at(ta, tb, c)] = scale * (Dv[0] * adj[0] + Dv[1] * adj[1] + Dv[2] * adj[2]);
        Gyz[L::at(ta, tb, c)] = scale * (Dv[0] * adj[3] + Dv[1] * adj[4] + Dv[2] * adj[5]);
        gz[c] = scale * (Dv[0] * adj[6] + Dv[1] * adj[7] + Dv[2] * adj[8]);
        h[c] = v;
      }
    }
    if (!PROJ) der_zT_add2<NQ>(P.rt.D, gz, h);
  }

  // ---- transposes ---------------------------------------------------------------------------------------------------------
  if (!PROJ) {
    __syncthreads();
    if (livex) der_xT2<NQ>(P.rt.D, Gxx, xb, xc);
    if (livey) der_yT2<NQ>(P.rt.D, Gyy, ya, yc);
  }
  __syncthreads();
  if (livez) {
    if (!PROJ) {
#pragma unroll
      for (int c = 0; c < NQ; c++) h[c] += Gxz[L::at(ta, tb, c)] + Gyz[L::at(ta, tb, c)];
    }
    bwd_z2<NP, NQ>(P.bt, h, Az, ta, tb);
  }
  __syncthreads();
  if (livey) bwd_y2<NP, NQ>(P.bt, Ay, ya, yc);
  __syncthreads();
  if (livex) {
    double o[NP];
    bwd_x2<NP, NQ>(P.bt, Ax, xb, xc, o);
#pragma unroll
    for (int i = 0; i < NP; i++) {
      if (idx[i] < P.n) atomicAdd(P.y + idx[i], o[i]);
      else if (idx[i] >= P.nc) atomicAdd(P.yaux + (idx[i] - P.n), o[i]);       // ghosts; Dirichlet rows are dropped
    }
  }
}

}  // namespace dohp
