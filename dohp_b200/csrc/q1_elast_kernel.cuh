// Embedded-Q1 assembly of the elasticity Jacobian (ElastJacobian, src/fs/tests/elast.c:343-407), block size 3, sm_100a.
//
// Same organisation as q1_kernel.cuh (state on the SPARSE points by sum factorisation, patch matrices reduced inside the
// element through shared memory, one RED.ADD.F64 per CSR value) with the 24 x 24 patch matrix split by component pair:
//   K[(i,fi),(j,fj)] = sum_q d_i^T T_q^{fi fj} d_j ,
//   T^{fi fj} = jw ( lambda g_fi (x) g_fj + mu g_fj (x) g_fi + mu (F_fi . F_fj) jq jq^T + delta_{fi fj} jq S jq^T ),  g_f = jq F_f
// which is ElastPointwiseJacobian (elast.c:270-280: dS = lambda tr(dE) 1 + 2 mu dE, Dv = F dS + dF S) applied to the unit
// trial function e_fj phi_j and tested with e_fi phi_i, with the patch basis gradient G_b = jq^T d_b folded into the
// 3 x 3 tensor (jq = the inverse Jacobian the reference maps the patch basis with, quirk (ii) of q1.cu; d_b = the
// reference-space patch basis with deriv1 = +-1, quirk (i)).  K^{fj fi}[j][i] = K^{fi fj}[i][j], so per patch
//   3 "diagonal" units (fi = fj, symmetric 8 x 8, 36 accumulators, stored full) and
//   6 "off-diagonal half" units (fi < fj, rows 4h..4h+3 x 8 columns, 32 accumulators)
// are computed, one thread each, warps homogeneous in the unit type.
// Phases per slab of patch layers:  B x/y contraction -> D one thread per SPARSE point: F, geometry, and the per-point
// ingredients (jw, g, F_f.F_f', jq jq^T, jw jq S jq^T: 28 doubles) -> C units -> G one thread per (row node, fi, fj): the
// 27 neighbour blocks from the <= 8 patches around the node in registers, then 27 RED.ADD.F64 in which the nine lanes of a
// node write the nine consecutive values of one BAIJ block.
#pragma once
#include "q1_kernel.cuh"

namespace dohp {

#ifdef DOHP_PROFILE
#define Q1E_DBG(bit) (A.dbg & (bit))      // measurement ablations (DOHP_Q1_DBG): 1 no REDs, 2 no patch units, 4 no gather, 8 no point phase, 16 no x/y contraction
#else
#define Q1E_DBG(bit) false
#endif
#ifndef Q1E_QUNROLL
#define Q1E_QUNROLL 8                     // the eight points of a patch unrolled: the reference-space patch basis becomes immediate operands
#endif

constexpr int Q1E_QU = Q1E_QUNROLL;

// tile of patches processed at a time: SL layers (i) x TJ columns (j) x all P (k); P >= 7 needs column tiles to fit in shared memory
#ifndef Q1E_TJ6
#define Q1E_TJ6 6
#endif
__host__ __device__ constexpr int q1e_slab(int P) { return P == 1 ? 1 : P == 2 ? 2 : P == 3 ? 3 : P == 4 ? 2 : 1; }
__host__ __device__ constexpr int q1e_tj(int P) { return P == 6 ? Q1E_TJ6 : P <= 6 ? P : 2; }      // P = 6 measured: whole layers (1 block/SM) 77 ms, 3-column tiles (2 blocks/SM) 88 ms
__host__ __device__ constexpr int q1e_epb(int P) { return P == 1 ? 8 : P == 2 ? 2 : 1; }

template <int P> struct Q1ECfg {
  static constexpr int N = P + 1, QS = 2 * P, N3 = N * N * N;
  static constexpr int SL = q1e_slab(P), TJ = q1e_tj(P), EPB = q1e_epb(P), QL = 2 * SL, QJ = 2 * TJ;
  static constexpr int NPB = EPB * SL * TJ * P;                   // patches per tile per block
  static constexpr int KBLK = 65, KSTR = 6 * KBLK + 1;            // 3 (diagonal pairs, full 8x8) + 3 (fi < fj) blocks of 64, padded: the nine
                                                                  // lanes of a row node read different blocks at the same offset (gather phase)
  static constexpr int NPD = 28;                                  // per-point ingredients
  static constexpr int THREADS = ((9 * NPB + 31) / 32) * 32 < 64 ? 64 : ((9 * NPB + 31) / 32) * 32;
  static constexpr int NW = THREADS / 32;
  static constexpr int MINB = 384 / THREADS < 1 ? 1 : (384 / THREADS > 2 ? 2 : 384 / THREADS);      // about 168 registers
  static constexpr int TAB = 2 * QS * N + 2 * QS;
  static constexpr int XU = 36 + 3 * N3;                          // vertices, affine geometry (adj[9], det, 1/det, flag), nodal state [3][n^3]
  static constexpr int DS = 3 * (2 * QL * N * N + 3 * QL * QJ * N);
  static constexpr int BX = TAB, BD = BX + EPB * XU, BP = BD + EPB * DS, BK = BP + NPD * 8 * NPB;
  static constexpr int NREC = EPB * (SL + 1) * (TJ + 1) * N;       // row nodes of a tile: their 32-byte records are staged by cp.async while
  static constexpr int BR = (BK + NPB * KSTR + 1) / 2 * 2;                     // the patch units run
  static constexpr int SMEM_DOUBLES = BR + 4 * NREC;
};

// reference-space patch basis at point q = (q0,q1,q2): interp1[q][b] = IA if b is on the point's side else IB, deriv1 = -1,+1
// (efstopo.c:665-666); yz[(b1,b2)] etc. are the products the three derivative components and the value are made of
DI void q1_patch_basis(int q0, int q1, int q2, double (&yz)[4], double (&xz)[4], double (&xy)[4])
{
  constexpr double IA = 0.78867513459481287, IB = 0.21132486540518713;
  const double ax[2] = {q0 ? IB : IA, q0 ? IA : IB}, ay[2] = {q1 ? IB : IA, q1 ? IA : IB}, az[2] = {q2 ? IB : IA, q2 ? IA : IB};
#pragma unroll
  for (int m = 0; m < 4; m++) { yz[m] = ay[m >> 1] * az[m & 1]; xz[m] = ax[m >> 1] * az[m & 1]; xy[m] = ax[m >> 1] * ay[m & 1]; }
}
#define Q1_D0(b, yz) (((b) >> 2) ? yz[(b) & 3] : -yz[(b) & 3])
#define Q1_D1(b, xz) ((((b) >> 1) & 1) ? xz[((b) >> 2) * 2 + ((b) & 1)] : -xz[((b) >> 2) * 2 + ((b) & 1)])
#define Q1_D2(b, xy) (((b) & 1) ? xy[(b) >> 1] : -xy[(b) >> 1])

// record bytes t with (t / 3) % 3 == jrow, as a mask of the packed word w (neighbour columns of the tile before / after this one)
__host__ __device__ constexpr unsigned q1_dropmask(int w, int jrow)
{
  unsigned m = 0;
  for (int b = 0; b < 4; b++) { const int t = 4 * w + b; if (t < 27 && (t / 3) % 3 == jrow) m |= 0xFFu << (8 * b); }
  return m;
}
// slot of h_f[k] among the 28 per-point doubles of the Stokes variant
__host__ __device__ constexpr int q1s_h(int f, int k) { return f * 3 + k < 3 ? 13 + f * 3 + k : 22 + (f * 3 + k - 3); }

template <int P, int KIND> __global__ void __launch_bounds__(Q1ECfg<P>::THREADS, Q1ECfg<P>::MINB) q1_elast_kernel(const Q1Launch A)
{
  using C = Q1ECfg<P>;
  constexpr int N = C::N, QS = C::QS, N3 = C::N3, SL = C::SL, TJ = C::TJ, QL = C::QL, QJ = C::QJ, EPB = C::EPB, T = C::THREADS, NPB = C::NPB, KSTR = C::KSTR;
  extern __shared__ double sm[];
  const double *Bs = sm, *Ds = Bs + QS * N, *qs = Ds + QS * N, *ws = qs + QS;
  double *pd = sm + C::BP;
  const int tid = threadIdx.x;
  const long e0 = (long)blockIdx.x * EPB;
  for (int i = tid; i < C::TAB; i += T) sm[i] = A.tab[i];
  for (int i = tid; i < EPB * 24; i += T) {
    const int el = i / 24, r = i - el * 24;
    sm[C::BX + el * C::XU + r] = e0 + el < A.nelem ? A.coords[(e0 + el) * 24 + r] : (double)((r / 3 >> (2 - r % 3)) & 1);
  }
  for (int i = tid; i < EPB * N3; i += T) {              // dFSGlobalToExpanded(INHOMOGENEOUS), fs.c:543-586
    const int el = i / N3, nd = i - el * N3;
    double v0 = 0, v1 = 0, v2 = 0;
    if (e0 + el < A.nelem) {
      const long gnode = (e0 + el) * N3 + nd;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(A.noff + gnode * 32));
      const int idx = A.e2l[gnode];
      const double *src = idx < A.n_own ? A.x + (long)idx * 3 : A.xaux + (long)(idx - A.n_own) * 3;
      v0 = src[0]; v1 = src[1]; v2 = src[2];
    }
    double *U = sm + C::BX + el * C::XU + 36;
    U[nd] = v0; U[N3 + nd] = v1; U[2 * N3 + nd] = v2;
  }
  __syncthreads();
  if (tid < EPB) {                                       // parallelepiped elements: one inverse Jacobian serves every point (and quirk (ii) is void)
    double *X = sm + C::BX + tid * C::XU;
    const bool aff = elem_is_affine(X);
    if (aff) {
      LineGeom lg;
      PointGeom pg;
      line_geom(X, 0.0, 0.0, lg, false);
      point_geom_adj(lg, 0.0, pg);
#pragma unroll
      for (int i = 0; i < 9; i++) X[24 + i] = pg.adj[i];
      X[33] = pg.det; X[34] = pg.rdet;
    }
    X[35] = aff ? 1.0 : 0.0;
  }
  __syncthreads();
  const int lane = tid & 31, warp = tid >> 5;

#pragma unroll 1
  for (int s0 = 0; s0 < P; s0 += SL) {
    const int nl = SL < P - s0 ? SL : P - s0;
    // ---- B: x then y contraction, 3 components ---------------------------------------------------------------------------------
    for (int it = tid; it < (Q1E_DBG(16) ? 0 : EPB * 3 * 2 * nl * N * N); it += T) {
      const int el = it / (3 * 2 * nl * N * N), r = it - el * (3 * 2 * nl * N * N), d = r / (2 * nl * N * N), r1 = r - d * (2 * nl * N * N), ql = r1 / (N * N), jk = r1 - ql * (N * N), qa = 2 * s0 + ql;
      const double *U = sm + C::BX + el * C::XU + 36 + d * N3;
      double a0 = 0, a1 = 0;
#pragma unroll
      for (int i = 0; i < N; i++) { const double u = U[i * N * N + jk]; a0 = fma(Bs[qa * N + i], u, a0); a1 = fma(Ds[qa * N + i], u, a1); }
      double *X0 = sm + C::BD + el * C::DS + d * (QL * N * N);
      X0[r1] = a0;
      X0[3 * QL * N * N + r1] = a1;
    }
    __syncthreads();
#pragma unroll 1
    for (int j0 = 0; j0 < P; j0 += TJ) {
    const int nj = TJ < P - j0 ? TJ : P - j0;             // patch columns in this tile
    // row records of the tile's nodes -> shared memory, asynchronously (consumed by the gather phase G)
    for (int i = tid; i < 2 * C::NREC; i += T) {
      const int idx = i >> 1, el = idx / ((SL + 1) * (TJ + 1) * N), r = idx - el * ((SL + 1) * (TJ + 1) * N), il = r / ((TJ + 1) * N), r2 = r - il * ((TJ + 1) * N), jl = r2 / N, k = r2 - jl * N;
      if (il <= nl && jl <= nj && e0 + el < A.nelem) {
        const long gnode = (e0 + el) * N3 + ((s0 + il) * N + (j0 + jl)) * N + k;
        const unsigned dst = (unsigned)__cvta_generic_to_shared(sm + C::BR + 4 * idx + 2 * (i & 1));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(A.noff + gnode * 32 + 16 * (i & 1)) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (int it = tid; it < (Q1E_DBG(16) ? 0 : EPB * 3 * 2 * nl * 2 * nj * N); it += T) {
      const int el = it / (3 * 2 * nl * 2 * nj * N), r = it - el * (3 * 2 * nl * 2 * nj * N), d = r / (2 * nl * 2 * nj * N), r1 = r - d * (2 * nl * 2 * nj * N);
      const int ql = r1 / (2 * nj * N), r2 = r1 - ql * (2 * nj * N), qbl = r2 / N, k = r2 - qbl * N, qb = 2 * j0 + qbl;
      const double *X0 = sm + C::BD + el * C::DS + d * (QL * N * N) + ql * N * N, *X1 = X0 + 3 * QL * N * N;
      double v = 0, vx = 0, vy = 0;
#pragma unroll
      for (int j = 0; j < N; j++) {
        const double b = Bs[qb * N + j], x0 = X0[j * N + k];
        v = fma(b, x0, v); vx = fma(b, X1[j * N + k], vx); vy = fma(Ds[qb * N + j], x0, vy);
      }
      double *Vd = sm + C::BD + el * C::DS + 6 * QL * N * N + d * (QL * QJ * N) + (ql * QJ + qbl) * N + k;
      Vd[0] = v;
      Vd[3 * QL * QJ * N] = vx;
      Vd[6 * QL * QJ * N] = vy;
    }
    __syncthreads();

    // ---- D: one thread per SPARSE point ------------------------------------------------------------------------------------------
    for (int it = tid; it < (Q1E_DBG(8) ? 0 : EPB * 2 * nl * 2 * nj * QS); it += T) {
      const int el = it / (2 * nl * 2 * nj * QS), r = it - el * (2 * nl * 2 * nj * QS), ql = r / (2 * nj * QS), r2 = r - ql * (2 * nj * QS), qbl = r2 / QS, qc = r2 - qbl * QS;
      if (e0 + el >= A.nelem) continue;
      const int pl = ql >> 1, pjl = qbl >> 1, pj = j0 + pjl, pk = qc >> 1, q = ((ql & 1) << 2) | ((qbl & 1) << 1) | (qc & 1), qa = 2 * s0 + ql, qb = 2 * j0 + qbl;
      const int patch = el * (nl * nj * P) + (pl * nj + pjl) * P + pk;
      const double *X = sm + C::BX + el * C::XU;
      // displacement gradient in reference coordinates: z contraction
      double g[3][3];
#pragma unroll
      for (int d = 0; d < 3; d++) {
        const double *v = sm + C::BD + el * C::DS + 6 * QL * N * N + d * (QL * QJ * N) + (ql * QJ + qbl) * N, *vx = v + 3 * QL * QJ * N, *vy = v + 6 * QL * QJ * N;
        double g0 = 0, g1 = 0, g2 = 0;
#pragma unroll
        for (int k = 0; k < N; k++) {
          const double b = Bs[qc * N + k];
          g0 = fma(b, vx[k], g0); g1 = fma(b, vy[k], g1); g2 = fma(Ds[qc * N + k], v[k], g2);
        }
        g[d][0] = g0; g[d][1] = g1; g[d][2] = g2;
      }
      // geometry at the true point (jacobi.c:383-414) and the inverse Jacobian of element-natural point pt*8+q (quirk (ii))
      PointGeom pg;
      double jq[9];
      if (X[35] != 0.0) {
#pragma unroll
        for (int i = 0; i < 9; i++) { pg.adj[i] = X[24 + i]; jq[i] = X[24 + i] * X[34]; }
        pg.det = X[33]; pg.rdet = X[34];
      } else {
        LineGeom lg;
        line_geom(X, qs[qa], qs[qb], lg, false);
        point_geom_adj(lg, qs[qc], pg);
        const int pt = ((s0 + pl) * P + pj) * P + pk, m = pt * 8 + q, mc = m % QS, mb = (m / QS) % QS, ma = m / (QS * QS);
        PointGeom pq;
        line_geom(X, qs[ma], qs[mb], lg, false);
        point_geom_adj(lg, qs[mc], pq);
#pragma unroll
        for (int i = 0; i < 9; i++) jq[i] = pq.adj[i] * pq.rdet;
      }
      const double jw = ws[qa] * ws[qb] * ws[qc] * pg.det;
      double F[3][3];                                    // elast: F = 1 + Du, ElastPointwiseComputeStore (elast.c:225); stokes: Du
#pragma unroll
      for (int d = 0; d < 3; d++)
#pragma unroll
        for (int x = 0; x < 3; x++) F[d][x] = (KIND == Q1_ELAST && d == x ? 1.0 : 0.0) + pg.rdet * (g[d][0] * pg.adj[x] + g[d][1] * pg.adj[3 + x] + g[d][2] * pg.adj[6 + x]);
      double *o = pd + patch;                            // pd[(c*8+q)*NPB + patch]
      constexpr int ST = 8 * NPB;
      o += q * NPB;
      o[0] = jw;
      constexpr int s6a[6] = {0, 1, 2, 0, 0, 1}, s6b[6] = {0, 1, 2, 1, 2, 2};                     // [00,11,22,01,02,12]
#pragma unroll
      for (int i = 0; i < 6; i++) {
        const int a = s6a[i], b = s6b[i];
        o[(16 + i) * ST] = jq[a * 3] * jq[b * 3] + jq[a * 3 + 1] * jq[b * 3 + 1] + jq[a * 3 + 2] * jq[b * 3 + 2];      // (jq jq^T)[a][b]
      }
      if (KIND == Q1_ELAST) {
#pragma unroll
        for (int f = 0; f < 3; f++)
#pragma unroll
          for (int k = 0; k < 3; k++) o[(1 + f * 3 + k) * ST] = jq[k * 3] * F[f][0] + jq[k * 3 + 1] * F[f][1] + jq[k * 3 + 2] * F[f][2];      // g_f = jq F_f
        double S[6];
        {
          double E[6];                                     // E = (F^T F - 1)/2, S = lambda tr(E) 1 + 2 mu E (elast.c:227-241)
#pragma unroll
          for (int i = 0; i < 6; i++) E[i] = 0.5 * (F[0][s6a[i]] * F[0][s6b[i]] + F[1][s6a[i]] * F[1][s6b[i]] + F[2][s6a[i]] * F[2][s6b[i]] - (i < 3 ? 1.0 : 0.0));
          const double tr = E[0] + E[1] + E[2];
#pragma unroll
          for (int i = 0; i < 6; i++) S[i] = (i < 3 ? A.c0 * tr : 0.0) + 2 * A.c1 * E[i];
        }
        double Sj[3][3];                                   // S jq^T: Sj[x][l] = sum_b S[x][b] jq[l][b]
        const double Sf[3][3] = {{S[0], S[3], S[4]}, {S[3], S[1], S[5]}, {S[4], S[5], S[2]}};
#pragma unroll
        for (int x = 0; x < 3; x++)
#pragma unroll
          for (int l = 0; l < 3; l++) Sj[x][l] = Sf[x][0] * jq[l * 3] + Sf[x][1] * jq[l * 3 + 1] + Sf[x][2] * jq[l * 3 + 2];
#pragma unroll
        for (int i = 0; i < 6; i++) {
          const int a = s6a[i], b = s6b[i];
          o[(10 + i) * ST] = F[a][0] * F[b][0] + F[a][1] * F[b][1] + F[a][2] * F[b][2];             // F_a . F_b
          o[(22 + i) * ST] = jw * (jq[a * 3] * Sj[0][b] + jq[a * 3 + 1] * Sj[1][b] + jq[a * 3 + 2] * Sj[2][b]);         // jw (jq S jq^T)[a][b]
        }
      } else {
        // StokesPointwiseComputeStore (stokes.c:447-453) at the SPARSE point: Dw = sym(Du), gamma, eta, eta'
        double Dw[3][3];
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
          for (int b = 0; b < 3; b++) Dw[a][b] = 0.5 * (F[a][b] + F[b][a]);
        double colon = 0;
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
          for (int b = 0; b < 3; b++) colon += Dw[a][b] * Dw[a][b];
        const double gamma = A.c1 + 0.5 * colon, eta = A.c0 * (A.c2 != 0 ? exp(A.c2 * log(gamma)) : 1.0), deta = A.c2 * eta / gamma;
        o[10 * ST] = 0.5 * eta * jw;                       // A2
        o[11 * ST] = deta * jw;                            // A3
        o[12 * ST] = A.c3 * jw;                            // A1 = lambda jw, StokesLambda (stokes.c:455-458)
#pragma unroll
        for (int f = 0; f < 3; f++)
#pragma unroll
          for (int k = 0; k < 3; k++) {
            o[(1 + f * 3 + k) * ST] = jq[k * 3 + f];                                                                   // c_f = column f of jq
            o[q1s_h(f, k) * ST] = jq[k * 3] * Dw[f][0] + jq[k * 3 + 1] * Dw[f][1] + jq[k * 3 + 2] * Dw[f][2];          // h_f = jq Dw_f
          }
      }
    }
    __syncthreads();

    // ---- C: one thread per (patch, unit); threads [0, 3 npatch) diagonal units, [3 npatch, 9 npatch) off-diagonal halves ---------
    {
      const int npatch = EPB * nl * nj * P;
      constexpr int ST = 8 * NPB;
      if (tid < 3 * npatch) {
        const int fi = tid / npatch, patch = tid - fi * npatch;
        if (e0 + patch / (nl * nj * P) < A.nelem) {
          double K[36];
#pragma unroll
          for (int i = 0; i < 36; i++) K[i] = 0;
          if (!Q1E_DBG(2))
#pragma unroll Q1E_QU
          for (int q = 0; q < 8; q++) {
            const double *o = pd + q * NPB + patch;
            const double g0 = o[(1 + fi * 3) * ST], g1 = o[(2 + fi * 3) * ST], g2 = o[(3 + fi * 3) * ST];
            double T00, T11, T22, T01, T02, T12;
            if (KIND == Q1_ELAST) {
              const double jw = o[0], ab = (A.c0 + A.c1) * jw, cm = A.c1 * jw * o[(10 + fi) * ST];
              T00 = ab * g0 * g0 + cm * o[16 * ST] + o[22 * ST]; T11 = ab * g1 * g1 + cm * o[17 * ST] + o[23 * ST]; T22 = ab * g2 * g2 + cm * o[18 * ST] + o[24 * ST];
              T01 = ab * g0 * g1 + cm * o[19 * ST] + o[25 * ST]; T02 = ab * g0 * g2 + cm * o[20 * ST] + o[26 * ST]; T12 = ab * g1 * g2 + cm * o[21 * ST] + o[27 * ST];
            } else {
              // T = (A1 + A2) c c^T + A3 h h^T + A2 jq jq^T  (StokesPointwiseJacobian_A on sym(e_f (x) G_j), stokes.c:491-496,700-719)
              const double a2 = o[10 * ST], a3 = o[11 * ST], ab = o[12 * ST] + a2;
              double h[3];
#pragma unroll
              for (int k = 0; k < 3; k++) h[k] = fi == 0 ? o[q1s_h(0, k) * ST] : fi == 1 ? o[q1s_h(1, k) * ST] : o[q1s_h(2, k) * ST];
              T00 = ab * g0 * g0 + a3 * h[0] * h[0] + a2 * o[16 * ST]; T11 = ab * g1 * g1 + a3 * h[1] * h[1] + a2 * o[17 * ST]; T22 = ab * g2 * g2 + a3 * h[2] * h[2] + a2 * o[18 * ST];
              T01 = ab * g0 * g1 + a3 * h[0] * h[1] + a2 * o[19 * ST]; T02 = ab * g0 * g2 + a3 * h[0] * h[2] + a2 * o[20 * ST]; T12 = ab * g1 * g2 + a3 * h[1] * h[2] + a2 * o[21 * ST];
            }
            double yz[4], xz[4], xy[4];
            q1_patch_basis(q >> 2, (q >> 1) & 1, q & 1, yz, xz, xy);
#pragma unroll
            for (int j = 0; j < 8; j++) {
              const double dj0 = Q1_D0(j, yz), dj1 = Q1_D1(j, xz), dj2 = Q1_D2(j, xy);
              const double w0 = T00 * dj0 + T01 * dj1 + T02 * dj2, w1 = T01 * dj0 + T11 * dj1 + T12 * dj2, w2 = T02 * dj0 + T12 * dj1 + T22 * dj2;
#pragma unroll
              for (int i = 0; i <= j; i++) {
                double acc = K[q1_sym(i, j)];
                acc = fma(Q1_D0(i, yz), w0, acc); acc = fma(Q1_D1(i, xz), w1, acc); acc = fma(Q1_D2(i, xy), w2, acc);
                K[q1_sym(i, j)] = acc;
              }
            }
          }
          double *Kp = sm + C::BK + patch * KSTR + fi * C::KBLK;
#pragma unroll
          for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < 8; j++) Kp[i * 8 + j] = K[q1_sym(i, j)];
        }
      } else if (tid < 9 * npatch && !A.reduced) {
        const int u = (tid - 3 * npatch) / npatch, patch = tid - 3 * npatch - u * npatch, pr = u >> 1, h = u & 1;
        const int fi = pr == 2 ? 1 : 0, fj = pr == 0 ? 1 : 2;                                      // (0,1) (0,2) (1,2)
        if (e0 + patch / (nl * nj * P) < A.nelem) {
          double K[32];
#pragma unroll
          for (int i = 0; i < 32; i++) K[i] = 0;
          if (!Q1E_DBG(2))
#pragma unroll Q1E_QU
          for (int q = 0; q < 8; q++) {
            const double *o = pd + q * NPB + patch;
            const double gi0 = o[(1 + fi * 3) * ST], gi1 = o[(2 + fi * 3) * ST], gi2 = o[(3 + fi * 3) * ST];
            const double gj0 = o[(1 + fj * 3) * ST], gj1 = o[(2 + fj * 3) * ST], gj2 = o[(3 + fj * 3) * ST];
            // T[k][l] = la gi[k] gj[l] + mu gj[k] gi[l] + (elast) cm M[k][l] / (stokes) a3 hi[k] hj[l]   (k: test side, l: trial side)
            double la, mu, T00, T11, T22, T01, T10, T02, T20, T12, T21;
            if (KIND == Q1_ELAST) {
              const double jw = o[0];
              la = A.c0 * jw; mu = A.c1 * jw;
              const double cm = mu * o[(13 + pr) * ST];
              T00 = cm * o[16 * ST]; T11 = cm * o[17 * ST]; T22 = cm * o[18 * ST];
              T01 = T10 = cm * o[19 * ST]; T02 = T20 = cm * o[20 * ST]; T12 = T21 = cm * o[21 * ST];
            } else {
              la = o[12 * ST]; mu = o[10 * ST];
              const double a3 = o[11 * ST];
              double hi[3], hj[3];
#pragma unroll
              for (int k = 0; k < 3; k++) {
                hi[k] = fi == 0 ? o[q1s_h(0, k) * ST] : o[q1s_h(1, k) * ST];
                hj[k] = fj == 1 ? o[q1s_h(1, k) * ST] : o[q1s_h(2, k) * ST];
              }
              T00 = a3 * hi[0] * hj[0]; T11 = a3 * hi[1] * hj[1]; T22 = a3 * hi[2] * hj[2];
              T01 = a3 * hi[0] * hj[1]; T10 = a3 * hi[1] * hj[0]; T02 = a3 * hi[0] * hj[2]; T20 = a3 * hi[2] * hj[0]; T12 = a3 * hi[1] * hj[2]; T21 = a3 * hi[2] * hj[1];
            }
            T00 += (la + mu) * gi0 * gj0; T11 += (la + mu) * gi1 * gj1; T22 += (la + mu) * gi2 * gj2;
            T01 += la * gi0 * gj1 + mu * gj0 * gi1; T10 += la * gi1 * gj0 + mu * gj1 * gi0;
            T02 += la * gi0 * gj2 + mu * gj0 * gi2; T20 += la * gi2 * gj0 + mu * gj2 * gi0;
            T12 += la * gi1 * gj2 + mu * gj1 * gi2; T21 += la * gi2 * gj1 + mu * gj2 * gi1;
            double yz[4], xz[4], xy[4];
            q1_patch_basis(q >> 2, (q >> 1) & 1, q & 1, yz, xz, xy);
            // the four test functions of this half: corners 4h..4h+3, i.e. b0 = h
            double t0[4], t1[4], t2[4];                                                             // d_i for i = 4h + m
#pragma unroll
            for (int m = 0; m < 4; m++) {
              t0[m] = h ? yz[m] : -yz[m];
              const double xzm = h ? xz[2 + (m & 1)] : xz[m & 1], xym = h ? xy[2 + (m >> 1)] : xy[m >> 1];      // static indices, runtime select
              t1[m] = (m >> 1) ? xzm : -xzm;
              t2[m] = (m & 1) ? xym : -xym;
            }
#pragma unroll
            for (int j = 0; j < 8; j++) {
              const double dj0 = Q1_D0(j, yz), dj1 = Q1_D1(j, xz), dj2 = Q1_D2(j, xy);
              const double w0 = T00 * dj0 + T01 * dj1 + T02 * dj2, w1 = T10 * dj0 + T11 * dj1 + T12 * dj2, w2 = T20 * dj0 + T21 * dj1 + T22 * dj2;
#pragma unroll
              for (int m = 0; m < 4; m++) {
                double acc = K[m * 8 + j];
                acc = fma(t0[m], w0, acc); acc = fma(t1[m], w1, acc); acc = fma(t2[m], w2, acc);
                K[m * 8 + j] = acc;
              }
            }
          }
          double *Kp = sm + C::BK + patch * KSTR + (3 + pr) * C::KBLK + h * 32;
#pragma unroll
          for (int i = 0; i < 32; i++) Kp[i] = K[i];
        }
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    // ---- G: one thread per (row node, fi, fj): 27 lanes = 3 nodes per warp pass (dFSMatSetValuesBlockedExpanded, fs.c:826-930) -----
    {
      const int sub = lane / 9, f = lane - sub * 9, fi = f / 3, fj = f - fi * 3;
      // block of the packed patch matrix and its row/column strides for this component pair (transposed read for fi > fj)
      const int lo = fi < fj ? fi : fj, hi = fi < fj ? fj : fi;
      const int blk = fi == fj ? fi : 3 + (lo == 0 ? hi - 1 : 2);
      const int sr = fi <= fj ? 8 : 1, sc = fi <= fj ? 1 : 8;
      for (int base = warp * 3; base < (Q1E_DBG(4) ? 0 : EPB * (SL + 1) * (TJ + 1) * N); base += C::NW * 3) {
        const int idx = base + sub;
        const int el = idx / ((SL + 1) * (TJ + 1) * N), r = idx - el * ((SL + 1) * (TJ + 1) * N), il = r / ((TJ + 1) * N), r2 = r - il * ((TJ + 1) * N), jl = r2 / N, k = r2 - jl * N;
        uint4 lo4 = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu), hi4 = lo4;
        if (lane < 27 && idx < EPB * (SL + 1) * (TJ + 1) * N && il <= nl && jl <= nj && e0 + el < A.nelem) {
          lo4 = reinterpret_cast<const uint4 *>(sm + C::BR)[idx * 2];
          hi4 = reinterpret_cast<const uint4 *>(sm + C::BR)[idx * 2 + 1];
        }
        const int rp = (int)hi4.w;                         // CSR start of the node's block row, -1: no row
        if (rp < 0 || (A.reduced && fi != fj)) continue;       // reduced: the off-diagonal units were not computed
        unsigned W[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
        double sums[27];
#pragma unroll
        for (int t = 0; t < 27; t++) sums[t] = 0;
        const double *Kel = sm + C::BK + (el * (nl * nj * P) + (il * nj + jl) * P + k) * KSTR + blk * C::KBLK;
#pragma unroll
        for (int abc = 0; abc < 8; abc++) {
          const int a = abc >> 2, b = (abc >> 1) & 1, c = abc & 1;          // the node is corner (a,b,c) of patch (il-a, j-b, k-c)
          if ((unsigned)(il - a) < (unsigned)nl && (unsigned)(jl - b) < (unsigned)nj && (unsigned)(k - c) < (unsigned)P) {
            const double *Kp = Kel - ((a * nj + b) * P + c) * KSTR + abc * sr;
#pragma unroll
            for (int col = 0; col < 8; col++) {
              const int ca = col >> 2, cb = (col >> 1) & 1, cc = col & 1;
              sums[(ca - a + 1) * 9 + (cb - b + 1) * 3 + (cc - c + 1)] += Kp[col * sc];
            }
          }
        }
        // neighbour entries whose patches belong to another slab / column tile are written by that pass: mark them dropped (0xFF)
        // in the packed record, so that one test per entry is left
        if (SL < P) {
          if (il == 0) { W[0] = W[1] = 0xFFFFFFFFu; W[2] |= 0x000000FFu; }                       // t in [0, 9): layer below
          if (il == nl) { W[4] |= 0xFFFF0000u; W[5] = 0xFFFFFFFFu; W[6] |= 0x00FFFFFFu; }         // t in [18, 27): layer above
        }
        if (TJ < P) {
#pragma unroll
          for (int w = 0; w < 7; w++) {
            if (jl == 0) W[w] |= q1_dropmask(w, 0);
            if (jl == nj) W[w] |= q1_dropmask(w, 2);
          }
        }
        if (Q1E_DBG(1)) continue;
        double *vb = A.vals + ((long)rp * 9 + f);           // the nine lanes of a node: nine consecutive doubles of every block
#pragma unroll
        for (int t = 0; t < 27; t++) {
          const unsigned rel = (W[t >> 2] >> (8 * (t & 3))) & 0xFFu;
          q1_red_unless_dropped(vb + rel * 9, sums[t], rel);
        }
      }
    }
    __syncthreads();
    }                                                       // j tiles
  }
}

}  // namespace dohp
