// Second-generation tensor stages (sm_100a, FP64 on the DFMA pipe), used by the fused operator kernels.
//
// Two changes against tensor_stages.cuh, both driven by the round-1 ncu captures (profiles/r1e_*: shared-memory pipe 75-88 % busy
// with 23-38 % of the wavefronts bank conflicts; FP64 pipe 31-58 %):
//  (1) EVEN-ODD 1-D contractions.  Every 1-D table Dohp builds is (anti)symmetric under the reflection of the reference interval:
//      interp[Q-1-a][P-1-i] = interp[a][i] (GLL nodes and Gauss points are symmetric sets, src/jacobi/impls/tensor/tensor.c:44-68),
//      and the collocation derivative changes sign.  Splitting a line into its even and odd parts turns the n x n product into two
//      (n/2) x (n/2) products: n^2/2 + 2n FP64 instructions instead of n^2 (n >= 5; the host verifies the symmetry of the tables it
//      is given and refuses others).  The same split serves the transposed products.
//  (2) a CONFLICT-FREE shared-memory scheme: element layout a*SA + b*SB + c with strides chosen per n, the z lines in natural thread
//      order, and HOST-BUILT thread maps for the x and y passes (any thread may own any line of the block: the map deals the lines
//      of the block to half-warps so that the 16 lanes of a half-warp hit 16 different 64-bit banks; smem_map.cpp).
#pragma once
#include "tensor_stages.cuh"

namespace dohp {

template <int NQ> struct Lay2;
#define DOHP_LAY2(N, A_, B_)                                                                                                     \
  template <> struct Lay2<N> {                                                                                                     \
    static constexpr int SA = A_, SB = B_, CS = (N - 1) * (A_ + B_ + 1) + 1;                                                       \
    DI static int at(int a, int b, int c) { return a * SA + b * SB + c; }                                                          \
  };
DOHP_LAY2(2, 8, 3) DOHP_LAY2(3, 9, 3) DOHP_LAY2(4, 17, 4) DOHP_LAY2(5, 25, 5) DOHP_LAY2(6, 42, 7) DOHP_LAY2(7, 49, 7)
DOHP_LAY2(8, 72, 9) DOHP_LAY2(9, 81, 9) DOHP_LAY2(10, 110, 11) DOHP_LAY2(11, 121, 11)
#undef DOHP_LAY2

// thread-map entry: element slot << 16 | first line coordinate << 8 | second line coordinate ; slot >= EPB means "no line"
DI void map_decode(int m, int &le, int &u, int &v) { le = m >> 16; u = (m >> 8) & 255; v = m & 255; }

constexpr bool use_eo(int n) { return n >= 5; }

// y[NR] = A x[NC] and y[NC] = A^T x[NR] for a table with A[NR-1-r][NC-1-c] = SIGN * A[r][c]
template <int NR, int NC, int SIGN, bool EO> struct Mat1D;

template <int NR, int NC, int SIGN> struct Mat1D<NR, NC, SIGN, false> {
  double a[NR * NC];
  DI void fwd(const double (&x)[NC], double (&y)[NR]) const
  {
#pragma unroll
    for (int r = 0; r < NR; r++) {
      double s = 0;
#pragma unroll
      for (int c = 0; c < NC; c++) s = fma(a[r * NC + c], x[c], s);
      y[r] = s;
    }
  }
  DI void bwd(const double (&x)[NR], double (&y)[NC]) const
  {
#pragma unroll
    for (int c = 0; c < NC; c++) {
      double s = 0;
#pragma unroll
      for (int r = 0; r < NR; r++) s = fma(a[r * NC + c], x[r], s);
      y[c] = s;
    }
  }
};

template <int NR, int NC, int SIGN> struct Mat1D<NR, NC, SIGN, true> {
  static constexpr int hr = NR / 2, hc = NC / 2, HR = (NR + 1) / 2, HC = (NC + 1) / 2;
  static constexpr bool mr = NR & 1, mc = NC & 1;
  double e[HR * HC], o[HR * HC];      // e = (A[r][c] + A[r][NC-1-c]) / 2, o = (A[r][c] - A[r][NC-1-c]) / 2 ; middle column: e = A[r][hc], o = 0
  DI void fwd(const double (&x)[NC], double (&y)[NR]) const
  {
    double xe[HC], xo[hc > 0 ? hc : 1];
#pragma unroll
    for (int c = 0; c < hc; c++) { xe[c] = x[c] + x[NC - 1 - c]; xo[c] = x[c] - x[NC - 1 - c]; }
    if (mc) xe[hc] = x[hc];
#pragma unroll
    for (int r = 0; r < hr; r++) {
      double E = 0, O = 0;
#pragma unroll
      for (int c = 0; c < HC; c++) E = fma(e[r * HC + c], xe[c], E);
#pragma unroll
      for (int c = 0; c < hc; c++) O = fma(o[r * HC + c], xo[c], O);
      y[r] = E + O;
      y[NR - 1 - r] = SIGN > 0 ? E - O : O - E;
    }
    if (mr) {
      double s = 0;
      if (SIGN > 0) {
#pragma unroll
        for (int c = 0; c < HC; c++) s = fma(e[hr * HC + c], xe[c], s);
      } else {
#pragma unroll
        for (int c = 0; c < hc; c++) s = fma(o[hr * HC + c], xo[c], s);
      }
      y[hr] = s;
    }
  }
  DI void bwd(const double (&x)[NR], double (&y)[NC]) const
  {
    double xe[hr > 0 ? hr : 1], xo[hr > 0 ? hr : 1];
#pragma unroll
    for (int r = 0; r < hr; r++) { xe[r] = x[r] + x[NR - 1 - r]; xo[r] = x[r] - x[NR - 1 - r]; }
    const double xm = mr ? x[hr] : 0.0;
#pragma unroll
    for (int c = 0; c < hc; c++) {
      double E = 0, O = 0;
      if (SIGN > 0) {
#pragma unroll
        for (int r = 0; r < hr; r++) { E = fma(e[r * HC + c], xe[r], E); O = fma(o[r * HC + c], xo[r], O); }
        if (mr) E = fma(e[hr * HC + c], xm, E);
      } else {
#pragma unroll
        for (int r = 0; r < hr; r++) { E = fma(e[r * HC + c], xo[r], E); O = fma(o[r * HC + c], xe[r], O); }
        if (mr) O = fma(o[hr * HC + c], xm, O);
      }
      y[c] = E + O;
      y[NC - 1 - c] = E - O;
    }
    if (mc) {
      double s = 0;
      if (SIGN > 0) {
#pragma unroll
        for (int r = 0; r < hr; r++) s = fma(e[r * HC + hc], xe[r], s);
        if (mr) s = fma(e[hr * HC + hc], xm, s);
      } else {
#pragma unroll
        for (int r = 0; r < hr; r++) s = fma(e[r * HC + hc], xo[r], s);
      }
      y[hc] = s;
    }
  }
};

// host: fill from the full row-major table; returns the largest violation of the assumed symmetry
template <int NR, int NC, int SIGN> inline double mat1d_fill(Mat1D<NR, NC, SIGN, false> &m, const double *A)
{
  for (int i = 0; i < NR * NC; i++) m.a[i] = A[i];
  return 0;
}
template <int NR, int NC, int SIGN> inline double mat1d_fill(Mat1D<NR, NC, SIGN, true> &m, const double *A)
{
  using M = Mat1D<NR, NC, SIGN, true>;
  double viol = 0;
  for (int r = 0; r < NR; r++)
    for (int c = 0; c < NC; c++) {
      const double d = A[(NR - 1 - r) * NC + (NC - 1 - c)] - SIGN * A[r * NC + c];
      viol = d > viol ? d : (-d > viol ? -d : viol);
    }
  for (int r = 0; r < M::HR; r++)
    for (int c = 0; c < M::HC; c++) {
      if (c < M::hc) {
        m.e[r * M::HC + c] = 0.5 * (A[r * NC + c] + A[r * NC + NC - 1 - c]);
        m.o[r * M::HC + c] = 0.5 * (A[r * NC + c] - A[r * NC + NC - 1 - c]);
      } else {
        m.e[r * M::HC + c] = A[r * NC + c];
        m.o[r * M::HC + c] = 0;
      }
    }
  return viol;
}

template <int NP, int NQ> using InterpMat = Mat1D<NQ, NP, +1, use_eo(NQ)>;      // interp[a][i]
template <int NQ> using DerivMat = Mat1D<NQ, NQ, -1, use_eo(NQ)>;              // collocation derivative Dt[a][m] = l_m'(q_a)
template <int NQ> struct RuleTab2 { DerivMat<NQ> D; double w[NQ], q[NQ]; };

// ---- stages on the Lay2 layout; the caller passes the coordinates of the line this thread owns in the pass ---------------------------
template <int NP, int NQ> DI void fwd_x2(const InterpMat<NP, NQ> &B, const double (&r)[NP], double *A, int j, int k)
{
  double y[NQ];
  B.fwd(r, y);
#pragma unroll
  for (int a = 0; a < NQ; a++) A[Lay2<NQ>::at(a, j, k)] = y[a];
}
template <int NP, int NQ> DI void fwd_y2(const InterpMat<NP, NQ> &B, double *A, int a, int k)
{
  double r[NP], y[NQ];
#pragma unroll
  for (int j = 0; j < NP; j++) r[j] = A[Lay2<NQ>::at(a, j, k)];
  B.fwd(r, y);
#pragma unroll
  for (int b = 0; b < NQ; b++) A[Lay2<NQ>::at(a, b, k)] = y[b];
}
template <int NP, int NQ> DI void fwd_z2(const InterpMat<NP, NQ> &B, const double *A, int a, int b, double (&u)[NQ])
{
  double r[NP];
#pragma unroll
  for (int k = 0; k < NP; k++) r[k] = A[Lay2<NQ>::at(a, b, k)];
  B.fwd(r, u);
}
template <int NQ> DI void put_z2(double *A, int a, int b, const double (&u)[NQ])
{
#pragma unroll
  for (int c = 0; c < NQ; c++) A[Lay2<NQ>::at(a, b, c)] = u[c];
}
template <int NQ> DI void der_x2(const DerivMat<NQ> &D, const double *A, double *G, int b, int c)
{
  double r[NQ], y[NQ];
#pragma unroll
  for (int m = 0; m < NQ; m++) r[m] = A[Lay2<NQ>::at(m, b, c)];
  D.fwd(r, y);
#pragma unroll
  for (int a = 0; a < NQ; a++) G[Lay2<NQ>::at(a, b, c)] = y[a];
}
template <int NQ> DI void der_y2(const DerivMat<NQ> &D, const double *A, double *G, int a, int c)
{
  double r[NQ], y[NQ];
#pragma unroll
  for (int m = 0; m < NQ; m++) r[m] = A[Lay2<NQ>::at(a, m, c)];
  D.fwd(r, y);
#pragma unroll
  for (int b = 0; b < NQ; b++) G[Lay2<NQ>::at(a, b, c)] = y[b];
}
template <int NQ> DI void der_xT2(const DerivMat<NQ> &D, double *G, int b, int c)
{
  double r[NQ], y[NQ];
#pragma unroll
  for (int a = 0; a < NQ; a++) r[a] = G[Lay2<NQ>::at(a, b, c)];
  D.bwd(r, y);
#pragma unroll
  for (int m = 0; m < NQ; m++) G[Lay2<NQ>::at(m, b, c)] = y[m];
}
template <int NQ> DI void der_yT2(const DerivMat<NQ> &D, double *G, int a, int c)
{
  double r[NQ], y[NQ];
#pragma unroll
  for (int b = 0; b < NQ; b++) r[b] = G[Lay2<NQ>::at(a, b, c)];
  D.bwd(r, y);
#pragma unroll
  for (int m = 0; m < NQ; m++) G[Lay2<NQ>::at(a, m, c)] = y[m];
}
template <int NQ> DI void der_zT_add2(const DerivMat<NQ> &D, const double (&g)[NQ], double (&h)[NQ])
{
  double y[NQ];
  D.bwd(g, y);
#pragma unroll
  for (int m = 0; m < NQ; m++) h[m] += y[m];
}
template <int NP, int NQ> DI void bwd_z2(const InterpMat<NP, NQ> &B, const double (&g)[NQ], double *A, int a, int b)
{
  double y[NP];
  B.bwd(g, y);
#pragma unroll
  for (int k = 0; k < NP; k++) A[Lay2<NQ>::at(a, b, k)] = y[k];
}
template <int NP, int NQ> DI void bwd_y2(const InterpMat<NP, NQ> &B, double *A, int a, int k)
{
  double r[NQ], y[NP];
#pragma unroll
  for (int b = 0; b < NQ; b++) r[b] = A[Lay2<NQ>::at(a, b, k)];
  B.bwd(r, y);
#pragma unroll
  for (int j = 0; j < NP; j++) A[Lay2<NQ>::at(a, j, k)] = y[j];
}
template <int NP, int NQ> DI void bwd_x2(const InterpMat<NP, NQ> &B, const double *A, int j, int k, double (&o)[NP])
{
  double r[NQ];
#pragma unroll
  for (int a = 0; a < NQ; a++) r[a] = A[Lay2<NQ>::at(a, j, k)];
  B.bwd(r, o);
}

// ---- stash staging: per-thread ring in shared memory filled by cp.async (LDGSTS) ----------------------------------------------------------
// The Jacobian-mode kernels read what the residual evaluation stored per quadrature point (dRulesetIterator stash,
// src/fs/interface/fsrulesetit.c:779-798).  Every thread consumes only the points of its own z line, in order, so each thread runs a
// private ring of RING points through shared memory: the first RING points are requested at kernel start (they arrive while the
// contractions run), point c + RING is requested right after point c has been read.  No registers are held and no barrier is needed
// (producer = consumer).  Round 1 stalled up to 51 % of the issue slots on these loads (profiles/r1e_stalls.txt, long_sb).
DI void cp_async8(unsigned dst, const double *src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory"); }
DI void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
DI void cp_wait_pending(int n)      // returns when at most n of the most recently committed groups are still in flight
{
  switch (n) {
  case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
  case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
  case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
  case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
  case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
  case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
  case 6: asm volatile("cp.async.wait_group 6;" ::: "memory"); break;
  case 7: asm volatile("cp.async.wait_group 7;" ::: "memory"); break;
  case 8: asm volatile("cp.async.wait_group 8;" ::: "memory"); break;
  case 9: asm volatile("cp.async.wait_group 9;" ::: "memory"); break;
  default: asm volatile("cp.async.wait_group 10;" ::: "memory"); break;
  }
}
// ring of RING points x npl planes per thread; thread-private column `t` of [slot][plane][T]
template <int NQ, int T, int NPLMAX, int RING> struct StashRing {
  unsigned sbase;                 // shared-window address of this thread's column
  const double *stg;              // generic pointer to the same
  const double *src;              // stash + (e*NQ)*T + t
  long splane;
  int npl;                        // planes used
  DI void init(double *column, const double *stash, long e, int t, long splane_, int npl_)
  {
    stg = column; sbase = (unsigned)__cvta_generic_to_shared(column); src = stash + (e * NQ) * T + t; splane = splane_; npl = npl_;
  }
  DI void request(int c) const
  {
#pragma unroll
    for (int k = 0; k < NPLMAX; k++)
      if (k < npl) cp_async8(sbase + (unsigned)((((c % RING) * npl + k) * T) * 8), src + k * splane + (long)c * T);
    cp_commit();
  }
  DI void prime() const
  {
#pragma unroll
    for (int c = 0; c < (RING < NQ ? RING : NQ); c++) request(c);
  }
  DI void wait(int c) const { cp_wait_pending((RING - 1) < (NQ - 1 - c) ? (RING - 1) : (NQ - 1 - c)); }
  DI double get(int c, int k) const { return stg[((c % RING) * npl + k) * T]; }
  DI void refill(int c) const { if (c + RING < NQ) request(c + RING); }
};

// ---- per-element geometry of an AFFINE element (a parallelepiped: the Jacobian is constant) --------------------------------------------
// geo[0..8] = adj (det J * J^-1, row k = reference direction), geo[9] = det, geo[10] = 1/det,
// geo[11..16] = M = adj adj^T / det, symmetric [00,01,02,11,12,22]: for the linear Poisson operator the whole pointwise step is
// Dv_ref = w_q M Du_ref (TensorRuleMapping, the identity physics and TensorRuleMappingTranspose folded, efstopo.c:557-623).
#define DOHP_NGEO 17
DI void affine_geometry(const double *X, double *geo)
{
  LineGeom lg;
  line_geom(X, 0.0, 0.0, lg, false);
  PointGeom pg;
  point_geom_adj(lg, 0.0, pg);
#pragma unroll
  for (int i = 0; i < 9; i++) geo[i] = pg.adj[i];
  geo[9] = pg.det;
  geo[10] = pg.rdet;
  int n = 11;
#pragma unroll
  for (int k = 0; k < 3; k++)
#pragma unroll
    for (int l = k; l < 3; l++) geo[n++] = (pg.adj[k * 3] * pg.adj[l * 3] + pg.adj[k * 3 + 1] * pg.adj[l * 3 + 1] + pg.adj[k * 3 + 2] * pg.adj[l * 3 + 2]) * pg.rdet;
}

}  // namespace dohp
