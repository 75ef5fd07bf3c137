// Sum-factorised tensor-product stages shared by every element kernel (sm_100a, FP64 on the CUDA-core DFMA pipe).
//
// What the reference does per element (src/jacobi/impls/tensor/inlinetmulthex.h:19-100 driven by dEFSApply_Tensor_Hex,
// src/jacobi/impls/tensor/efstopo.c:368-423): INTERP = (B (x) B (x) B) u ; GRAD = three such products with the
// derivative table in one direction (18 n^4 flops) ; the transposes likewise.  What these stages do instead:
//   * interpolate once to the quadrature points (3 contractions), then differentiate BY COLLOCATION on the rule's own
//     points (3 contractions with the Q x Q matrix Dt = l_b'(q_a)) — mathematically identical for the polynomial
//     bases Dohp builds (deriv = interp . D_gll, src/jacobi/impls/tensor/tensor.c:62-68) and 12 n^4 instead of 18 n^4;
//   * every 1-D table entry is an immediate constant-bank operand of the DFMA: the tables travel as __grid_constant__
//     kernel parameters and all loops are unrolled at compile time, so a contraction is nothing but
//     "n loads of a column -> n^2 DFMA -> n stores";
//   * one thread owns one grid line (column) of the element; between directions the element is re-laid out through
//     shared memory; the z line stays in registers across the pointwise physics.
// Shared-memory element layout: value (a,b,c) at a*PS + b*SZ + c.
#pragma once
#include <cuda_runtime.h>

#ifndef DI
#define DI __device__ __forceinline__
#endif

namespace dohp {

template <int NQ> struct Lay {
  static constexpr int SZ = NQ | 1;              // odd line stride: the (a,b)-thread pattern t*SZ+c is conflict free
  static constexpr int PS = NQ * SZ;             // plane stride
  static constexpr int CS = NQ * PS;             // one component of one element
  DI static int at(int a, int b, int c) { return a * PS + b * SZ + c; }
};

// 1-D tables of one (P,Q) pair, passed by value inside kernel parameters
template <int NP, int NQ> struct BasisTab { double B[NQ * NP]; };          // interp[a][i]
template <int NQ> struct RuleTab { double Dt[NQ * NQ]; double w[NQ]; double q[NQ]; };   // collocation derivative, weights, points

// ---- forward: nodes (i,j,k) -> points (a,b,c) -------------------------------------------------------------------------
// x: thread (j,k), column r[i] in registers (just gathered from HBM) -> A[a][j][k]
template <int NP, int NQ> DI void fwd_x(const double (&B)[NQ * NP], const double (&r)[NP], double *A, int j, int k)
{
#pragma unroll
  for (int a = 0; a < NQ; a++) {
    double s = 0;
#pragma unroll
    for (int i = 0; i < NP; i++) s = fma(B[a * NP + i], r[i], s);
    A[Lay<NQ>::at(a, j, k)] = s;
  }
}
// y: thread (a,k), in place on its own line
template <int NP, int NQ> DI void fwd_y(const double (&B)[NQ * NP], double *A, int a, int k)
{
  double r[NP];
#pragma unroll
  for (int j = 0; j < NP; j++) r[j] = A[Lay<NQ>::at(a, j, k)];
#pragma unroll
  for (int b = 0; b < NQ; b++) {
    double s = 0;
#pragma unroll
    for (int j = 0; j < NP; j++) s = fma(B[b * NP + j], r[j], s);
    A[Lay<NQ>::at(a, b, k)] = s;
  }
}
// z: thread (a,b), result stays in registers
template <int NP, int NQ> DI void fwd_z(const double (&B)[NQ * NP], const double *A, int a, int b, double (&u)[NQ])
{
  double r[NP];
#pragma unroll
  for (int k = 0; k < NP; k++) r[k] = A[Lay<NQ>::at(a, b, k)];
#pragma unroll
  for (int c = 0; c < NQ; c++) {
    double s = 0;
#pragma unroll
    for (int k = 0; k < NP; k++) s = fma(B[c * NP + k], r[k], s);
    u[c] = s;
  }
}
template <int NQ> DI void put_z(double *A, int a, int b, const double (&u)[NQ])
{
#pragma unroll
  for (int c = 0; c < NQ; c++) A[Lay<NQ>::at(a, b, c)] = u[c];
}
template <int NQ> DI void get_z(const double *A, int a, int b, double (&u)[NQ])
{
#pragma unroll
  for (int c = 0; c < NQ; c++) u[c] = A[Lay<NQ>::at(a, b, c)];
}

// ---- collocated derivatives at the points ------------------------------------------------------------------------------
template <int NQ> DI void der_z(const double (&Dt)[NQ * NQ], const double (&u)[NQ], double (&g)[NQ])
{
#pragma unroll
  for (int c = 0; c < NQ; c++) {
    double s = 0;
#pragma unroll
    for (int m = 0; m < NQ; m++) s = fma(Dt[c * NQ + m], u[m], s);
    g[c] = s;
  }
}
// thread (b,c): G[a][b][c] = sum_m Dt[a][m] A[m][b][c]
template <int NQ> DI void der_x(const double (&Dt)[NQ * NQ], const double *A, double *G, int b, int c)
{
  double r[NQ];
#pragma unroll
  for (int m = 0; m < NQ; m++) r[m] = A[Lay<NQ>::at(m, b, c)];
#pragma unroll
  for (int a = 0; a < NQ; a++) {
    double s = 0;
#pragma unroll
    for (int m = 0; m < NQ; m++) s = fma(Dt[a * NQ + m], r[m], s);
    G[Lay<NQ>::at(a, b, c)] = s;
  }
}
// thread (a,c): G[a][b][c] = sum_m Dt[b][m] A[a][m][c]
template <int NQ> DI void der_y(const double (&Dt)[NQ * NQ], const double *A, double *G, int a, int c)
{
  double r[NQ];
#pragma unroll
  for (int m = 0; m < NQ; m++) r[m] = A[Lay<NQ>::at(a, m, c)];
#pragma unroll
  for (int b = 0; b < NQ; b++) {
    double s = 0;
#pragma unroll
    for (int m = 0; m < NQ; m++) s = fma(Dt[b * NQ + m], r[m], s);
    G[Lay<NQ>::at(a, b, c)] = s;
  }
}

// ---- transposes -----------------------------------------------------------------------------------------------------------
// h[m] += sum_c Dt[c][m] g[c]
template <int NQ> DI void der_zT_add(const double (&Dt)[NQ * NQ], const double (&g)[NQ], double (&h)[NQ])
{
#pragma unroll
  for (int m = 0; m < NQ; m++) {
    double s = h[m];
#pragma unroll
    for (int c = 0; c < NQ; c++) s = fma(Dt[c * NQ + m], g[c], s);
    h[m] = s;
  }
}
// thread (b,c), in place on its own x line: G[m][b][c] <- sum_a Dt[a][m] G[a][b][c]
template <int NQ> DI void der_xT(const double (&Dt)[NQ * NQ], double *G, int b, int c)
{
  double r[NQ];
#pragma unroll
  for (int a = 0; a < NQ; a++) r[a] = G[Lay<NQ>::at(a, b, c)];
#pragma unroll
  for (int m = 0; m < NQ; m++) {
    double s = 0;
#pragma unroll
    for (int a = 0; a < NQ; a++) s = fma(Dt[a * NQ + m], r[a], s);
    G[Lay<NQ>::at(m, b, c)] = s;
  }
}
template <int NQ> DI void der_yT(const double (&Dt)[NQ * NQ], double *G, int a, int c)
{
  double r[NQ];
#pragma unroll
  for (int b = 0; b < NQ; b++) r[b] = G[Lay<NQ>::at(a, b, c)];
#pragma unroll
  for (int m = 0; m < NQ; m++) {
    double s = 0;
#pragma unroll
    for (int b = 0; b < NQ; b++) s = fma(Dt[b * NQ + m], r[b], s);
    G[Lay<NQ>::at(a, m, c)] = s;
  }
}
// z^T: thread (a,b), g[c] in registers -> A[a][b][k]
template <int NP, int NQ> DI void bwd_z(const double (&B)[NQ * NP], const double (&g)[NQ], double *A, int a, int b)
{
#pragma unroll
  for (int k = 0; k < NP; k++) {
    double s = 0;
#pragma unroll
    for (int c = 0; c < NQ; c++) s = fma(B[c * NP + k], g[c], s);
    A[Lay<NQ>::at(a, b, k)] = s;
  }
}
// y^T: thread (a,k), in place
template <int NP, int NQ> DI void bwd_y(const double (&B)[NQ * NP], double *A, int a, int k)
{
  double r[NQ];
#pragma unroll
  for (int b = 0; b < NQ; b++) r[b] = A[Lay<NQ>::at(a, b, k)];
#pragma unroll
  for (int j = 0; j < NP; j++) {
    double s = 0;
#pragma unroll
    for (int b = 0; b < NQ; b++) s = fma(B[b * NP + j], r[b], s);
    A[Lay<NQ>::at(a, j, k)] = s;
  }
}
// x^T: thread (j,k), result in registers (scattered to HBM by the caller)
template <int NP, int NQ> DI void bwd_x(const double (&B)[NQ * NP], const double *A, int j, int k, double (&o)[NP])
{
  double r[NQ];
#pragma unroll
  for (int a = 0; a < NQ; a++) r[a] = A[Lay<NQ>::at(a, j, k)];
#pragma unroll
  for (int i = 0; i < NP; i++) {
    double s = 0;
#pragma unroll
    for (int a = 0; a < NQ; a++) s = fma(B[a * NP + i], r[a], s);
    o[i] = s;
  }
}

// ---- quadrature geometry of a trilinear (Q1) hex ----------------------------------------------------------------------------
// What dRulesetIteratorGetPatchApplied does for link 0 (src/fs/interface/fsrulesetit.c:545-560): GRAD of the Q1
// coordinate space with NULL jinv, then dRuleComputePhysical (src/jacobi/interface/jacobi.c:383-414) = cofactor inverse
// (dGeomInvert3, include/dohpgeom.h:57-75) and jw = w0 w1 w2 det.  X = the element's 8 vertices [i][j][k][xyz] in tensor
// order.  For a trilinear map d x/d xi_0 does not depend on xi_0 (etc.), so a thread that owns the line (a,b,:) keeps
// d x/d xi_2 and the k-halves of the other two columns in registers.
struct LineGeom {
  double e0[2][3], e1[2][3], j2[3];      // dx/dxi0 and dx/dxi1 at k=0,1 ; dx/dxi2
  double p0[2][3];                       // position at k=0,1
};
DI void line_geom(const double *X, double qa, double qb, LineGeom &g, bool want_pos)
{
  const double a0 = 0.5 * (1 - qa), a1 = 0.5 * (1 + qa), b0 = 0.5 * (1 - qb), b1 = 0.5 * (1 + qb);
#pragma unroll
  for (int d = 0; d < 3; d++) {
#pragma unroll
    for (int k = 0; k < 2; k++) {
      const double x00 = X[(0 * 4 + 0 * 2 + k) * 3 + d], x01 = X[(0 * 4 + 1 * 2 + k) * 3 + d];
      const double x10 = X[(1 * 4 + 0 * 2 + k) * 3 + d], x11 = X[(1 * 4 + 1 * 2 + k) * 3 + d];
      g.e0[k][d] = 0.5 * (b0 * (x10 - x00) + b1 * (x11 - x01));
      g.e1[k][d] = 0.5 * (a0 * (x01 - x00) + a1 * (x11 - x10));
      g.p0[k][d] = a0 * (b0 * x00 + b1 * x01) + a1 * (b0 * x10 + b1 * x11);
    }
    g.j2[d] = 0.5 * (g.p0[1][d] - g.p0[0][d]);
  }
  (void)want_pos;
}
// 1/x to full double precision for a normal positive x: MUFU.RCP64H seed (about 20 bits) + two Newton steps.  The
// compiler's IEEE division carries a slow-path call for denormals/infinities that det(J) > 0 never needs.
DI double fast_rcp(double x)
{
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}
// Geometry at one point in adjugate form: adj = det(J) * J^{-1}, i.e. jinv[k][e] = d xi_k / d x_e = adj[k*3+e] * rdet
// (dGeomInvert3, include/dohpgeom.h:57-75, without the nine multiplications by 1/det, which the callers fold into
// one scale factor).
struct PointGeom { double adj[9], det, rdet; };
DI void point_geom_adj(const LineGeom &g, double qc, PointGeom &pg)
{
  const double c0 = 0.5 * (1 - qc), c1 = 0.5 * (1 + qc);
  double J[9];                             // J[d][e] = d x_d / d xi_e
#pragma unroll
  for (int d = 0; d < 3; d++) {
    J[d * 3 + 0] = c0 * g.e0[0][d] + c1 * g.e0[1][d];
    J[d * 3 + 1] = c0 * g.e1[0][d] + c1 * g.e1[1][d];
    J[d * 3 + 2] = g.j2[d];
  }
  pg.adj[0] = J[4] * J[8] - J[5] * J[7];
  pg.adj[1] = J[2] * J[7] - J[1] * J[8];
  pg.adj[2] = J[1] * J[5] - J[2] * J[4];
  pg.adj[3] = J[5] * J[6] - J[3] * J[8];
  pg.adj[4] = J[0] * J[8] - J[2] * J[6];
  pg.adj[5] = J[2] * J[3] - J[0] * J[5];
  pg.adj[6] = J[3] * J[7] - J[4] * J[6];
  pg.adj[7] = J[1] * J[6] - J[0] * J[7];
  pg.adj[8] = J[0] * J[4] - J[1] * J[3];
  pg.det = J[0] * pg.adj[0] + J[1] * pg.adj[3] + J[2] * pg.adj[6];
  pg.rdet = fast_rcp(pg.det);
}
// A parallelepiped (affine map): the Jacobian is the same at every point of the element, so a thread evaluates the
// geometry once instead of once per quadrature point.  Exact comparison: any rounding difference just selects the
// general path.
DI bool line_is_affine(const LineGeom &g)
{
  bool a = true;
#pragma unroll
  for (int d = 0; d < 3; d++) a = a && g.e0[0][d] == g.e0[1][d] && g.e1[0][d] == g.e1[1][d];
  return a;
}
// jinv[k][e] = d xi_k / d x_e, returns det(J)
DI double point_geom(const LineGeom &g, double qc, double (&jinv)[9])
{
  PointGeom pg;
  point_geom_adj(g, qc, pg);
#pragma unroll
  for (int i = 0; i < 9; i++) jinv[i] = pg.adj[i] * pg.rdet;
  return pg.det;
}
DI void point_pos(const LineGeom &g, double qc, double (&x)[3])
{
  const double c0 = 0.5 * (1 - qc), c1 = 0.5 * (1 + qc);
#pragma unroll
  for (int d = 0; d < 3; d++) x[d] = c0 * g.p0[0][d] + c1 * g.p0[1][d];
}

}  // namespace dohp
