// Host-side 1-D tables (setup only).  Restates WHAT the reference tabulates, with its own arithmetic:
//   rules  : TensorGetRule, src/jacobi/quadrature/impls/tensor/tensorquad.c:31-118 (zwgj / zwglj of polylib.c:182-295)
//   bases  : TensorBasisCreate, src/jacobi/impls/tensor/tensor.c:22-72 (zwglj nodes, Imglj interpolation, Dglj collocation
//            derivative, deriv = interp . D)
// Legendre polynomials by the three-term recurrence in long double, Newton for the roots, barycentric Lagrange formulas
// for interpolation/differentiation matrices.  tests/test_abi.py checks these against the oracle to ~1e-14.
#include "internal.h"
#include <cmath>
#include <vector>

namespace dohp {

typedef long double ld;

static void legendre(int n, ld x, ld *P, ld *dP)
{
  ld p0 = 1, p1 = x;
  if (n == 0) { *P = 1; *dP = 0; return; }
  for (int k = 2; k <= n; k++) { ld p2 = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k; p0 = p1; p1 = p2; }
  *P = p1;
  *dP = n * (x * p1 - p0) / (x * x - 1);       // not used at |x|=1
}

void gauss_legendre(int n, double *z, double *w)
{
  const ld pi = 3.14159265358979323846264338327950288L;
  for (int i = 0; i < n; i++) {
    ld x = -cosl(pi * (i + 0.75L) / (n + 0.5L)), P, dP;
    for (int it = 0; it < 100; it++) {
      legendre(n, x, &P, &dP);
      ld dx = P / dP;
      x -= dx;
      if (fabsl(dx) < 1e-19L) break;
    }
    legendre(n, x, &P, &dP);
    z[i] = (double)x;
    w[i] = (double)(2 / ((1 - x * x) * dP * dP));
  }
  for (int i = 0; i < n / 2; i++) {              // enforce exact symmetry
    double a = 0.5 * (z[n - 1 - i] - z[i]); z[i] = -a; z[n - 1 - i] = a;
    double b = 0.5 * (w[i] + w[n - 1 - i]); w[i] = w[n - 1 - i] = b;
  }
  if (n % 2) z[n / 2] = 0;
}

void gauss_lobatto(int n, double *z, double *w)
{
  const ld pi = 3.14159265358979323846264338327950288L;
  const int m = n - 1;                           // nodes: +-1 and the roots of P_m'
  z[0] = -1; z[n - 1] = 1;
  for (int i = 1; i < n - 1; i++) {
    ld x = -cosl(pi * i / m), P, dP;
    for (int it = 0; it < 100; it++) {
      legendre(m, x, &P, &dP);
      // P_m'' from the Legendre ODE: (1-x^2) P'' = 2x P' - m(m+1) P
      ld d2 = (2 * x * dP - m * (m + 1) * P) / (1 - x * x);
      ld dx = dP / d2;
      x -= dx;
      if (fabsl(dx) < 1e-19L) break;
    }
    z[i] = (double)x;
  }
  for (int i = 0; i < n / 2; i++) { double a = 0.5 * (z[n - 1 - i] - z[i]); z[i] = -a; z[n - 1 - i] = a; }
  if (n % 2) z[n / 2] = 0;
  for (int i = 0; i < n; i++) {
    ld P, dP;
    legendre(m, (ld)z[i], &P, &dP);
    w[i] = (double)(2 / ((ld)m * (m + 1) * P * P));
  }
}

int rule_1d(int method, int family, int order, int *size, double *coord, double *weight)
{
  if (!(0 <= order && order < 100)) return seterr(1, "rule order %d out of bounds", order);      // tensorquad.c:60
  int n;
  switch (method) {
  case dB200_QMETHOD_SPARSE: {                    // tensorquad.c:31-50,68-80
    n = order > 0 ? order : 1;
    if (family != dB200_GAUSS_LEGENDRE && family != dB200_GAUSS_LOBATTO) return seterr(56, "Gauss family %d with SPARSE", family);
    if (n == 1) break;
    const double x0 = family == dB200_GAUSS_LEGENDRE ? -0.57735026918962573 : -1, x1 = -x0;
    const int ne = n / 2;
    std::vector<double> iv(ne + 1), un(ne + 1);
    gauss_lobatto(ne + 1, iv.data(), un.data());
    for (int i = 0; i < ne; i++) {
      const double h = iv[i + 1] - iv[i];
      coord[2 * i] = iv[i] + (x0 + 1) * h / 2.;
      coord[2 * i + 1] = iv[i] + (x1 + 1) * h / 2.;
      weight[2 * i] = weight[2 * i + 1] = h / 2.;
    }
    *size = n;
    return 0;
  }
  case dB200_QMETHOD_FAST:                        // tensorquad.c:81-95
    if (family == dB200_GAUSS_LEGENDRE) n = 1 + order / 2;
    else if (family == dB200_GAUSS_LOBATTO) n = 2 + order / 2;
    else return seterr(56, "Gauss family %d with FAST", family);
    break;
  case dB200_QMETHOD_SELF:                        // tensorquad.c:96-104
    if (family != dB200_GAUSS_LOBATTO) return seterr(56, "Gauss family %d with SELF", family);
    n = 1 + order / 2;
    break;
  default: return seterr(56, "quadrature method %d", method);
  }
  *size = n;
  if (n == 1) { coord[0] = 0; weight[0] = 2; return 0; }     // tensorquad.c:109-111
  if (method == dB200_QMETHOD_FAST && family == dB200_GAUSS_LEGENDRE) gauss_legendre(n, coord, weight);
  else gauss_lobatto(n, coord, weight);
  return 0;
}

// Lagrange basis on nodes z[P] evaluated at x[Q]: L[q][j]; and the collocation derivative D[i][j] = l_j'(z_i)
static void lagrange_eval(int P, const double *z, int Q, const double *x, double *L)
{
  std::vector<ld> bw(P);
  for (int j = 0; j < P; j++) { ld p = 1; for (int m = 0; m < P; m++) if (m != j) p *= (ld)z[j] - (ld)z[m]; bw[j] = 1 / p; }
  for (int q = 0; q < Q; q++) {
    int hit = -1;
    for (int j = 0; j < P; j++) if (x[q] == z[j]) hit = j;
    for (int j = 0; j < P; j++) {
      if (hit >= 0) { L[q * P + j] = j == hit ? 1.0 : 0.0; continue; }
      ld p = bw[j];
      for (int m = 0; m < P; m++) if (m != j) p *= (ld)x[q] - (ld)z[m];
      L[q * P + j] = (double)p;
    }
  }
}
void collocation_derivative(int P, const double *z, double *D)
{
  std::vector<ld> bw(P);
  for (int j = 0; j < P; j++) { ld p = 1; for (int m = 0; m < P; m++) if (m != j) p *= (ld)z[j] - (ld)z[m]; bw[j] = 1 / p; }
  for (int i = 0; i < P; i++) {
    ld diag = 0;
    for (int j = 0; j < P; j++) {
      if (i == j) continue;
      ld v = (bw[j] / bw[i]) / ((ld)z[i] - (ld)z[j]);
      D[i * P + j] = (double)v;
      diag -= v;
    }
    D[i * P + i] = (double)diag;
  }
}

int basis_1d(int P, int Q, const double *qcoord, double *interp, double *deriv, double *node, double *weight)
{
  if (P < 2) return seterr(56, "Zeroth order Tensor basis not available (P=%d)", P);         // tensor.c:76
  if (P > 14) return seterr(1, "optimal scaling not available for this order (P=%d > 14)", P); // tensor.c:92-95
  std::vector<double> z(P), w(P), L((size_t)P * Q), D((size_t)P * P);
  gauss_lobatto(P, z.data(), w.data());
  lagrange_eval(P, z.data(), Q, qcoord, L.data());
  collocation_derivative(P, z.data(), D.data());
  if (interp) for (int i = 0; i < P * Q; i++) interp[i] = L[i];
  if (deriv)
    for (int i = 0; i < Q; i++)
      for (int j = 0; j < P; j++) {
        double s = 0;
        for (int k = 0; k < P; k++) s += L[i * P + k] * D[k * P + j];                        // tensor.c:62-68
        deriv[i * P + j] = s;
      }
  if (node) for (int i = 0; i < P; i++) node[i] = z[i];
  if (weight) for (int i = 0; i < P; i++) weight[i] = w[i];
  return 0;
}

}  // namespace dohp

extern "C" int dB200Rule1D(int method, int family, int order, int *size, double *coord, double *weight)
{ return dohp::rule_1d(method, family, order, size, coord, weight); }
extern "C" int dB200Basis1D(int P, int Q, const double *qcoord, double *interp, double *deriv, double *node, double *weight)
{ return dohp::basis_1d(P, Q, qcoord, interp, deriv, node, weight); }
