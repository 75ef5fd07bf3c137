// Element-dense assembly: the dEFSGetExplicitSparse_Basic path of the reference (src/jacobi/interface/jacobi.c:611-643,
// SURVEY.md 8a row Q1-b), used where the patches of the rule do not line up with a Q1 sub-element structure:
//   * EllipJacobian with -ellip_jac_qmethod fast (src/fs/tests/ellip.c:474-536 on the FAST rule: one "patch" = the element,
//     all P^3 basis functions; pinned upstream by ellip-morph-qfast.refout: LU of this matrix solves the Jacobian exactly);
//   * StokesJacobianAssemble_Pressure (src/fs/tests/stokes.c:739-794): the pressure basis (degree p-1) on the velocity
//     SPARSE rule: D = scaled mass  sum_q phi_i jw/eta phi_j,  Daux = Neumann Laplacian  sum_q grad phi_i . grad phi_j jw.
// Every pair of nodes sharing an element couples.  The reference walks the patches and applies the element basis to every
// unit vector (O(P^3) dEFSApply per element); here the element matrix
//     K[i][j] = sum_q sum_{a,b} Phi^a_i(q) C^{ab}_q Phi^b_j(q),   Phi^0 = phi, Phi^{1..3} = reference derivatives,
// is formed term by term with sum factorisation over PAIRS of 1-D indices (three contractions per term), with the per-point
// coefficients C^{00} (mass) and C^{kl} = jw jq A jq^T (stiffness; jq = the inverse Jacobian the reference maps the basis
// with).  Quirk (iii) of the reference (jacobi.c:619-625: cjinv is permuted with qidx the wrong way round, so the gradient at
// element point m is mapped with the inverse Jacobian of point patchind[m]) is replicated for the SPARSE rule; it is the
// identity on the FAST rule.  This path is a completeness path (setup-time preconditioners of small problems), written for
// correctness: one block per element, K in shared memory (P^6 doubles), binary-searched atomic scatter.
#include "internal.h"
#include "tensor_stages.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <algorithm>
#include <cstdlib>
#include <cstring>

using namespace dohp;

#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return seterr(700 + (int)e_, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); } while (0)
#define CHK(call) do { int e_ = (call); if (e_) return e_; } while (0)
static inline int nblk(long n, int t) { return (int)((n + t - 1) / t); }
static bool is_dev(const void *p)
{
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}
template <class T> static int to_dev(dB200Ctx c, const T *p, size_t count, const T **dev, T **owned)
{
  *owned = nullptr;
  if (is_dev(p)) { *dev = p; return 0; }
  CU(cudaMalloc((void **)owned, std::max<size_t>(count, 1) * sizeof(T)));
  CU(cudaMemcpyAsync(*owned, p, count * sizeof(T), cudaMemcpyHostToDevice, c->stream));
  *dev = *owned;
  return 0;
}

__device__ __forceinline__ int dn_row(int loc, int n, int nc) { return loc < n ? loc : (loc >= nc ? n + (loc - nc) : -1); }
__device__ __forceinline__ long dn_find(const int *rowptr, const int *colidx, int row, int col)
{
  int lo = rowptr[row], hi = rowptr[row + 1] - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1, v = colidx[mid];
    if (v == col) return mid;
    if (v < col) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}

// ---- pattern: sorted unique (row, column) pairs of all node pairs sharing an element -------------------------------------------
__global__ void dense_keys_kernel(long npairs, int P3, int n, int nc, const int *e2l, const int *l2g, unsigned long long *keys)
{
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= npairs) return;
  const long e = g / ((long)P3 * P3);
  const int r = (int)(g - e * P3 * P3), i = r / P3, j = r - i * P3;
  const int row = dn_row(e2l[e * P3 + i], n, nc), col = l2g[e2l[e * P3 + j]];
  keys[g] = (row >= 0 && col >= 0) ? ((unsigned long long)row << 32 | (unsigned)col) : ~0ULL;
}
__global__ void dense_split_kernel(long nu, const unsigned long long *keys, int *cnt, int *colidx)
{
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nu || keys[g] == ~0ULL) return;
  atomicAdd(cnt + (int)(keys[g] >> 32), 1);
  colidx[g] = (int)(keys[g] & 0xffffffffu);
}

struct DenseCache {
  int nrows = 0;
  long long nnz = 0;
  int *d_rowptr = nullptr, *d_colidx = nullptr;
};
void dohp_dense_free(dB200Op o)
{
  for (int f = 0; f < DB200_MAXFIELDS; f++) {
    DenseCache *pc = (DenseCache *)o->densecache[f];
    if (!pc) continue;
    cudaFree(pc->d_rowptr); cudaFree(pc->d_colidx);
    delete pc;
    o->densecache[f] = nullptr;
  }
}

extern "C" int dB200OpDensePattern(dB200Op o, int ifield, const int *l2g, long long *nrows, long long *nnz, int *rowptr, int *colidx)
{
  dB200Ctx c = o->ctx;
  CU(cudaSetDevice(c->device));
  if (ifield < 0 || ifield >= o->nspaces) return seterr(63, "dense pattern: field %d out of range", ifield);
  dB200Space s = o->space[ifield];
  DenseCache *pc = (DenseCache *)o->densecache[ifield];
  if (!pc) {
    const int P3 = s->P * s->P * s->P, nr = s->n + s->ngh;
    const long npairs = (long)o->nelem * P3 * P3;
    const int *dl2g;
    int *own_l2g;
    CHK(to_dev(c, l2g, (size_t)s->nc + s->ngh, &dl2g, &own_l2g));
    unsigned long long *k0 = nullptr, *k1 = nullptr, *ku = nullptr;
    long long *dnum = nullptr, hnum = 0;
    void *tmp = nullptr;
    size_t tb = 0, tb2 = 0, tb3 = 0;
    const size_t np1 = std::max<long>(npairs, 1);
    CU(cudaMalloc(&k0, np1 * 8)); CU(cudaMalloc(&k1, np1 * 8)); CU(cudaMalloc(&ku, np1 * 8)); CU(cudaMalloc(&dnum, 8));
    if (npairs) dense_keys_kernel<<<nblk(npairs, 256), 256, 0, c->stream>>>(npairs, P3, s->n, s->nc, s->d_e2l, dl2g, k0);
    CU(cub::DeviceRadixSort::SortKeys(nullptr, tb, k0, k1, (int)npairs, 0, 64, c->stream));
    CU(cub::DeviceSelect::Unique(nullptr, tb2, k1, ku, dnum, (int)npairs, c->stream));
    int *cnt = nullptr;
    CU(cudaMalloc(&cnt, ((size_t)nr + 1) * sizeof(int)));
    pc = new DenseCache;
    CU(cudaMalloc(&pc->d_rowptr, ((size_t)nr + 1) * sizeof(int)));
    CU(cub::DeviceScan::ExclusiveSum(nullptr, tb3, cnt, pc->d_rowptr, nr + 1, c->stream));
    CU(cudaMalloc(&tmp, std::max(tb, std::max(tb2, tb3))));
    if (npairs >= (1L << 31)) return seterr(56, "dense pattern: %ld node pairs exceed this build's 32-bit sort", npairs);
    CU(cub::DeviceRadixSort::SortKeys(tmp, tb, k0, k1, (int)npairs, 0, 64, c->stream));
    CU(cub::DeviceSelect::Unique(tmp, tb2, k1, ku, dnum, (int)npairs, c->stream));
    CU(cudaMemcpyAsync(&hnum, dnum, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaMemsetAsync(cnt, 0, ((size_t)nr + 1) * sizeof(int), c->stream));
    CU(cudaMalloc(&pc->d_colidx, std::max<size_t>(hnum, 1) * sizeof(int)));
    if (hnum) dense_split_kernel<<<nblk(hnum, 256), 256, 0, c->stream>>>(hnum, ku, cnt, pc->d_colidx);   // sentinel (if any) sorts last
    CU(cub::DeviceScan::ExclusiveSum(tmp, tb3, cnt, pc->d_rowptr, nr + 1, c->stream));
    int last = 0;
    CU(cudaMemcpyAsync(&last, pc->d_rowptr + nr, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    c->nlaunch += 6;
    pc->nrows = nr; pc->nnz = last;
    cudaFree(k0); cudaFree(k1); cudaFree(ku); cudaFree(dnum); cudaFree(tmp); cudaFree(cnt);
    if (own_l2g) cudaFree(own_l2g);
    o->densecache[ifield] = pc;
  }
  if (nrows) *nrows = pc->nrows;
  if (nnz) *nnz = pc->nnz;
  if (rowptr) CU(cudaMemcpyAsync(rowptr, pc->d_rowptr, ((size_t)pc->nrows + 1) * sizeof(int), is_dev(rowptr) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
  if (colidx) CU(cudaMemcpyAsync(colidx, pc->d_colidx, (size_t)pc->nnz * sizeof(int), is_dev(colidx) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

// ---- element kernel ----------------------------------------------------------------------------------------------------------------
struct DenseArgs {
  int which, nelem;
  int Q, Pu, bsu, Pm;             // rule points per direction; state field nodes / components; matrix field nodes
  int sparse;                     // 1: SPARSE rule (quirk (iii) aliasing of the inverse Jacobian), 0: FAST rule
  int fi, fj, bs2;                // elast: component pair of this pass; entries per CSR block (1 or 9)
  const double *tab;              // device: Bu[Q][Pu], Du[Q][Pu], Bm[Q][Pm], Dm[Q][Pm], qs[Q], ws[Q]
  const int *e2l_u, *e2l_m, *l2g;
  int nu_own, nm_own, nm_nc;
  const double *coords, *x, *xaux;
  const int *rowptr, *colidx;
  double *vals;
  int *flag;                      // device: bit 0 set when a coupling is missing from the pattern
  double c0, c1, c2, c3;          // ellip: (p-2)/2, eps^2, lambda ; stokes: B, eps^2/2, (p-2)/2 ; elast: lambda, mu
};

// Per-point coefficients of element point m (element-natural numbering): out[0] mass, out[1 + 3k + l] stiffness C^{kl} (k: test side, l: trial
// side) already folded with the inverse Jacobian the reference maps the basis with (quirk (iii) on the SPARSE rule).  Elasticity: the pass
// of component pair (fi, fj), C = T^{fi fj} of csrc/q1_elast_kernel.cuh (ElastPointwiseJacobian on unit trial functions, elast.c:270-280).
__device__ void dense_point_coeffs(const DenseArgs &A, int m, const double *Bu, const double *Du, const double *qs, const double *ws, const double *X, const double *U, double (&out)[10])
{
  const int Q = A.Q, Pu = A.Pu, Pu3 = Pu * Pu * Pu, bsu = A.bsu;
  {
    const int qc = m % Q, qb = (m / Q) % Q, qa = m / (Q * Q);
    LineGeom lg;
    PointGeom pg;
    line_geom(X, qs[qa], qs[qb], lg, false);
    point_geom_adj(lg, qs[qc], pg);
    const double jw = ws[qa] * ws[qb] * ws[qc] * pg.det;
    double jinv[9], jq[9];
    for (int i = 0; i < 9; i++) jinv[i] = pg.adj[i] * pg.rdet;
    if (A.sparse) {                                        // quirk (iii): the basis gradient at point m is mapped with jinv of point patchind[m]
      const int p = Q / 2, pt = m / 8, q = m % 8, pk = pt % p, pj = (pt / p) % p, pi = pt / (p * p);
      const int ma = 2 * pi + (q >> 2), mb = 2 * pj + ((q >> 1) & 1), mc = 2 * pk + (q & 1);
      PointGeom pq;
      line_geom(X, qs[ma], qs[mb], lg, false);
      point_geom_adj(lg, qs[mc], pq);
      for (int i = 0; i < 9; i++) jq[i] = pq.adj[i] * pq.rdet;
    } else {
      for (int i = 0; i < 9; i++) jq[i] = jinv[i];
    }
    // state: value and physical gradient of the bsu components (direct evaluation)
    double uval[3] = {0, 0, 0}, Dux[3][3];
    for (int d = 0; d < bsu; d++) {
      double g0 = 0, g1 = 0, g2 = 0, v = 0;
      for (int i = 0; i < Pu; i++)
        for (int j = 0; j < Pu; j++) {
          const double bij = Bu[qa * Pu + i] * Bu[qb * Pu + j], dij = Du[qa * Pu + i] * Bu[qb * Pu + j], bdj = Bu[qa * Pu + i] * Du[qb * Pu + j];
          for (int k = 0; k < Pu; k++) {
            const double u = U[d * Pu3 + (i * Pu + j) * Pu + k], bk = Bu[qc * Pu + k];
            v = fma(bij * bk, u, v); g0 = fma(dij * bk, u, g0); g1 = fma(bdj * bk, u, g1); g2 = fma(bij * Du[qc * Pu + k], u, g2);
          }
        }
      uval[d] = v;
      for (int x = 0; x < 3; x++) Dux[d][x] = g0 * jinv[x] + g1 * jinv[3 + x] + g2 * jinv[6 + x];      // efstopo.c:574
    }
    double c0 = 0, Axy[6] = {1, 1, 1, 0, 0, 0};            // A = [00,11,22,01,02,12]
    if (A.which == dB200_DENSE_ELLIP_FAST) {               // EllipPointwiseComputeStore + Jacobian on unit trial functions (ellip.c:333-377,495-525)
      const double gamma = 0.5 * (Dux[0][0] * Dux[0][0] + Dux[0][1] * Dux[0][1] + Dux[0][2] * Dux[0][2]), espg = A.c1 + gamma;
      const double eta = A.c0 != 0 ? pow(espg, A.c0) : 1.0, sq = A.c0 != 0 ? sqrt(-A.c0 * eta / espg) : 0.0;
      const double s[3] = {sq * Dux[0][0], sq * Dux[0][1], sq * Dux[0][2]};
      Axy[0] = eta - s[0] * s[0]; Axy[1] = eta - s[1] * s[1]; Axy[2] = eta - s[2] * s[2];
      Axy[3] = -s[0] * s[1]; Axy[4] = -s[0] * s[2]; Axy[5] = -s[1] * s[2];
      c0 = A.c2 != 0 ? -jw * A.c2 * exp(uval[0]) : 0.0;
    } else if (A.which != dB200_DENSE_ELAST_FAST) {        // StokesPointwiseComputeStore (stokes.c:447-453)
      double colon = 0;
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) { const double d = 0.5 * (Dux[a][b] + Dux[b][a]); colon += d * d; }
      const double gamma = A.c1 + 0.5 * colon, eta = A.c0 * (A.c2 != 0 ? pow(gamma, A.c2) : 1.0);
      c0 = A.which == dB200_DENSE_STOKES_D ? jw / eta : 0.0;             // stokes.c:781
    }
    out[0] = c0;
    if (A.which == dB200_DENSE_ELAST_FAST) {
      double F[3][3], S[3][3], g[3][3];                    // F = 1 + Du (elast.c:225), S = lambda tr(E) 1 + 2 mu E (elast.c:227-241), g_f = jq F_f
      for (int d = 0; d < 3; d++)
        for (int x = 0; x < 3; x++) F[d][x] = (d == x ? 1.0 : 0.0) + Dux[d][x];
      double tr = 0;
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) {
          const double E = 0.5 * (F[0][a] * F[0][b] + F[1][a] * F[1][b] + F[2][a] * F[2][b] - (a == b ? 1.0 : 0.0));
          S[a][b] = 2 * A.c1 * E;
          if (a == b) tr += E;
        }
      for (int a = 0; a < 3; a++) S[a][a] += A.c0 * tr;
      for (int f = 0; f < 3; f++)
        for (int k = 0; k < 3; k++) g[f][k] = jq[k * 3] * F[f][0] + jq[k * 3 + 1] * F[f][1] + jq[k * 3 + 2] * F[f][2];
      const int fi = A.fi, fj = A.fj;
      const double FF = F[fi][0] * F[fj][0] + F[fi][1] * F[fj][1] + F[fi][2] * F[fj][2];
      for (int k = 0; k < 3; k++)
        for (int l = 0; l < 3; l++) {
          double M = 0, SS = 0;
          for (int x = 0; x < 3; x++) {
            M += jq[k * 3 + x] * jq[l * 3 + x];
            for (int y = 0; y < 3; y++) SS += jq[k * 3 + x] * S[x][y] * jq[l * 3 + y];
          }
          out[1 + 3 * k + l] = jw * (A.c0 * g[fi][k] * g[fj][l] + A.c1 * g[fj][k] * g[fi][l] + A.c1 * FF * M + (fi == fj ? SS : 0.0));
        }
      return;
    }
    const double Af[3][3] = {{Axy[0], Axy[3], Axy[4]}, {Axy[3], Axy[1], Axy[5]}, {Axy[4], Axy[5], Axy[2]}};
    for (int k = 0; k < 3; k++)
      for (int l = 0; l < 3; l++) {
        double v = 0;
        for (int x = 0; x < 3; x++)
          for (int y = 0; y < 3; y++) v += jq[k * 3 + x] * Af[x][y] * jq[l * 3 + y];
        out[1 + 3 * k + l] = A.which == dB200_DENSE_STOKES_D ? 0.0 : jw * v;
      }
  }
}

__global__ void __launch_bounds__(256) dense_elem_kernel(const DenseArgs A)
{
  extern __shared__ double sm[];
  const int Q = A.Q, Pu = A.Pu, Pm = A.Pm, Q3 = Q * Q * Q, Pu3 = Pu * Pu * Pu, Pm3 = Pm * Pm * Pm, bsu = A.bsu;
  const int tid = threadIdx.x, nth = blockDim.x;
  const long e = blockIdx.x;
  double *Bu = sm, *Du = Bu + Q * Pu, *Bm = Du + Q * Pu, *Dm = Bm + Q * Pm, *qs = Dm + Q * Pm, *ws = qs + Q, *X = ws + Q;
  double *U = X + 24, *Cq = U + bsu * Pu3, *W1 = Cq + 10 * Q3, *W2 = W1 + Q * Q * Pm * Pm, *K = W2 + Q * Pm * Pm * Pm * Pm;
  for (int i = tid; i < 2 * Q * Pu + 2 * Q * Pm + 2 * Q; i += nth) sm[i] = A.tab[i];
  for (int i = tid; i < 24; i += nth) X[i] = A.coords[e * 24 + i];
  for (int i = tid; i < Pu3; i += nth) {                  // dFSGlobalToExpanded(INHOMOGENEOUS)
    const int idx = A.e2l_u[e * Pu3 + i];
    for (int d = 0; d < bsu; d++) U[d * Pu3 + i] = idx < A.nu_own ? A.x[(long)idx * bsu + d] : A.xaux[(long)(idx - A.nu_own) * bsu + d];
  }
  for (int i = tid; i < Pm3 * Pm3; i += nth) K[i] = 0;
  __syncthreads();
  // ---- per-point coefficients: Cq[0] mass, Cq[1 + 3k + l] stiffness ------------------------------------------------------------------
  for (int m = tid; m < Q3; m += nth) {
    double cq[10];
    dense_point_coeffs(A, m, Bu, Du, qs, ws, X, U, cq);
    for (int t = 0; t < 10; t++) Cq[t * Q3 + m] = cq[t];
  }
  __syncthreads();
  // ---- terms: mass (a = b = 0) and the nine (k,l) derivative pairs ----------------------------------------------------------------
  const int PP = Pm * Pm;
  for (int term = 0; term < 10; term++) {
    int ka = -1, kb = -1, cidx = 0;
    if (term > 0) {
      ka = (term - 1) / 3; kb = (term - 1) % 3;
      cidx = term;
      if (A.which == dB200_DENSE_STOKES_D) break;
    } else if (A.which == dB200_DENSE_STOKES_DAUX || A.which == dB200_DENSE_ELAST_FAST) continue;
    const double *C = Cq + cidx * Q3;
    const double *X2 = ka == 2 ? Dm : Bm, *Y2 = kb == 2 ? Dm : Bm, *X1 = ka == 1 ? Dm : Bm, *Y1 = kb == 1 ? Dm : Bm, *X0 = ka == 0 ? Dm : Bm, *Y0 = kb == 0 ? Dm : Bm;
    for (int it = tid; it < Q * Q * PP; it += nth) {      // W1[qa][qb][i2][j2] = sum_qc X2[qc][i2] Y2[qc][j2] C[qa][qb][qc]
      const int ab = it / PP, ij = it - ab * PP, i2 = ij / Pm, j2 = ij - i2 * Pm;
      double s = 0;
      for (int qc = 0; qc < Q; qc++) s = fma(X2[qc * Pm + i2] * Y2[qc * Pm + j2], C[ab * Q + qc], s);
      W1[it] = s;
    }
    __syncthreads();
    for (int it = tid; it < Q * PP * PP; it += nth) {     // W2[qa][i1][j1][i2j2] = sum_qb X1[qb][i1] Y1[qb][j1] W1[qa][qb][i2j2]
      const int qa = it / (PP * PP), r = it - qa * PP * PP, ij1 = r / PP, ij2 = r - ij1 * PP, i1 = ij1 / Pm, j1 = ij1 - i1 * Pm;
      double s = 0;
      for (int qb = 0; qb < Q; qb++) s = fma(X1[qb * Pm + i1] * Y1[qb * Pm + j1], W1[(qa * Q + qb) * PP + ij2], s);
      W2[it] = s;
    }
    __syncthreads();
    for (int it = tid; it < Pm3 * Pm3; it += nth) {       // K[i][j] += sum_qa X0[qa][i0] Y0[qa][j0] W2[qa][i1 j1][i2 j2]
      const int i = it / Pm3, j = it - i * Pm3, i0 = i / PP, i1 = (i / Pm) % Pm, i2 = i % Pm, j0 = j / PP, j1 = (j / Pm) % Pm, j2 = j % Pm;
      double s = 0;
      for (int qa = 0; qa < Q; qa++) s = fma(X0[qa * Pm + i0] * Y0[qa * Pm + j0], W2[(qa * PP + i1 * Pm + j1) * PP + i2 * Pm + j2], s);
      K[it] += s;
    }
    __syncthreads();
  }
  // ---- dFSMatSetValuesBlockedExpanded (fs.c:826-930): Dirichlet rows/columns dropped ------------------------------------------------
  for (int it = tid; it < Pm3 * Pm3; it += nth) {
    const int i = it / Pm3, j = it - i * Pm3;
    const int row = dn_row(A.e2l_m[e * Pm3 + i], A.nm_own, A.nm_nc), col = A.l2g[A.e2l_m[e * Pm3 + j]];
    if (row < 0 || col < 0) continue;
    const long pos = dn_find(A.rowptr, A.colidx, row, col);
    if (pos >= 0) atomicAdd(A.vals + pos * A.bs2 + (A.fi * 3 + A.fj), K[it]);
    else atomicOr(A.flag, 1);                              // coupling missing from the caller's pattern: reported, never dropped silently
  }
}

// ---- column path: any degree -------------------------------------------------------------------------------------------------------------
// When the element matrix does not fit in shared memory (matrix-field degree > 4) the columns are formed one at a time, which is what
// the reference does (dEFSGetExplicit applies the basis to every unit vector, jacobi.c:508-559): kernel 1 stores the per-point
// coefficients of a chunk of elements, kernel 2 (one block per element and column j) evaluates f^a(q) = sum_b C^{ab}_q Phi^b_j(q) on the
// fly and applies the transposed basis by sum factorisation, K[:,j] = sum_a sum_q Phi^a_i(q) f^a(q), then scatters the column.
__global__ void __launch_bounds__(128) dense_coeff_kernel(const DenseArgs A, long e0, double *cq_out)
{
  extern __shared__ double sm[];
  const int Q = A.Q, Pu = A.Pu, Pm = A.Pm, Q3 = Q * Q * Q, Pu3 = Pu * Pu * Pu, bsu = A.bsu;
  const int tid = threadIdx.x, nth = blockDim.x;
  const long e = e0 + blockIdx.x;
  double *Bu = sm, *Du = Bu + Q * Pu, *qs = Du + Q * Pu + 2 * Q * Pm, *ws = qs + Q, *X = ws + Q, *U = X + 24;
  for (int i = tid; i < 2 * Q * Pu + 2 * Q * Pm + 2 * Q; i += nth) sm[i] = A.tab[i];
  for (int i = tid; i < 24; i += nth) X[i] = A.coords[e * 24 + i];
  for (int i = tid; i < Pu3; i += nth) {
    const int idx = A.e2l_u[e * Pu3 + i];
    for (int d = 0; d < bsu; d++) U[d * Pu3 + i] = idx < A.nu_own ? A.x[(long)idx * bsu + d] : A.xaux[(long)(idx - A.nu_own) * bsu + d];
  }
  __syncthreads();
  for (int m = tid; m < Q3; m += nth) {
    double cq[10];
    dense_point_coeffs(A, m, Bu, Du, qs, ws, X, U, cq);
    for (int t = 0; t < 10; t++) cq_out[((long)blockIdx.x * 10 + t) * Q3 + m] = cq[t];
  }
}

__global__ void __launch_bounds__(128) dense_column_kernel(const DenseArgs A, long e0, const double *cq_in)
{
  extern __shared__ double sm[];
  const int Q = A.Q, Pm = A.Pm, Q3 = Q * Q * Q, Pm3 = Pm * Pm * Pm, PP = Pm * Pm;
  const int tid = threadIdx.x, nth = blockDim.x;
  const int el = blockIdx.x / Pm3, j = blockIdx.x - el * Pm3, j0 = j / PP, j1 = (j / Pm) % Pm, j2 = j % Pm;
  const long e = e0 + el;
  double *Bm = sm, *Dm = Bm + Q * Pm, *T1 = Dm + Q * Pm, *T2 = T1 + 4 * Q * Q * Pm;
  for (int i = tid; i < 2 * Q * Pm; i += nth) sm[i] = A.tab[2 * Q * A.Pu + i];
  __syncthreads();
  const double *C = cq_in + (long)el * 10 * Q3;
  const bool mass = A.which != dB200_DENSE_STOKES_DAUX && A.which != dB200_DENSE_ELAST_FAST, stiff = A.which != dB200_DENSE_STOKES_D;
  // stage 1: T1[a][qa][qb][i2] = sum_qc X2^a[qc][i2] f^a(qa,qb,qc),  X2^a = Dm for a == 3 else Bm
  for (int it = tid; it < Q * Q * Pm; it += nth) {
    const int i2 = it % Pm, ab = it / Pm, qa = ab / Q, qb = ab - qa * Q;
    const double b0 = Bm[qa * Pm + j0], d0 = Dm[qa * Pm + j0], b1 = Bm[qb * Pm + j1], d1 = Dm[qb * Pm + j1];
    double t0 = 0, t1 = 0, t2 = 0, t3 = 0;
    for (int qc = 0; qc < Q; qc++) {
      const int m = ab * Q + qc;
      const double b2 = Bm[qc * Pm + j2], d2 = Dm[qc * Pm + j2];
      const double phi = b0 * b1 * b2, g0 = d0 * b1 * b2, g1 = b0 * d1 * b2, g2 = b0 * b1 * d2;      // Phi^b_j at the point
      const double xb = Bm[qc * Pm + i2], xd = Dm[qc * Pm + i2];
      if (mass) t0 = fma(xb, C[m] * phi, t0);
      if (stiff) {
        t1 = fma(xb, C[Q3 + m] * g0 + C[2 * Q3 + m] * g1 + C[3 * Q3 + m] * g2, t1);             // f^k = sum_l C^{kl} Phi^l_j
        t2 = fma(xb, C[4 * Q3 + m] * g0 + C[5 * Q3 + m] * g1 + C[6 * Q3 + m] * g2, t2);
        t3 = fma(xd, C[7 * Q3 + m] * g0 + C[8 * Q3 + m] * g1 + C[9 * Q3 + m] * g2, t3);
      }
    }
    T1[it] = t0; T1[Q * Q * Pm + it] = t1; T1[2 * Q * Q * Pm + it] = t2; T1[3 * Q * Q * Pm + it] = t3;
  }
  __syncthreads();
  // stage 2: T2[a][qa][i1][i2] = sum_qb X1^a[qb][i1] T1[a][qa][qb][i2],  X1^a = Dm for a == 2 else Bm
  for (int it = tid; it < 4 * Q * PP; it += nth) {
    const int a = it / (Q * PP), r = it - a * (Q * PP), qa = r / PP, i12 = r - qa * PP, i1 = i12 / Pm, i2 = i12 - i1 * Pm;
    const double *X1 = a == 2 ? Dm : Bm, *src = T1 + a * (Q * Q * Pm) + qa * (Q * Pm) + i2;
    double sacc = 0;
    for (int qb = 0; qb < Q; qb++) sacc = fma(X1[qb * Pm + i1], src[qb * Pm], sacc);
    T2[it] = sacc;
  }
  __syncthreads();
  // stage 3 + dFSMatSetValuesBlockedExpanded (fs.c:826-930): K[i][j] = sum_a sum_qa X0^a[qa][i0] T2[a][qa][i1][i2],  X0^a = Dm for a == 1
  const int col = A.l2g[A.e2l_m[e * Pm3 + j]];
  for (int i = tid; i < Pm3; i += nth) {
    const int i0 = i / PP, i12 = i - i0 * PP;
    double k = 0;
    for (int a = 0; a < 4; a++) {
      const double *X0 = a == 1 ? Dm : Bm, *src = T2 + a * (Q * PP) + i12;
      for (int qa = 0; qa < Q; qa++) k = fma(X0[qa * Pm + i0], src[qa * PP], k);
    }
    const int row = dn_row(A.e2l_m[e * Pm3 + i], A.nm_own, A.nm_nc);
    if (row < 0 || col < 0) continue;
    const long pos = dn_find(A.rowptr, A.colidx, row, col);
    if (pos >= 0) atomicAdd(A.vals + pos * A.bs2 + (A.fi * 3 + A.fj), k);
    else atomicOr(A.flag, 1);
  }
}

extern "C" int dB200OpAssembleDense(dB200Op o, int which, const double *const gx[], const int *l2g, const int *rowptr, const int *colidx, double *vals)
{
  dB200Ctx c = o->ctx;
  CU(cudaSetDevice(c->device));
  const bool stokes = which == dB200_DENSE_STOKES_D || which == dB200_DENSE_STOKES_DAUX;
  const bool elast = which == dB200_DENSE_ELAST_FAST;
  if (which == dB200_DENSE_ELLIP_FAST ? o->kind != dB200_OP_ELLIP : elast ? o->kind != dB200_OP_ELAST : (!stokes || o->kind != dB200_OP_STOKES))
    return seterr(62, "dense assembly %d does not apply to operator kind %d", which, o->kind);
  const int bs2 = elast ? 9 : 1;                           // BAIJ(3) blocks for elasticity (elast.c:479), scalar entries otherwise
  const int fm = stokes ? 1 : 0;
  dB200Space su = o->space[0], smx = o->space[fm];
  DenseCache *pc = (DenseCache *)o->densecache[fm];
  if ((which == dB200_DENSE_ELLIP_FAST || elast) && su->P == 2) {
    // degree 1: the FAST rule has Q = 8 points in one patch, so dEFSGetExplicitSparse_Tensor_Hex does NOT fall back to the
    // Basic path (efstopo.c:668) and the reference assembles the Q1 patch with its deriv1 = +-1 quirk; same pattern
    if (!(rowptr && colidx) && !pc) return seterr(73, "dense assembly without a pattern: call dB200OpDensePattern first or pass rowptr/colidx");
    return dB200OpAssembleQ1(o, gx, l2g, rowptr && colidx ? rowptr : pc->d_rowptr, rowptr && colidx ? colidx : pc->d_colidx, vals);
  }
  if (!gx || !gx[0] || !l2g || !vals) return seterr(85, "dense assembly: null argument");
  struct Tmp { std::vector<void *> p; ~Tmp() { for (void *q : p) cudaFree(q); } } tmp;      // device temporaries freed on every exit path
  const int *drow, *dcol, *dl2g;
  int *own_row = nullptr, *own_col = nullptr, *own_l2g = nullptr;
  long long nnz;
  const int nr = smx->n + smx->ngh;
  if (rowptr && colidx) {
    int last;
    CU(cudaMemcpy(&last, rowptr + nr, sizeof(int), is_dev(rowptr) ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost));
    nnz = last;
    CHK(to_dev(c, rowptr, (size_t)nr + 1, &drow, &own_row));
    if (own_row) tmp.p.push_back(own_row);
    CHK(to_dev(c, colidx, (size_t)nnz, &dcol, &own_col));
    if (own_col) tmp.p.push_back(own_col);
  } else {
    if (!pc) return seterr(73, "dense assembly without a pattern: call dB200OpDensePattern first or pass rowptr/colidx");
    drow = pc->d_rowptr; dcol = pc->d_colidx; nnz = pc->nnz;
  }
  CHK(to_dev(c, l2g, (size_t)smx->nc + smx->ngh, &dl2g, &own_l2g));
  if (own_l2g) tmp.p.push_back(own_l2g);
  // rule and tables: FAST = the operator's rule; SPARSE = 2 Gauss points per GLL sub-interval of the velocity space (stokes.c:292-298)
  const int Pu = su->P, Pm = smx->P, p = Pu - 1;
  int Q;
  std::vector<double> qs(2 * p + 4), ws(2 * p + 4);
  if (stokes) { CHK(rule_1d(dB200_QMETHOD_SPARSE, dB200_GAUSS_LEGENDRE, 2 * p, &Q, qs.data(), ws.data())); }
  else { Q = o->rule->Q[0]; for (int i = 0; i < Q; i++) { qs[i] = o->rule->coord[0][i]; ws[i] = o->rule->weight[0][i]; } }
  std::vector<double> tab((size_t)2 * Q * Pu + 2 * Q * Pm + 2 * Q);
  CHK(basis_1d(Pu, Q, qs.data(), tab.data(), tab.data() + (size_t)Q * Pu, nullptr, nullptr));
  CHK(basis_1d(Pm, Q, qs.data(), tab.data() + (size_t)2 * Q * Pu, tab.data() + (size_t)2 * Q * Pu + (size_t)Q * Pm, nullptr, nullptr));
  for (int i = 0; i < Q; i++) { tab[(size_t)2 * Q * Pu + 2 * Q * Pm + i] = qs[i]; tab[(size_t)2 * Q * Pu + 2 * Q * Pm + Q + i] = ws[i]; }
  double *dtab = nullptr;
  CU(cudaMalloc((void **)&dtab, tab.size() * sizeof(double) + sizeof(int)));
  tmp.p.push_back(dtab);
  int *dflag = (int *)(dtab + tab.size());
  CU(cudaMemsetAsync(dflag, 0, sizeof(int), c->stream));
  CU(cudaMemcpyAsync(dtab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  const size_t nbu = (size_t)su->n * su->bs;
  const double *dx;
  double *own_x = nullptr, *dvals, *own_vals = nullptr;
  CHK(to_dev(c, gx[0], nbu, &dx, &own_x));
  if (own_x) tmp.p.push_back(own_x);
  if (is_dev(vals)) dvals = vals;
  else { CU(cudaMalloc((void **)&own_vals, std::max<size_t>(nnz * bs2, 1) * sizeof(double))); dvals = own_vals; tmp.p.push_back(own_vals); }
  CU(cudaMemsetAsync(dvals, 0, (size_t)nnz * bs2 * sizeof(double), c->stream));                       // MatZeroEntries
  if (!su->d_xaux_inhomog) CHK(dB200OpSetDirichlet(o, 0, nullptr));
  CHK(dohp_halo_forward(su, dx, su->d_xaux_inhomog));
  DenseArgs A;
  memset(&A, 0, sizeof A);
  A.which = which; A.nelem = o->nelem; A.Q = Q; A.Pu = Pu; A.bsu = su->bs; A.Pm = Pm; A.sparse = stokes ? 1 : 0; A.tab = dtab;
  A.e2l_u = su->d_e2l; A.e2l_m = smx->d_e2l; A.l2g = dl2g; A.nu_own = su->n; A.nm_own = smx->n; A.nm_nc = smx->nc;
  A.coords = o->d_coords; A.x = dx; A.xaux = su->d_xaux_inhomog; A.rowptr = drow; A.colidx = dcol; A.vals = dvals; A.flag = dflag;
  A.bs2 = bs2;
  if (elast) { A.c0 = o->elast.lambda; A.c1 = o->elast.mu; }
  else if (stokes) { A.c0 = o->stokes.B; A.c1 = 0.5 * o->stokes.eps * o->stokes.eps; A.c2 = 0.5 * (o->stokes.p - 2); }
  else { A.c0 = (o->ellip.p - 2) / 2; A.c1 = o->ellip.epsilon * o->ellip.epsilon; A.c2 = o->ellip.lambda; }
  const size_t Q3 = (size_t)Q * Q * Q, Pm2 = (size_t)Pm * Pm, Pm3 = Pm2 * Pm;
  const size_t smem = sizeof(double) * (tab.size() + 24 + (size_t)su->bs * Pu * Pu * Pu + 10 * Q3 + (size_t)Q * Q * Pm2 + (size_t)Q * Pm2 * Pm2 + Pm3 * Pm3);
  if (smem <= 227 * 1024 && !getenv("DOHP_B200_DENSE_COLUMNS")) {
    if (o->nelem) {
      CU(cudaFuncSetAttribute(dense_elem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      for (int pr = 0; pr < bs2; pr++) {                   // elasticity: one pass per component pair (fi, fj)
        A.fi = pr / 3; A.fj = pr % 3;
        dense_elem_kernel<<<o->nelem, 256, smem, c->stream>>>(A);
        CU(cudaGetLastError());
        c->nlaunch++;
      }
    }
  } else if (o->nelem) {
    // column path (any degree): per-point coefficients of a chunk of elements in HBM, then one block per (element, column)
    const size_t smem1 = sizeof(double) * (tab.size() + 24 + (size_t)su->bs * Pu * Pu * Pu);
    const size_t smem2 = sizeof(double) * (2 * (size_t)Q * Pm + 4 * (size_t)Q * Q * Pm + 4 * (size_t)Q * Pm2);
    if (smem1 > 227 * 1024 || smem2 > 227 * 1024) return seterr(56, "dense assembly: %zu / %zu B of shared memory needed", smem1, smem2);
    const long chunk = std::max<long>(1, std::min<long>(o->nelem, (256L << 20) / (long)(10 * Q3 * sizeof(double))));
    double *dcq = nullptr;
    CU(cudaMalloc((void **)&dcq, (size_t)chunk * 10 * Q3 * sizeof(double)));
    tmp.p.push_back(dcq);
    CU(cudaFuncSetAttribute(dense_coeff_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    CU(cudaFuncSetAttribute(dense_column_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    for (long e0 = 0; e0 < o->nelem; e0 += chunk) {
      const long ne = std::min<long>(chunk, o->nelem - e0);
      for (int pr = 0; pr < bs2; pr++) {
        A.fi = pr / 3; A.fj = pr % 3;
        dense_coeff_kernel<<<(unsigned)ne, 128, smem1, c->stream>>>(A, e0, dcq);
        dense_column_kernel<<<(unsigned)(ne * Pm3), 128, smem2, c->stream>>>(A, e0, dcq);
        CU(cudaGetLastError());
        c->nlaunch += 2;
      }
    }
  }
  if (own_vals) CU(cudaMemcpyAsync(vals, dvals, (size_t)nnz * bs2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  int hflag = 0;
  CU(cudaMemcpyAsync(&hflag, dflag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (hflag) return seterr(56, "dense assembly: the pattern lacks a coupling of the element matrices (row/column pair not found); values would be incomplete");
  return 0;
}
