// Embedded-Q1 preconditioning matrix: sparsity pattern and assembly on the device.
//
// Reference path: XJacobian (src/fs/tests/ellip.c:474-536, elast.c:343-407) walks the p^3 Q1 patches of every element on
// the SPARSE rule (2 Gauss points per GLL sub-interval, src/jacobi/quadrature/impls/tensor/tensorquad.c:31-50,120-169),
// builds the 8-function patch basis with dEFSGetExplicitSparse_Tensor_Hex (src/jacobi/impls/tensor/efstopo.c:660-717),
// forms K[8*bs][8*bs] with the pointwise Jacobian on unit trial functions and adds it through
// dFSMatSetValuesBlockedExpanded (src/fs/interface/fs.c:826-930) into a PETSc matrix whose nonzero structure is
// whatever was inserted.  Here: the pattern is built once on the device as a BLOCK CSR (one block per coupled node
// pair, sorted columns, = PETSc BAIJ(bs) / AIJ for bs=1), and one kernel per assembly evaluates the state on the SPARSE
// points, the patch matrices and scatters them with FP64 atomics into the value array (bs*bs per block, row-major).
//
// Reference quirks replicated on purpose (each is pinned by a golden file):
//   (i)  the patch basis derivative uses deriv1 = +-1 whatever the sub-interval width (efstopo.c:665-666);
//   (ii) the physical map of that derivative takes dXdx = cjinv[(patch*8+q)*9] from the ELEMENT-natural array of inverse
//        Jacobians, i.e. the inverse Jacobian of a different point of the element on non-affine meshes (efstopo.c:692,
//        fsrulesetit.c:677-683); the state and jw are taken at the true point.
#include "internal.h"
#include "tensor_stages.cuh"
#include "q1_kernel.cuh"
#include <cub/device/device_scan.cuh>
#include <algorithm>
#include <cstring>
#include <cstdlib>

using namespace dohp;

#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return seterr(700 + (int)e_, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); } while (0)
#define CHK(call) do { int e_ = (call); if (e_) return e_; } while (0)
static inline int nblk(long n, int t) { return (int)((n + t - 1) / t); }

// ---- pattern ----------------------------------------------------------------------------------------------------------------
// element-local node of corner b = (iii,jjj,kkk) of patch (i,j,k): efstopo.c:676-684
__device__ __forceinline__ int patch_corner(int n, int pi, int pj, int pk, int b) { return ((pi + (b >> 2)) * n + pj + ((b >> 1) & 1)) * n + pk + (b & 1); }

__global__ void q1_count_kernel(long npatch_total, int p, int nown, const int *e2l, int *cnt)
{
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= npatch_total) return;
  const int n = p + 1, p3 = p * p * p;
  const long e = g / p3;
  const int pt = (int)(g - e * p3), pk = pt % p, pj = (pt / p) % p, pi = pt / (p * p);
  for (int b = 0; b < 8; b++) {
    const int row = e2l[e * n * n * n + patch_corner(n, pi, pj, pk, b)];
    if (row < nown) atomicAdd(cnt + row, 8);                 // upper bound: 8 columns per patch containing the row
  }
}
__global__ void q1_insert_kernel(long npatch_total, int p, int nown, const int *e2l, const int *l2g, const long long *off, int *tab)
{
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= npatch_total) return;
  const int n = p + 1, p3 = p * p * p;
  const long e = g / p3;
  const int pt = (int)(g - e * p3), pk = pt % p, pj = (pt / p) % p, pi = pt / (p * p);
  int loc[8], col[8];
  for (int b = 0; b < 8; b++) { loc[b] = e2l[e * n * n * n + patch_corner(n, pi, pj, pk, b)]; col[b] = l2g[loc[b]]; }
  for (int i = 0; i < 8; i++) {
    if (loc[i] >= nown) continue;                            // rows: owned explicit nodes (Dirichlet rows dropped, fsbuild.c:31-34)
    int *seg = tab + off[loc[i]];
    const int cap = (int)(off[loc[i] + 1] - off[loc[i]]);
    for (int j = 0; j < 8; j++) {
      if (col[j] < 0) continue;                              // Dirichlet columns dropped
      for (int s = 0; s < cap; s++) {
        const int old = atomicCAS(seg + s, 0x7fffffff, col[j]);
        if (old == 0x7fffffff || old == col[j]) break;
      }
    }
  }
}
__global__ void q1_sort_kernel(int nown, const long long *off, int *tab, int *len)
{
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nown) return;
  int *seg = tab + off[r];
  const int cap = (int)(off[r + 1] - off[r]);
  int m = 0;
  while (m < cap && seg[m] != 0x7fffffff) m++;
  for (int a = 1; a < m; a++) {                              // insertion sort: PETSc keeps the columns of a row sorted
    const int v = seg[a];
    int b = a - 1;
    while (b >= 0 && seg[b] > v) { seg[b + 1] = seg[b]; b--; }
    seg[b + 1] = v;
  }
  len[r] = m;
}
__global__ void q1_compact_kernel(int nown, const long long *off, const int *tab, const int *rowptr, int *colidx)
{
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nown) return;
  const int *seg = tab + off[r];
  const int m = rowptr[r + 1] - rowptr[r];
  for (int a = 0; a < m; a++) colidx[rowptr[r] + a] = seg[a];
}
__global__ void fill_int_kernel(long n, int v, int *a)
{
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n) a[g] = v;
}
__global__ void widen_kernel(long n, const int *a, long long *b)
{
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n) b[g] = a[g];
}

// position of block (row, col) in the block CSR (columns sorted): binary search
__device__ __forceinline__ long q1_find(const int *rowptr, const int *colidx, int row, int col)
{
  int lo = rowptr[row], hi = rowptr[row + 1] - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1, v = colidx[mid];
    if (v == col) return mid;
    if (v < col) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}
// "COO preallocation" (what MatSetPreallocationCOO does once per pattern): for every (patch, trial corner j, test corner i)
// the offset of block (row_i, col_j) inside row_i, one byte each (0xFF = dropped: Dirichlet row or column), laid out
// [patch][j][i] so that the lane owning trial corner j reads its eight offsets as one 64-bit word.
__global__ void q1_relpos_kernel(long npatch_total, int p, int nown, const int *e2l, const int *l2g, const int *rowptr, const int *colidx, unsigned char *relpos, int *overflow)
{
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= npatch_total * 8) return;
  const long patch = g >> 3;
  const int j = (int)(g & 7);
  const int n = p + 1, p3 = p * p * p;
  const long e = patch / p3;
  const int pt = (int)(patch - e * p3), pk = pt % p, pj = (pt / p) % p, pi = pt / (p * p);
  const int col = l2g[e2l[e * n * n * n + patch_corner(n, pi, pj, pk, j)]];
  unsigned long long word = 0;
  for (int i = 0; i < 8; i++) {
    const int row = e2l[e * n * n * n + patch_corner(n, pi, pj, pk, i)];
    unsigned int rel = 0xFF;
    if (col >= 0 && row < nown) {
      const long pos = q1_find(rowptr, colidx, row, col);
      if (pos >= 0) { rel = (unsigned int)(pos - rowptr[row]); if (rel >= 0xFF) { atomicOr(overflow, 1); rel = 0xFF; } }
    }
    word |= (unsigned long long)rel << (8 * i);
  }
  reinterpret_cast<unsigned long long *>(relpos)[g] = word;
}


// Per-(element, node) row record of the gather assembly (q1_kernel.cuh): byte t < 27 = offset inside the node's CSR row of
// the column of the neighbour node at (i,j,k) + (t/9-1, t/3%3-1, t%3-1), 0xFF = dropped (outside the element, Dirichlet
// row or column, row not owned); 5 pad bytes.
__global__ void q1_noff_kernel(long nnode_total, int n, int nown, const int *e2l, const int *l2g, const int *rowptr, const int *colidx, unsigned char *noff, int *overflow)
{
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nnode_total * 32) return;
  const long gnode = g >> 5;
  const int t = (int)(g & 31);
  unsigned char rel = t < 27 ? 0xFF : 0;
  if (t < 27) {
    const int n3 = n * n * n;
    const long e = gnode / n3;
    const int nd = (int)(gnode - e * n3), k = nd % n, j = (nd / n) % n, i = nd / (n * n);
    const int ci = i + t / 9 - 1, cj = j + (t / 3) % 3 - 1, ck = k + t % 3 - 1;
    const int row = e2l[gnode];
    if (row < nown && (unsigned)ci < (unsigned)n && (unsigned)cj < (unsigned)n && (unsigned)ck < (unsigned)n) {
      const int col = l2g[e2l[e * n3 + (ci * n + cj) * n + ck]];
      if (col >= 0) {
        const long pos = q1_find(rowptr, colidx, row, col);
        if (pos >= 0) { const long r = pos - rowptr[row]; if (r >= 0xFF) atomicOr(overflow, 1); else rel = (unsigned char)r; }
        else atomicOr(overflow, 2);                          // the pattern does not contain a coupling
      }
    }
  }
  noff[g] = rel;
}
static int q1_build_noff(dB200Op o, const int *drow, const int *dcol, const int *dl2g, long long nnzb, unsigned char **out)
{
  dB200Ctx c = o->ctx;
  dB200Space s = o->space[0];
  const int n = s->P;
  const long nnode = (long)o->nelem * n * n * n;
  unsigned char *noff = nullptr;
  int *ovf = nullptr, hovf = 0;
  *out = nullptr;
  if (!nnode) return 0;
  CU(cudaMalloc((void **)&noff, (size_t)nnode * 32));
  CU(cudaMalloc((void **)&ovf, sizeof(int)));
  CU(cudaMemsetAsync(ovf, 0, sizeof(int), c->stream));
  q1_noff_kernel<<<nblk(nnode * 32, 256), 256, 0, c->stream>>>(nnode, n, s->n, s->d_e2l, dl2g, drow, dcol, noff, ovf);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(&hovf, ovf, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  c->nlaunch += 1;
  cudaFree(ovf);
  if (hovf) {
    cudaFree(noff);
    return seterr(56, hovf & 2 ? "Q1 assembly: the given pattern does not contain every coupling of the Q1 patches" : "Q1 assembly: a row with 255 or more blocks is not supported");
  }
  *out = noff;
  return 0;
}

struct Q1Cache {
  int nrows = 0;
  long long nnzb = 0;
  int *d_rowptr = nullptr, *d_colidx = nullptr;
  unsigned char *d_relpos = nullptr;        // [npatch][8][8] or NULL (a row with >= 255 blocks: fall back to searching)
  unsigned char *d_noff = nullptr;          // [nelem*n^3][32] row records of the gather assembly (built on first use)
};

static int q1_build_pattern(dB200Op o, const int *d_l2g, Q1Cache *pc)
{
  dB200Ctx c = o->ctx;
  dB200Space s = o->space[0];
  const int p = s->P - 1, nown = s->n;
  const long npatch = (long)o->nelem * p * p * p;
  int *cnt = nullptr, *tab = nullptr, *len = nullptr;
  long long *cnt64 = nullptr, *off = nullptr;
  void *tmp = nullptr;
  size_t tmpb = 0;
  CU(cudaMalloc(&cnt, ((size_t)nown + 1) * sizeof(int)));
  CU(cudaMalloc(&cnt64, ((size_t)nown + 1) * sizeof(long long)));
  CU(cudaMalloc(&off, ((size_t)nown + 1) * sizeof(long long)));
  CU(cudaMemsetAsync(cnt, 0, ((size_t)nown + 1) * sizeof(int), c->stream));
  if (npatch) q1_count_kernel<<<nblk(npatch, 256), 256, 0, c->stream>>>(npatch, p, nown, s->d_e2l, cnt);
  widen_kernel<<<nblk(nown + 1, 256), 256, 0, c->stream>>>(nown + 1, cnt, cnt64);
  CU(cub::DeviceScan::ExclusiveSum(nullptr, tmpb, cnt64, off, nown + 1, c->stream));
  CU(cudaMalloc(&tmp, tmpb));
  CU(cub::DeviceScan::ExclusiveSum(tmp, tmpb, cnt64, off, nown + 1, c->stream));
  long long total = 0;
  CU(cudaMemcpyAsync(&total, off + nown, sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaMalloc(&tab, std::max<size_t>(total, 1) * sizeof(int)));
  fill_int_kernel<<<nblk(std::max<long long>(total, 1), 256), 256, 0, c->stream>>>(total, 0x7fffffff, tab);
  if (npatch) q1_insert_kernel<<<nblk(npatch, 128), 128, 0, c->stream>>>(npatch, p, nown, s->d_e2l, d_l2g, off, tab);
  CU(cudaMalloc(&len, ((size_t)nown + 1) * sizeof(int)));
  CU(cudaMemsetAsync(len, 0, ((size_t)nown + 1) * sizeof(int), c->stream));
  if (nown) q1_sort_kernel<<<nblk(nown, 128), 128, 0, c->stream>>>(nown, off, tab, len);
  CU(cudaMalloc(&pc->d_rowptr, ((size_t)nown + 1) * sizeof(int)));
  size_t tmpb2 = 0;
  CU(cub::DeviceScan::ExclusiveSum(nullptr, tmpb2, len, pc->d_rowptr, nown + 1, c->stream));
  if (tmpb2 > tmpb) { CU(cudaFree(tmp)); CU(cudaMalloc(&tmp, tmpb2)); }
  CU(cub::DeviceScan::ExclusiveSum(tmp, tmpb2, len, pc->d_rowptr, nown + 1, c->stream));
  int nnz = 0;
  CU(cudaMemcpyAsync(&nnz, pc->d_rowptr + nown, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaMalloc(&pc->d_colidx, std::max<size_t>(nnz, 1) * sizeof(int)));
  if (nown) q1_compact_kernel<<<nblk(nown, 128), 128, 0, c->stream>>>(nown, off, tab, pc->d_rowptr, pc->d_colidx);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(c->stream));
  c->nlaunch += 8;
  pc->nrows = nown; pc->nnzb = nnz;
  if (npatch) {
    int *ovf = cnt;                                          // reuse as the overflow flag
    CU(cudaMemsetAsync(ovf, 0, sizeof(int), c->stream));
    CU(cudaMalloc((void **)&pc->d_relpos, (size_t)npatch * 64));
    q1_relpos_kernel<<<nblk(npatch * 8, 256), 256, 0, c->stream>>>(npatch, p, nown, s->d_e2l, d_l2g, pc->d_rowptr, pc->d_colidx, pc->d_relpos, ovf);
    int hovf = 0;
    CU(cudaMemcpyAsync(&hovf, ovf, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->nlaunch++;
    if (hovf) { cudaFree(pc->d_relpos); pc->d_relpos = nullptr; }
  }
  cudaFree(cnt); cudaFree(cnt64); cudaFree(off); cudaFree(tab); cudaFree(len); cudaFree(tmp);
  return 0;
}

// ---- assembly ---------------------------------------------------------------------------------------------------------------
struct Q1Args {
  int nelem, p, n_own, nc;
  const int *e2l, *l2g;
  const double *coords, *x, *xaux;
  const double *tab;             // device: Bs[2p][n], Ds[2p][n], qs[2p], ws[2p]
  const int *rowptr, *colidx;
  const unsigned char *relpos;   // cached COO->CSR offsets or NULL (search)
  double *vals;
  double c0, c1, c2;             // ellip: (p-2)/2, eps^2, lambda ; elast: lambda, mu
};

__device__ __forceinline__ double shfl8(double v, int src) { return __shfl_sync(0xffffffffu, v, src, 8); }


template <int KIND> __global__ void __launch_bounds__(256, KIND == 0 ? 2 : 1) q1_assemble_kernel(const Q1Args A)
{
  constexpr int BS = KIND == Q1_ELAST ? 3 : 1;
  extern __shared__ double sm[];
  const int p = A.p, n = p + 1, Qs = 2 * p, n3 = n * n * n;
  const long e = blockIdx.x;
  // shared layout
  double *Bs = sm, *Ds = Bs + Qs * n, *qs = Ds + Qs * n, *ws = qs + Qs, *X = ws + Qs;     // tables, 24 vertex coordinates
  double *U = X + 24;                                 // [BS][n][n][n]
  double *X0 = U + BS * n3, *X1 = X0 + BS * Qs * n * n;   // [BS][Qs][n][n] value / x-derivative after the x contraction
  double *V = X1 + BS * Qs * n * n, *Vx = V + BS * Qs * Qs * n, *Vy = Vx + BS * Qs * Qs * n;   // [BS][Qs][Qs][n]
  const int tid = threadIdx.x, nth = blockDim.x;
  for (int i = tid; i < 2 * Qs * n + 2 * Qs; i += nth) sm[i] = A.tab[i];
  for (int i = tid; i < 24; i += nth) X[i] = A.coords[e * 24 + i];
  for (int i = tid; i < n3; i += nth) {              // dFSGlobalToExpanded(INHOMOGENEOUS), fs.c:543-586
    const int idx = A.e2l[e * n3 + i];
    for (int d = 0; d < BS; d++) U[d * n3 + i] = idx < A.n_own ? A.x[(long)idx * BS + d] : A.xaux[(long)(idx - A.n_own) * BS + d];
  }
  __syncthreads();
  // state on the SPARSE points, sum-factorised: x then y here, z per point below
  for (int it = tid; it < BS * Qs * n * n; it += nth) {
    const int d = it / (Qs * n * n), r = it - d * Qs * n * n, qa = r / (n * n), jk = r - qa * n * n;
    double s0 = 0, s1 = 0;
    for (int i = 0; i < n; i++) { const double u = U[d * n3 + i * n * n + jk]; s0 = fma(Bs[qa * n + i], u, s0); s1 = fma(Ds[qa * n + i], u, s1); }
    X0[it] = s0; X1[it] = s1;
  }
  __syncthreads();
  for (int it = tid; it < BS * Qs * Qs * n; it += nth) {
    const int d = it / (Qs * Qs * n), r = it - d * Qs * Qs * n, qa = r / (Qs * n), r2 = r - qa * Qs * n, qb = r2 / n, k = r2 - qb * n;
    double v = 0, vx = 0, vy = 0;
    for (int j = 0; j < n; j++) {
      const double a0 = X0[(d * Qs + qa) * n * n + j * n + k], a1 = X1[(d * Qs + qa) * n * n + j * n + k];
      v = fma(Bs[qb * n + j], a0, v); vx = fma(Bs[qb * n + j], a1, vx); vy = fma(Ds[qb * n + j], a0, vy);
    }
    V[it] = v; Vx[it] = vx; Vy[it] = vy;
  }
  __syncthreads();

  // ---- patches: 8 lanes per patch; lane l evaluates quadrature point l, then owns trial node l ----------------------------
  const double i1[2][2] = {{0.78867513459481287, 0.21132486540518713}, {0.21132486540518713, 0.78867513459481287}};   // efstopo.c:665
  const int lane = tid & 7, grp = tid >> 3, ngrp = nth >> 3, p3 = p * p * p;
  for (int pt0 = 0; pt0 < p3; pt0 += ngrp) {
    const int pt = min(pt0 + grp, p3 - 1);            // idle groups redo the last patch (keeps the 8-lane shuffles uniform) and skip the scatter
    const bool active = pt0 + grp < p3;
    const int pk = pt % p, pj = (pt / p) % p, pi = pt / (p * p);
    // -- point `lane` of the patch
    const int ii = lane >> 2, jj = (lane >> 1) & 1, kk = lane & 1;
    const int qa = 2 * pi + ii, qb = 2 * pj + jj, qc = 2 * pk + kk;                      // patchind, tensorquad.c:144-158
    double g[BS][3], uval[BS];
    for (int d = 0; d < BS; d++) {
      const double *v = V + ((d * Qs + qa) * Qs + qb) * n, *vx = Vx + ((d * Qs + qa) * Qs + qb) * n, *vy = Vy + ((d * Qs + qa) * Qs + qb) * n;
      double s = 0, sx = 0, sy = 0, sz = 0;
      for (int k = 0; k < n; k++) { const double b = Bs[qc * n + k]; s = fma(b, v[k], s); sx = fma(b, vx[k], sx); sy = fma(b, vy[k], sy); sz = fma(Ds[qc * n + k], v[k], sz); }
      uval[d] = s; g[d][0] = sx; g[d][1] = sy; g[d][2] = sz;
    }
    LineGeom lg;
    PointGeom pg;
    line_geom(X, qs[qa], qs[qb], lg, false);
    point_geom_adj(lg, qs[qc], pg);
    const double jw_own = ws[qa] * ws[qb] * ws[qc] * pg.det;                             // jacobi.c:383-414 at the true point
    double Du[BS][3];
    for (int d = 0; d < BS; d++)
      for (int x = 0; x < 3; x++) Du[d][x] = pg.rdet * (g[d][0] * pg.adj[x] + g[d][1] * pg.adj[3 + x] + g[d][2] * pg.adj[6 + x]);
    // quirk (ii): the patch basis is mapped with the inverse Jacobian of element-natural point number pt*8+lane
    double jq_own[9];
    {
      const int m = pt * 8 + lane, mc = m % Qs, mb = (m / Qs) % Qs, ma = m / (Qs * Qs);
      LineGeom lq;
      PointGeom pq;
      line_geom(X, qs[ma], qs[mb], lq, false);
      point_geom_adj(lq, qs[mc], pq);
      for (int i = 0; i < 9; i++) jq_own[i] = pq.adj[i] * pq.rdet;
    }
    // per-point store of the example (EllipPointwiseComputeStore ellip.c:333-342 / ElastPointwiseComputeStore elast.c:225)
    constexpr int NST = KIND == Q1_ELAST ? 9 : 5;
    double st_own[NST];
    if (KIND == Q1_ELLIP) {
      const double gamma = 0.5 * (Du[0][0] * Du[0][0] + Du[0][1] * Du[0][1] + Du[0][2] * Du[0][2]), espg = A.c1 + gamma;
      const double eta = pow(espg, A.c0), sq = sqrt(-A.c0 * eta / espg);
      st_own[0] = eta; st_own[1] = sq * Du[0][0]; st_own[2] = sq * Du[0][1]; st_own[3] = sq * Du[0][2];
      st_own[4] = A.c2 != 0 ? A.c2 * exp(uval[0]) : 0;
    } else {
      for (int d = 0; d < BS; d++) for (int x = 0; x < 3; x++) st_own[(d * 3 + x) % NST] = (d == x ? 1.0 : 0.0) + Du[d][x];
    }
    // -- trial node `lane`: K[i][fi][fj] for all test nodes i
    double K[8][BS][BS];
    for (int i = 0; i < 8; i++) for (int a = 0; a < BS; a++) for (int b = 0; b < BS; b++) K[i][a][b] = 0;
    const int tj0 = lane >> 2, tj1 = (lane >> 1) & 1, tj2 = lane & 1;
    if (KIND == Q1_ELLIP) {
      // K[i][j] = sum_q  d_i(q)^T A_q d_j(q) - c_q phi_i(q) phi_j(q)   with   G_b = jq^T d_b (efstopo.c:707-709),
      // A_q = jw (eta jq jq^T - t t^T), t = jq s, c_q = jw lambda e^u: EllipPointwiseJacobian (ellip.c:369-377) applied to the
      // unit trial function and tested against (phi_i, G_i) (ellip.c:506-520), regrouped so that the per-point 3x3 tensor is
      // formed once (by the lane that owns the point) and the reference-space basis d_b(q), phi_b(q) are compile-time constants.
      double Aq_own[6], cq_own;
      {
        const double *jq = jq_own;
        const double t0 = jq[0] * st_own[1] + jq[1] * st_own[2] + jq[2] * st_own[3];
        const double t1 = jq[3] * st_own[1] + jq[4] * st_own[2] + jq[5] * st_own[3];
        const double t2 = jq[6] * st_own[1] + jq[7] * st_own[2] + jq[8] * st_own[3];
        const double etaw = st_own[0] * jw_own;
        Aq_own[0] = etaw * (jq[0] * jq[0] + jq[1] * jq[1] + jq[2] * jq[2]) - jw_own * t0 * t0;     // [0][0]
        Aq_own[1] = etaw * (jq[3] * jq[3] + jq[4] * jq[4] + jq[5] * jq[5]) - jw_own * t1 * t1;     // [1][1]
        Aq_own[2] = etaw * (jq[6] * jq[6] + jq[7] * jq[7] + jq[8] * jq[8]) - jw_own * t2 * t2;     // [2][2]
        Aq_own[3] = etaw * (jq[0] * jq[3] + jq[1] * jq[4] + jq[2] * jq[5]) - jw_own * t0 * t1;     // [0][1]
        Aq_own[4] = etaw * (jq[0] * jq[6] + jq[1] * jq[7] + jq[2] * jq[8]) - jw_own * t0 * t2;     // [0][2]
        Aq_own[5] = etaw * (jq[3] * jq[6] + jq[4] * jq[7] + jq[5] * jq[8]) - jw_own * t1 * t2;     // [1][2]
        cq_own = jw_own * st_own[4];
      }
      constexpr double IA = 0.78867513459481287, IB = 0.21132486540518713;        // interp1, efstopo.c:665
      // this lane's 1-D factors at point bit 0/1 in each direction, and the sign of deriv1 = {-1,+1}[b]
      const double ax[2] = {tj0 ? IB : IA, tj0 ? IA : IB}, ay[2] = {tj1 ? IB : IA, tj1 ? IA : IB}, az[2] = {tj2 ? IB : IA, tj2 ? IA : IB};
      const double sx = tj0 ? 1.0 : -1.0, sy = tj1 ? 1.0 : -1.0, sz = tj2 ? 1.0 : -1.0;
#pragma unroll
      for (int q = 0; q < 8; q++) {
        const int q0 = q >> 2, q1 = (q >> 1) & 1, q2 = q & 1;
        double Aq[6];
#pragma unroll
        for (int i = 0; i < 6; i++) Aq[i] = shfl8(Aq_own[i], q);
        const double cq = shfl8(cq_own, q);
        const double dj0 = sx * ay[q1] * az[q2], dj1 = ax[q0] * sy * az[q2], dj2 = ax[q0] * ay[q1] * sz, phij = ax[q0] * ay[q1] * az[q2];
        const double w0 = Aq[0] * dj0 + Aq[3] * dj1 + Aq[4] * dj2, w1 = Aq[3] * dj0 + Aq[1] * dj1 + Aq[5] * dj2, w2 = Aq[4] * dj0 + Aq[5] * dj1 + Aq[2] * dj2;
        const double cphi = cq * phij;
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const int b0 = i >> 2, b1 = (i >> 1) & 1, b2 = i & 1;
          const double ix = (q0 == b0) ? IA : IB, iy = (q1 == b1) ? IA : IB, iz = (q2 == b2) ? IA : IB;      // interp1[q][b]
          const double di0 = (b0 ? 1.0 : -1.0) * iy * iz, di1 = ix * (b1 ? 1.0 : -1.0) * iz, di2 = ix * iy * (b2 ? 1.0 : -1.0), phii = ix * iy * iz;
          K[i][0][0] += di0 * w0 + di1 * w1 + di2 * w2 - phii * cphi;
        }
      }
    } else {
#pragma unroll 1
    for (int q = 0; q < 8; q++) {
      const double jw = shfl8(jw_own, q);
      double jq[9], st[NST];
#pragma unroll
      for (int i = 0; i < 9; i++) jq[i] = shfl8(jq_own[i], q);
#pragma unroll
      for (int i = 0; i < NST; i++) st[i] = shfl8(st_own[i], q);
      const int q0 = q >> 2, q1 = (q >> 1) & 1, q2 = q & 1;
      // basis function b at point q: interp1/deriv1 products (efstopo.c:699-706); deriv1[..][0] = -1, [..][1] = +1
      auto basis = [&](int b0, int b1, int b2, double &phi, double (&G)[3]) {
        const double ix = i1[q0][b0], iy = i1[q1][b1], iz = i1[q2][b2];
        const double dX = (b0 ? 1.0 : -1.0) * iy * iz, dY = ix * (b1 ? 1.0 : -1.0) * iz, dZ = ix * iy * (b2 ? 1.0 : -1.0);
        phi = ix * iy * iz;
#pragma unroll
        for (int x = 0; x < 3; x++) G[x] = dX * jq[x] + dY * jq[3 + x] + dZ * jq[6 + x];
      };
      double phij, Gj[3];
      basis(tj0, tj1, tj2, phij, Gj);
      {                                                // ElastPointwiseJacobian on unit trial functions, elast.c:270-280,364-393
        double F[9], S[6];
#pragma unroll
        for (int i = 0; i < 9; i++) F[i] = st[i % NST];
        // S = lambda tr(E) 1 + 2 mu E, E = (F^T F - 1)/2 (elast.c:227-241)
        {
          double E[6];
          E[0] = 0.5 * (F[0] * F[0] + F[3] * F[3] + F[6] * F[6] - 1); E[1] = 0.5 * (F[1] * F[1] + F[4] * F[4] + F[7] * F[7] - 1);
          E[2] = 0.5 * (F[2] * F[2] + F[5] * F[5] + F[8] * F[8] - 1); E[3] = 0.5 * (F[0] * F[1] + F[3] * F[4] + F[6] * F[7]);
          E[4] = 0.5 * (F[0] * F[2] + F[3] * F[5] + F[6] * F[8]); E[5] = 0.5 * (F[1] * F[2] + F[4] * F[5] + F[7] * F[8]);
          const double tr = E[0] + E[1] + E[2];
          for (int i = 0; i < 6; i++) S[i] = (i < 3 ? A.c0 * tr : 0.0) + 2 * A.c1 * E[i];
        }
#pragma unroll
        for (int fj = 0; fj < BS; fj++) {
          // dF = e_fj (x) Gj ; dE = sym(F^T dF) ; dS ; Dv = jw (F dS + dF S)
          double dE[6], dS[6], Dv[9];
          const double *Frow = F + 3 * fj;             // (F^T dF)[a][b] = F[fj][a] Gj[b]
          dE[0] = Frow[0] * Gj[0]; dE[1] = Frow[1] * Gj[1]; dE[2] = Frow[2] * Gj[2];
          dE[3] = 0.5 * (Frow[0] * Gj[1] + Frow[1] * Gj[0]); dE[4] = 0.5 * (Frow[0] * Gj[2] + Frow[2] * Gj[0]); dE[5] = 0.5 * (Frow[1] * Gj[2] + Frow[2] * Gj[1]);
          const double dtr = dE[0] + dE[1] + dE[2];
          for (int i = 0; i < 6; i++) dS[i] = (i < 3 ? A.c0 * dtr : 0.0) + 2 * A.c1 * dE[i];
          for (int i = 0; i < 3; i++) {
            Dv[i * 3 + 0] = F[i * 3] * dS[0] + F[i * 3 + 1] * dS[3] + F[i * 3 + 2] * dS[4];
            Dv[i * 3 + 1] = F[i * 3] * dS[3] + F[i * 3 + 1] * dS[1] + F[i * 3 + 2] * dS[5];
            Dv[i * 3 + 2] = F[i * 3] * dS[4] + F[i * 3 + 1] * dS[5] + F[i * 3 + 2] * dS[2];
          }
          Dv[fj * 3 + 0] += Gj[0] * S[0] + Gj[1] * S[3] + Gj[2] * S[4];
          Dv[fj * 3 + 1] += Gj[0] * S[3] + Gj[1] * S[1] + Gj[2] * S[5];
          Dv[fj * 3 + 2] += Gj[0] * S[4] + Gj[1] * S[5] + Gj[2] * S[2];
#pragma unroll
          for (int i = 0; i < 8; i++) {
            double phii, Gi[3];
            basis(i >> 2, (i >> 1) & 1, i & 1, phii, Gi);
#pragma unroll
            for (int fi = 0; fi < BS; fi++) K[i][fi][fj % BS] += jw * (Gi[0] * Dv[fi * 3] + Gi[1] * Dv[fi * 3 + 1] + Gi[2] * Dv[fi * 3 + 2]);
          }
        }
      }
    }
    }
    // -- dFSMatSetValuesBlockedExpanded, fs.c:826-930: expanded -> local (boolean E) -> global; Dirichlet rows/cols dropped
    {
      const int locj = A.e2l[e * n3 + patch_corner(n, pi, pj, pk, lane)];
      if (A.relpos) {
        const int rp_own = locj < A.n_own ? A.rowptr[locj] : 0;       // lane i holds rowptr of test corner i
        const unsigned long long word = reinterpret_cast<const unsigned long long *>(A.relpos)[(e * p3 + pt) * 8 + lane];
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const int rp = __shfl_sync(0xffffffffu, rp_own, i, 8);
          const unsigned int rel = (unsigned int)(word >> (8 * i)) & 0xFF;
          if (active && rel != 0xFF) {
            const long pos = (long)rp + rel;
            for (int fi = 0; fi < BS; fi++) for (int fj = 0; fj < BS; fj++) atomicAdd(A.vals + (pos * BS + fi) * BS + fj, K[i][fi][fj]);
          }
        }
      } else {
        const int col = A.l2g[locj];
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const int row = __shfl_sync(0xffffffffu, locj, i, 8);        // all lanes shuffle; only active groups add
          if (active && col >= 0 && row < A.n_own) {
            const long pos = q1_find(A.rowptr, A.colidx, row, col);
            if (pos >= 0)                                               // always true with a pattern from dB200OpQ1Pattern
              for (int fi = 0; fi < BS; fi++) for (int fj = 0; fj < BS; fj++) atomicAdd(A.vals + (pos * BS + fi) * BS + fj, K[i][fi][fj]);
          }
        }
      }
    }
  }
}

// ---- API ------------------------------------------------------------------------------------------------------------------------
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

extern "C" int dB200OpQ1Pattern(dB200Op o, const int *l2g, long long *nrows, long long *nnz, int *rowptr, int *colidx)
{
  dB200Ctx c = o->ctx;
  CU(cudaSetDevice(c->device));
  dB200Space s = o->space[0];
  if (o->kind != dB200_OP_ELLIP && o->kind != dB200_OP_ELAST) return seterr(56, "Q1 pattern: operator kind %d not implemented", o->kind);
  Q1Cache *pc = (Q1Cache *)o->q1cache;
  if (!pc) {
    const int *dl2g;
    int *own;
    CHK(to_dev(c, l2g, (size_t)s->nc + s->ngh, &dl2g, &own));
    pc = new Q1Cache;
    int err = q1_build_pattern(o, dl2g, pc);
    if (own) cudaFree(own);
    if (err) { delete pc; return err; }
    o->q1cache = pc;
  }
  if (nrows) *nrows = pc->nrows;
  if (nnz) *nnz = pc->nnzb;
  if (rowptr) CU(cudaMemcpyAsync(rowptr, pc->d_rowptr, ((size_t)pc->nrows + 1) * sizeof(int), is_dev(rowptr) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
  if (colidx) CU(cudaMemcpyAsync(colidx, pc->d_colidx, (size_t)pc->nnzb * sizeof(int), is_dev(colidx) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

void dohp_q1_free(dB200Op o)
{
  Q1Cache *pc = (Q1Cache *)o->q1cache;
  if (!pc) return;
  cudaFree(pc->d_rowptr); cudaFree(pc->d_colidx);
  if (pc->d_relpos) cudaFree(pc->d_relpos);
  if (pc->d_noff) cudaFree(pc->d_noff);
  if (o->d_q1tab) cudaFree(o->d_q1tab);
  delete pc;
  o->q1cache = nullptr;
}

extern "C" int dB200OpAssembleQ1(dB200Op o, const double *const gx[], const int *l2g, const int *rowptr, const int *colidx, double *vals)
{
  dB200Ctx c = o->ctx;
  CU(cudaSetDevice(c->device));
  dB200Space s = o->space[0];
  if (o->kind != dB200_OP_ELLIP && o->kind != dB200_OP_ELAST) return seterr(56, "Q1 assembly: operator kind %d not implemented", o->kind);
  const int bs = s->bs, p = s->P - 1, n = p + 1, Qs = 2 * p;
  // pattern: the caller's (any block CSR with sorted columns that contains the couplings) or the cached one
  Q1Cache *pc = (Q1Cache *)o->q1cache;
  const int *drow = nullptr, *dcol = nullptr, *dl2g = nullptr;
  int *own_row = nullptr, *own_col = nullptr, *own_l2g = nullptr;
  long long nnzb;
  if (rowptr && colidx) {
    int last;
    CU(cudaMemcpy(&last, rowptr + s->n, sizeof(int), is_dev(rowptr) ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost));
    nnzb = last;
    CHK(to_dev(c, rowptr, (size_t)s->n + 1, &drow, &own_row));
    CHK(to_dev(c, colidx, (size_t)nnzb, &dcol, &own_col));
  } else {
    if (!pc) return seterr(73, "Q1 assembly without a pattern: call dB200OpQ1Pattern first or pass rowptr/colidx");
    drow = pc->d_rowptr; dcol = pc->d_colidx; nnzb = pc->nnzb;
  }
  const unsigned char *drel = (!(rowptr && colidx) && pc) ? pc->d_relpos : nullptr;
  CHK(to_dev(c, l2g, (size_t)s->nc + s->ngh, &dl2g, &own_l2g));
  // SPARSE rule and the high-order basis tabulated on it (EllipGetRegionIterator(EVAL_JACOBIAN), ellip.c:289-316)
  if (!o->d_q1tab) {
    std::vector<double> qs(2 * p + 2), ws(2 * p + 2), tab((size_t)2 * Qs * n + 2 * Qs);
    int size;
    CHK(rule_1d(dB200_QMETHOD_SPARSE, dB200_GAUSS_LEGENDRE, 2 * p, &size, qs.data(), ws.data()));
    if (size != Qs) return seterr(77, "SPARSE rule size %d != %d", size, Qs);
    CHK(basis_1d(n, Qs, qs.data(), tab.data(), tab.data() + (size_t)Qs * n, nullptr, nullptr));
    for (int i = 0; i < Qs; i++) { tab[(size_t)2 * Qs * n + i] = qs[i]; tab[(size_t)2 * Qs * n + Qs + i] = ws[i]; }
    CU(cudaMalloc((void **)&o->d_q1tab, tab.size() * sizeof(double)));
    CU(cudaMemcpyAsync(o->d_q1tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  const size_t nb = (size_t)s->n * bs;
  const double *dx;
  double *own_x = nullptr, *dvals, *own_vals = nullptr;
  CHK(to_dev(c, gx[0], nb, &dx, &own_x));
  const size_t nvals = (size_t)nnzb * bs * bs;
  if (is_dev(vals)) dvals = vals;
  else { CU(cudaMalloc((void **)&own_vals, std::max<size_t>(nvals, 1) * sizeof(double))); dvals = own_vals; }
  CU(cudaMemsetAsync(dvals, 0, nvals * sizeof(double), c->stream));                     // MatZeroEntries, ellip.c:489
  if (!s->d_xaux_inhomog) CHK(dB200OpSetDirichlet(o, 0, nullptr));
  if (s->ngh) return seterr(56, "Q1 assembly on a partitioned space is not implemented (off-process rows need the MatAssembly stash exchange)");
  unsigned char *tmp_noff = nullptr;
  if (o->kind == dB200_OP_ELLIP) {
    // gather assembly (q1_kernel.cuh): row records for the cached pattern are kept, for a caller's pattern built per call
    const unsigned char *dnoff;
    if (!(rowptr && colidx)) {
      if (!pc->d_noff) CHK(q1_build_noff(o, drow, dcol, dl2g, nnzb, &pc->d_noff));
      dnoff = pc->d_noff;
    } else {
      CHK(q1_build_noff(o, drow, dcol, dl2g, nnzb, &tmp_noff));
      dnoff = tmp_noff;
    }
    static const Q1Fn q1fn[] = {launch_q1_2, launch_q1_3, launch_q1_4, launch_q1_5, launch_q1_6, launch_q1_7, launch_q1_8, launch_q1_9, launch_q1_10, launch_q1_11};
    if (p < 1 || p > 10) return seterr(56, "Q1 assembly: degree %d not instantiated", p);
    Q1Launch L;
    L.P = p; L.kind = Q1_ELLIP; L.nelem = o->nelem; L.n_own = s->n; L.e2l = s->d_e2l; L.coords = o->d_coords; L.x = dx; L.xaux = s->d_xaux_inhomog;
    L.tab = o->d_q1tab; L.rowptr = drow; L.noff = dnoff; L.vals = dvals;
    L.dbg = getenv("DOHP_Q1_DBG") ? atoi(getenv("DOHP_Q1_DBG")) : 0;
    L.c0 = (o->ellip.p - 2) / 2; L.c1 = o->ellip.epsilon * o->ellip.epsilon; L.c2 = o->ellip.lambda;
    if (o->nelem) {
      const int err = q1fn[p - 1](L, c->stream);
      if (err) return seterr(700 + err, "Q1 assembly kernel launch failed (degree %d): %s", p, err > 0 ? cudaGetErrorString((cudaError_t)err) : "not instantiated");
      c->nlaunch++;
    }
  } else {
  Q1Args A;
  A.nelem = o->nelem; A.p = p; A.n_own = s->n; A.nc = s->nc; A.e2l = s->d_e2l; A.l2g = dl2g; A.coords = o->d_coords;
  A.x = dx; A.xaux = s->d_xaux_inhomog; A.tab = o->d_q1tab; A.rowptr = drow; A.colidx = dcol; A.relpos = drel; A.vals = dvals;
  A.c0 = o->elast.lambda; A.c1 = o->elast.mu; A.c2 = 0;
  const size_t smem = sizeof(double) * ((size_t)2 * Qs * n + 2 * Qs + 24 + (size_t)bs * (n * n * n + 2 * Qs * n * n + 3 * Qs * Qs * n));
  if (smem > 227 * 1024) return seterr(56, "Q1 assembly: degree %d with block size %d needs %zu B of shared memory per element (> 227 KB)", p, bs, smem);
  const int threads = p <= 2 ? 64 : (p <= 4 ? 128 : 256);
  if (o->nelem) {
    CU(cudaFuncSetAttribute(q1_assemble_kernel<Q1_ELAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    q1_assemble_kernel<Q1_ELAST><<<o->nelem, threads, smem, c->stream>>>(A);
    CU(cudaGetLastError());
    c->nlaunch++;
  }
  }
  if (own_vals) CU(cudaMemcpyAsync(vals, dvals, nvals * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (own_row) cudaFree(own_row);
  if (own_col) cudaFree(own_col);
  if (own_l2g) cudaFree(own_l2g);
  if (own_x) cudaFree(own_x);
  if (own_vals) cudaFree(own_vals);
  if (tmp_noff) cudaFree(tmp_noff);
  return 0;
}
