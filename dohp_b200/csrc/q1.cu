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

// Row of the (extended) local matrix that local node `loc` owns: owned explicit nodes first, then this rank's ghost
// nodes (their rows collect what MatSetValues would stash for the owner until MatAssemblyEnd); Dirichlet nodes have no
// row (negative ltog, fsbuild.c:31-34).  Local layout [explicit n | dirichlet nc-n | ghost ngh], vecd.c:123-151.
__device__ __forceinline__ int q1_row(int loc, int n, int nc) { return loc < n ? loc : (loc >= nc ? n + (loc - nc) : -1); }

__global__ void q1_count_kernel(long npatch_total, int p, int nown, int nc, const int *e2l, int *cnt)
{
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= npatch_total) return;
  const int n = p + 1, p3 = p * p * p;
  const long e = g / p3;
  const int pt = (int)(g - e * p3), pk = pt % p, pj = (pt / p) % p, pi = pt / (p * p);
  for (int b = 0; b < 8; b++) {
    const int row = q1_row(e2l[e * n * n * n + patch_corner(n, pi, pj, pk, b)], nown, nc);
    if (row >= 0) atomicAdd(cnt + row, 8);                   // upper bound: 8 columns per patch containing the row
  }
}
__global__ void q1_insert_kernel(long npatch_total, int p, int nown, int nc, const int *e2l, const int *l2g, const long long *off, int *tab)
{
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= npatch_total) return;
  const int n = p + 1, p3 = p * p * p;
  const long e = g / p3;
  const int pt = (int)(g - e * p3), pk = pt % p, pj = (pt / p) % p, pi = pt / (p * p);
  int loc[8], col[8];
  for (int b = 0; b < 8; b++) { loc[b] = e2l[e * n * n * n + patch_corner(n, pi, pj, pk, b)]; col[b] = l2g[loc[b]]; }
  for (int i = 0; i < 8; i++) {
    const int row = q1_row(loc[i], nown, nc);
    if (row < 0) continue;                                   // Dirichlet rows dropped (fsbuild.c:31-34)
    int *seg = tab + off[row];
    const int cap = (int)(off[row + 1] - off[row]);
    for (int j = 0; j < 8; j++) {
      if (col[j] < 0) continue;                              // Dirichlet columns dropped
      // open addressing inside the row's segment: start at a hash of the column so that a probe sequence is O(1) long
      // (the segment is at most half full: cap = 8 slots per patch containing the node, <= 27 distinct columns per 8 patches)
      int sl = (int)(((unsigned)col[j] * 2654435761u) % (unsigned)cap);
      for (int s = 0; s < cap; s++) {
        const int old = atomicCAS(seg + sl, 0x7fffffff, col[j]);
        if (old == 0x7fffffff || old == col[j]) break;
        sl = sl + 1 == cap ? 0 : sl + 1;
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
  for (int a = 0; a < cap; a++) { const int v = seg[a]; if (v != 0x7fffffff) seg[m++] = v; }      // gather the occupied slots (m <= a)
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
// Per-(element, node) row record of the gather assembly (q1_kernel.cuh): byte t < 27 = offset inside the node's CSR row of
// the column of the neighbour node at (i,j,k) + (t/9-1, t/3%3-1, t%3-1), 0xFF = dropped (outside the element, Dirichlet
// row or column, row not owned); byte 27 pad; bytes 28..31 = CSR start of the node's row (-1: no row).
__global__ void q1_noff_kernel(long nnode_total, int n, int nown, int nc, const int *e2l, const int *l2g, const int *rowptr, const int *colidx, unsigned char *noff, int *overflow)
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
    const int row = q1_row(e2l[gnode], nown, nc);
    if (row >= 0 && (unsigned)ci < (unsigned)n && (unsigned)cj < (unsigned)n && (unsigned)ck < (unsigned)n) {
      const int col = l2g[e2l[e * n3 + (ci * n + cj) * n + ck]];
      if (col >= 0) {
        const long pos = q1_find(rowptr, colidx, row, col);
        if (pos >= 0) { const long r = pos - rowptr[row]; if (r >= 0xFF) atomicOr(overflow, 1); else rel = (unsigned char)r; }
        else atomicOr(overflow, 2);                          // the pattern does not contain a coupling
      }
    }
  }
  if (t < 28) noff[g] = rel;
  else if (t == 28) {                                        // bytes 28..31: CSR start of the node's row (-1: Dirichlet node, no row)
    const int row = q1_row(e2l[gnode], nown, nc);
    reinterpret_cast<int *>(noff)[gnode * 8 + 7] = row >= 0 ? rowptr[row] : -1;
  }
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
  q1_noff_kernel<<<nblk(nnode * 32, 256), 256, 0, c->stream>>>(nnode, n, s->n, s->nc, s->d_e2l, dl2g, drow, dcol, noff, ovf);
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
  unsigned char *d_noff = nullptr;          // [nelem*n^3][32] row records of the gather assembly (built on first use)
  unsigned long long l2g_hash = 0;          // the pattern and the row records belong to ONE local-to-global map
};

// order-dependent checksum of an index array (device or host pointer): a cached pattern is only valid for the l2g it was built from
__global__ void q1_hash_kernel(long n, const int *a, unsigned long long *h)
{
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long v = 0;
  if (g < n) v = ((unsigned long long)(unsigned)a[g] + 0x9E3779B97F4A7C15ull) * (2ull * (unsigned long long)g + 1ull);
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0 && v) atomicAdd(h, v);
}
// frees device temporaries on every exit path
struct DevTmp {
  std::vector<void *> p;
  template <class T> T *keep(T *q) { if (q) p.push_back((void *)q); return q; }
  ~DevTmp() { for (void *q : p) cudaFree(q); }
};
static int q1_hash(dB200Ctx c, const int *d_l2g, size_t n, unsigned long long *h)
{
  unsigned long long *dh;
  CU(cudaMalloc((void **)&dh, sizeof *dh));
  DevTmp tmp;
  tmp.keep(dh);
  CU(cudaMemsetAsync(dh, 0, sizeof *dh, c->stream));
  if (n) q1_hash_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>((long)n, d_l2g, dh);
  CU(cudaGetLastError());
  c->nlaunch++;
  CU(cudaMemcpyAsync(h, dh, sizeof *dh, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

static int q1_build_pattern(dB200Op o, const int *d_l2g, Q1Cache *pc)
{
  dB200Ctx c = o->ctx;
  dB200Space s = o->space[0];
  const int p = s->P - 1, nrows = s->n + s->ngh;       // owned rows, then the rows of this rank's ghost nodes
  const long npatch = (long)o->nelem * p * p * p;
  int *cnt = nullptr, *tab = nullptr, *len = nullptr;
  long long *cnt64 = nullptr, *off = nullptr;
  void *tmp = nullptr;
  size_t tmpb = 0;
  CU(cudaMalloc(&cnt, ((size_t)nrows + 1) * sizeof(int)));
  CU(cudaMalloc(&cnt64, ((size_t)nrows + 1) * sizeof(long long)));
  CU(cudaMalloc(&off, ((size_t)nrows + 1) * sizeof(long long)));
  CU(cudaMemsetAsync(cnt, 0, ((size_t)nrows + 1) * sizeof(int), c->stream));
  if (npatch) q1_count_kernel<<<nblk(npatch, 256), 256, 0, c->stream>>>(npatch, p, s->n, s->nc, s->d_e2l, cnt);
  widen_kernel<<<nblk(nrows + 1, 256), 256, 0, c->stream>>>(nrows + 1, cnt, cnt64);
  CU(cub::DeviceScan::ExclusiveSum(nullptr, tmpb, cnt64, off, nrows + 1, c->stream));
  CU(cudaMalloc(&tmp, tmpb));
  CU(cub::DeviceScan::ExclusiveSum(tmp, tmpb, cnt64, off, nrows + 1, c->stream));
  long long total = 0;
  CU(cudaMemcpyAsync(&total, off + nrows, sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaMalloc(&tab, std::max<size_t>(total, 1) * sizeof(int)));
  fill_int_kernel<<<nblk(std::max<long long>(total, 1), 256), 256, 0, c->stream>>>(total, 0x7fffffff, tab);
  if (npatch) q1_insert_kernel<<<nblk(npatch, 128), 128, 0, c->stream>>>(npatch, p, s->n, s->nc, s->d_e2l, d_l2g, off, tab);
  CU(cudaMalloc(&len, ((size_t)nrows + 1) * sizeof(int)));
  CU(cudaMemsetAsync(len, 0, ((size_t)nrows + 1) * sizeof(int), c->stream));
  if (nrows) q1_sort_kernel<<<nblk(nrows, 128), 128, 0, c->stream>>>(nrows, off, tab, len);
  CU(cudaMalloc(&pc->d_rowptr, ((size_t)nrows + 1) * sizeof(int)));
  size_t tmpb2 = 0;
  CU(cub::DeviceScan::ExclusiveSum(nullptr, tmpb2, len, pc->d_rowptr, nrows + 1, c->stream));
  if (tmpb2 > tmpb) { CU(cudaFree(tmp)); CU(cudaMalloc(&tmp, tmpb2)); }
  CU(cub::DeviceScan::ExclusiveSum(tmp, tmpb2, len, pc->d_rowptr, nrows + 1, c->stream));
  int nnz = 0;
  CU(cudaMemcpyAsync(&nnz, pc->d_rowptr + nrows, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  // the same sum in 64 bits: rowptr/colidx are 32-bit like the reference's dInt (include/dohptype.h:30-34); fail loudly instead of wrapping
  long long nnz64 = 0;
  widen_kernel<<<nblk(nrows + 1, 256), 256, 0, c->stream>>>(nrows + 1, len, cnt64);
  CU(cub::DeviceScan::ExclusiveSum(tmp, tmpb, cnt64, cnt64, nrows + 1, c->stream));
  CU(cudaMemcpyAsync(&nnz64, cnt64 + nrows, sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (nnz64 >= (1LL << 31)) {
    cudaFree(cnt); cudaFree(cnt64); cudaFree(off); cudaFree(tab); cudaFree(len); cudaFree(tmp); cudaFree(pc->d_rowptr);
    pc->d_rowptr = nullptr;
    return seterr(63, "Q1 pattern: %lld blocks do not fit 32-bit row offsets (dInt = int); partition the space over more ranks", nnz64);
  }
  CU(cudaMalloc(&pc->d_colidx, std::max<size_t>(nnz, 1) * sizeof(int)));
  if (nrows) q1_compact_kernel<<<nblk(nrows, 128), 128, 0, c->stream>>>(nrows, off, tab, pc->d_rowptr, pc->d_colidx);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(c->stream));
  c->nlaunch += 9;
  pc->nrows = nrows; pc->nnzb = nnz;
  cudaFree(cnt); cudaFree(cnt64); cudaFree(off); cudaFree(tab); cudaFree(len); cudaFree(tmp);
  return 0;
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

// SPARSE rule and the high-order basis tabulated on it (XGetRegionIterator(EVAL_JACOBIAN), ellip.c:289-316): device table
// Bs[2p][n], Ds[2p][n], qs[2p], ws[2p]
static int q1_tables(dB200Op o)
{
  dB200Ctx c = o->ctx;
  const int p = o->space[0]->P - 1, n = p + 1, Qs = 2 * p;
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
  return 0;
}

extern "C" int dB200OpQ1Pattern(dB200Op o, const int *l2g, long long *nrows, long long *nnz, int *rowptr, int *colidx)
{
  dB200Ctx c = o->ctx;
  CU(cudaSetDevice(c->device));
  dB200Space s = o->space[0];
  if (o->kind != dB200_OP_ELLIP && o->kind != dB200_OP_ELAST && o->kind != dB200_OP_STOKES && o->kind != dB200_OP_PROJECTION) return seterr(56, "Q1 pattern: operator kind %d not implemented", o->kind);
  Q1Cache *pc = (Q1Cache *)o->q1cache;
  {
    const int *dl2g;
    int *own;
    CHK(to_dev(c, l2g, (size_t)s->nc + s->ngh, &dl2g, &own));
    DevTmp tmp;
    tmp.keep(own);
    unsigned long long h;
    CHK(q1_hash(c, dl2g, (size_t)s->nc + s->ngh, &h));
    if (pc && pc->l2g_hash != h) { dohp_q1_free(o); pc = nullptr; }            // a different numbering: the cached pattern and row records are stale
    if (!pc) {
      bool ghost_cols = s->ngh == 0;
      if (!ghost_cols) {                                                     // all ghost entries negative = built with "skip ghost l2g": columns would be dropped silently
        std::vector<int> gh(s->ngh);
        CU(cudaMemcpy(gh.data(), dl2g + s->nc, (size_t)s->ngh * sizeof(int), cudaMemcpyDeviceToHost));
        for (int v : gh) if (v >= 0) { ghost_cols = true; break; }
      }
      if (!ghost_cols) return seterr(63, "Q1 pattern: every ghost entry of l2g is negative (map built without ghost numbers): ghost columns would be dropped");
      pc = new Q1Cache;
      int err = q1_build_pattern(o, dl2g, pc);
      if (err) { delete pc; return err; }
      pc->l2g_hash = h;
      o->q1cache = pc;
    }
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
  dohp_q1owner_free(o);                 // the owner-side maps refer to the cached pattern
  if (o->d_q1tab) { cudaFree(o->d_q1tab); o->d_q1tab = nullptr; }
  Q1Cache *pc = (Q1Cache *)o->q1cache;
  if (!pc) return;
  cudaFree(pc->d_rowptr); cudaFree(pc->d_colidx);
  if (pc->d_noff) cudaFree(pc->d_noff);
  delete pc;
  o->q1cache = nullptr;
}

extern "C" int dB200OpAssembleQ1(dB200Op o, const double *const gx[], const int *l2g, const int *rowptr, const int *colidx, double *vals)
{
  dB200Ctx c = o->ctx;
  CU(cudaSetDevice(c->device));
  dB200Space s = o->space[0];
  if (o->kind != dB200_OP_ELLIP && o->kind != dB200_OP_ELAST && o->kind != dB200_OP_STOKES && o->kind != dB200_OP_PROJECTION) return seterr(56, "Q1 assembly: operator kind %d not implemented", o->kind);
  const int bs = s->bs, p = s->P - 1;
  if (p < 1 || p > 10) return seterr(56, "Q1 assembly: degree %d not instantiated (1..10)", p);
  if (!gx || !gx[0] || !l2g || !vals) return seterr(85, "Q1 assembly: null argument");
  DevTmp tmp;
  // pattern: the caller's (any block CSR with sorted columns that contains the couplings) or the cached one
  Q1Cache *pc = (Q1Cache *)o->q1cache;
  const int *drow = nullptr, *dcol = nullptr, *dl2g = nullptr;
  int *own_row = nullptr, *own_col = nullptr, *own_l2g = nullptr;
  long long nnzb;
  if (rowptr && colidx) {
    int last;
    CU(cudaMemcpy(&last, rowptr + s->n + s->ngh, sizeof(int), is_dev(rowptr) ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost));
    nnzb = last;
    CHK(to_dev(c, rowptr, (size_t)s->n + s->ngh + 1, &drow, &own_row));
    tmp.keep(own_row);
    CHK(to_dev(c, colidx, (size_t)nnzb, &dcol, &own_col));
    tmp.keep(own_col);
  } else {
    if (!pc) return seterr(73, "Q1 assembly without a pattern: call dB200OpQ1Pattern first or pass rowptr/colidx");
    drow = pc->d_rowptr; dcol = pc->d_colidx; nnzb = pc->nnzb;
  }
  CHK(to_dev(c, l2g, (size_t)s->nc + s->ngh, &dl2g, &own_l2g));
  tmp.keep(own_l2g);
  if (!(rowptr && colidx)) {
    unsigned long long h;
    CHK(q1_hash(c, dl2g, (size_t)s->nc + s->ngh, &h));
    if (h != pc->l2g_hash) return seterr(73, "Q1 assembly: l2g differs from the map the cached pattern was built with; call dB200OpQ1Pattern with the new map first");
  }
  CHK(q1_tables(o));
  const size_t nb = (size_t)s->n * bs;
  const double *dx;
  double *own_x = nullptr, *dvals, *own_vals = nullptr;
  CHK(to_dev(c, gx[0], nb, &dx, &own_x));
  tmp.keep(own_x);
  const size_t nvals = (size_t)nnzb * bs * bs;
  if (is_dev(vals)) dvals = vals;
  else { CU(cudaMalloc((void **)&own_vals, std::max<size_t>(nvals, 1) * sizeof(double))); dvals = tmp.keep(own_vals); }
  CU(cudaMemsetAsync(dvals, 0, nvals * sizeof(double), c->stream));                     // MatZeroEntries, ellip.c:489
  if (!s->d_xaux_inhomog) CHK(dB200OpSetDirichlet(o, 0, nullptr));
  CHK(dohp_halo_forward(s, dx, s->d_xaux_inhomog));                                     // ghost values of the state (VecGhostUpdate INSERT, fs.c:578-579)
  // gather assembly (q1_kernel.cuh / q1_elast_kernel.cuh): row records of the cached pattern are kept, those of a caller's
  // pattern are built per call
  unsigned char *tmp_noff = nullptr;
  const unsigned char *dnoff;
  if (!(rowptr && colidx)) {
    if (!pc->d_noff) CHK(q1_build_noff(o, drow, dcol, dl2g, nnzb, &pc->d_noff));
    dnoff = pc->d_noff;
  } else {
    CHK(q1_build_noff(o, drow, dcol, dl2g, nnzb, &tmp_noff));
    dnoff = tmp.keep(tmp_noff);
  }
  static const Q1Fn q1fn[] = {launch_q1_2, launch_q1_3, launch_q1_4, launch_q1_5, launch_q1_6, launch_q1_7, launch_q1_8, launch_q1_9, launch_q1_10, launch_q1_11};
  Q1Launch L;
  memset(&L, 0, sizeof L);
  L.P = p; L.kind = o->kind == dB200_OP_ELLIP || o->kind == dB200_OP_PROJECTION ? Q1_ELLIP : o->kind == dB200_OP_ELAST ? Q1_ELAST : Q1_STOKES_A; L.nelem = o->nelem; L.n_own = s->n; L.nc = s->nc; L.e2l = s->d_e2l; L.coords = o->d_coords;
  L.x = dx; L.xaux = s->d_xaux_inhomog; L.tab = o->d_q1tab; L.rowptr = drow; L.noff = dnoff; L.vals = dvals;
  L.reduced = o->q1reduced;
#ifdef DOHP_PROFILE
  L.dbg = getenv("DOHP_Q1_DBG") ? atoi(getenv("DOHP_Q1_DBG")) : 0;      // measurement ablations, profiling builds only
#endif
  if (o->kind == dB200_OP_PROJECTION) { L.c0 = 0; L.c1 = 1; L.c2 = -1; L.mass_only = 1; }      // K[i][j] = sum_q phi_i jw phi_j, ex1.c:458-464
  else if (o->kind == dB200_OP_ELLIP) { L.c0 = (o->ellip.p - 2) / 2; L.c1 = o->ellip.epsilon * o->ellip.epsilon; L.c2 = o->ellip.lambda; }
  else if (o->kind == dB200_OP_ELAST) { L.c0 = o->elast.lambda; L.c1 = o->elast.mu; L.c2 = 0; }
  else { L.c0 = o->stokes.B; L.c1 = 0.5 * o->stokes.eps * o->stokes.eps; L.c2 = 0.5 * (o->stokes.p - 2); L.c3 = o->stokes.nu / (1 - 2 * o->stokes.nu); }
  if (o->nelem) {
    const int err = q1fn[p - 1](L, c->stream);
    if (err) return seterr(700 + err, "Q1 assembly kernel launch failed (degree %d): %s", p, err > 0 ? cudaGetErrorString((cudaError_t)err) : "not instantiated");
    c->nlaunch++;
  }
  if (own_vals) CU(cudaMemcpyAsync(vals, dvals, nvals * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

// ---- Stokes "lumped mass" vector (StokesJacobianAssemble_Velocity, stokes.c:721-733) ----------------------------------------------
// Mentry_i = sum_q phi_i(q)^2 jw(q) of every Q1 patch is stored into the slot of quadrature point number i of the patch (the
// reference indexes the patch's quadrature-point array v with the basis-function index), the element's (2p)^3 slots go through
// INTERP_TRANSPOSE of the high-order basis on the SPARSE points (dRulesetIteratorCommitPatchApplied of the last patch,
// fsrulesetit.c:595-643) and are added into Mdiag (HOMOGENEOUS), the same value for the three components.  State independent.
__global__ void __launch_bounds__(128) q1_stokes_mdiag_kernel(int nelem, int p, int n_own, int nc, const int *e2l, const double *coords, const double *tab, double *md, double *mdaux)
{
  extern __shared__ double sm[];
  const int n = p + 1, Qs = 2 * p, tid = threadIdx.x, nth = blockDim.x;
  const long e = blockIdx.x;
  double *Bs = sm, *qs = Bs + 2 * Qs * n, *ws = qs + Qs, *X = ws + Qs, *vq = X + 24, *t1 = vq + Qs * Qs * Qs, *t2 = t1 + n * Qs * Qs;
  for (int i = tid; i < 2 * Qs * n + 2 * Qs; i += nth) sm[i] = tab[i];
  for (int i = tid; i < 24; i += nth) X[i] = coords[e * 24 + i];
  __syncthreads();
  constexpr double IA = 0.78867513459481287, IB = 0.21132486540518713;      // interp1, efstopo.c:665
  for (int pt = tid; pt < p * p * p; pt += nth) {
    const int pk = pt % p, pj = (pt / p) % p, pi = pt / (p * p);
    double jw[8];
    for (int q = 0; q < 8; q++) {
      const int qa = 2 * pi + (q >> 2), qb = 2 * pj + ((q >> 1) & 1), qc = 2 * pk + (q & 1);
      LineGeom lg;
      PointGeom pg;
      line_geom(X, qs[qa], qs[qb], lg, false);
      point_geom_adj(lg, qs[qc], pg);
      jw[q] = ws[qa] * ws[qb] * ws[qc] * pg.det;
    }
    for (int i = 0; i < 8; i++) {
      double m = 0;
      for (int q = 0; q < 8; q++) {
        const double phi = ((((q ^ i) >> 2) & 1) ? IB : IA) * ((((q ^ i) >> 1) & 1) ? IB : IA) * (((q ^ i) & 1) ? IB : IA);
        m += phi * jw[q] * phi;
      }
      vq[((2 * pi + (i >> 2)) * Qs + 2 * pj + ((i >> 1) & 1)) * Qs + 2 * pk + (i & 1)] = m;
    }
  }
  __syncthreads();
  for (int it = tid; it < n * Qs * Qs; it += nth) {         // x^T: t1[i][b][c] = sum_a B[a][i] vq[a][b][c]
    const int i = it / (Qs * Qs), bc = it - i * Qs * Qs;
    double s = 0;
    for (int a = 0; a < Qs; a++) s = fma(Bs[a * n + i], vq[a * Qs * Qs + bc], s);
    t1[it] = s;
  }
  __syncthreads();
  for (int it = tid; it < n * n * Qs; it += nth) {          // y^T: t2[i][j][c]
    const int i = it / (n * Qs), r = it - i * n * Qs, j = r / Qs, c = r - j * Qs;
    double s = 0;
    for (int b = 0; b < Qs; b++) s = fma(Bs[b * n + j], t1[(i * Qs + b) * Qs + c], s);
    t2[it] = s;
  }
  __syncthreads();
  for (int it = tid; it < n * n * n; it += nth) {           // z^T and dFSExpandedToGlobal(ADD), fs.c:592-640
    const int ij = it / n, k = it - ij * n;
    double s = 0;
    for (int c = 0; c < Qs; c++) s = fma(Bs[c * n + k], t2[ij * Qs + c], s);
    const int loc = e2l[e * n * n * n + it];
    double *dst = loc < n_own ? md + (long)loc * 3 : (loc >= nc ? mdaux + (long)(loc - n_own) * 3 : nullptr);      // Dirichlet rows dropped
    if (dst) { atomicAdd(dst, s); atomicAdd(dst + 1, s); atomicAdd(dst + 2, s); }
  }
}

extern "C" int dB200OpQ1SetReduced(dB200Op o, int flag) { o->q1reduced = flag ? 1 : 0; return 0; }

extern "C" int dB200OpStokesLumpedMass(dB200Op o, double *mdiag)
{
  dB200Ctx c = o->ctx;
  CU(cudaSetDevice(c->device));
  if (o->kind != dB200_OP_STOKES) return seterr(62, "dB200OpStokesLumpedMass: not a Stokes operator");
  dB200Space s = o->space[0];
  const int p = s->P - 1, n = p + 1, Qs = 2 * p;
  CHK(q1_tables(o));
  const size_t nb = (size_t)s->n * 3;
  double *dmd = mdiag, *own = nullptr;
  if (!is_dev(mdiag)) { CU(cudaMalloc((void **)&own, std::max<size_t>(nb, 1) * sizeof(double))); dmd = own; }
  CU(cudaMemsetAsync(dmd, 0, nb * sizeof(double), c->stream));                                   // VecZeroEntries(Mdiag), stokes.c:683
  if (s->ngh) CU(cudaMemsetAsync(s->d_yaux + (size_t)(s->nc - s->n) * 3, 0, (size_t)s->ngh * 3 * sizeof(double), c->stream));
  const size_t smem = sizeof(double) * ((size_t)2 * Qs * n + 2 * Qs + 24 + (size_t)Qs * Qs * Qs + (size_t)n * Qs * Qs + (size_t)n * n * Qs);
  if (smem > 227 * 1024) return seterr(56, "lumped mass: degree %d needs %zu B of shared memory", p, smem);
  if (o->nelem) {
    CU(cudaFuncSetAttribute(q1_stokes_mdiag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    q1_stokes_mdiag_kernel<<<o->nelem, 128, smem, c->stream>>>(o->nelem, p, s->n, s->nc, s->d_e2l, o->d_coords, o->d_q1tab, dmd, s->d_yaux);
    CU(cudaGetLastError());
    c->nlaunch++;
  }
  CHK(dohp_halo_reverse(s, s->d_yaux, dmd));
  if (own) { CU(cudaMemcpyAsync(mdiag, dmd, nb * sizeof(double), cudaMemcpyDeviceToHost, c->stream)); }
  CU(cudaStreamSynchronize(c->stream));
  if (own) cudaFree(own);
  return 0;
}
