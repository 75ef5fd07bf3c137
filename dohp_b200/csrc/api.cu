// C-ABI host layer of libdohp_b200.so (include/dohp_b200.h).  Host logic only: owns device buffers, stages host
// pointers through HBM, selects the kernel instantiation, enqueues on the context's stream.  There is no CPU compute
// path: every entry point that computes launches sm_100a kernels or fails.
#include "internal.h"
#include "launch.h"
#include "ellip_kernel.cuh"
#include "vec_kernel.cuh"
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <algorithm>

using namespace dohp;

// ---- errors -----------------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
namespace dohp {
int seterr(int code, const char *fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code ? code : 1;
}
}  // namespace dohp
extern "C" const char *dB200GetLastError(void) { return g_err; }

#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return seterr(700 + (int)e_, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); } while (0)
#define CHK(call) do { int e_ = (call); if (e_) return e_; } while (0)
#define LAUNCHED(ctx) ((ctx)->nlaunch++)

// ---- small kernels ---------------------------------------------------------------------------------------------------
namespace dohp {
__global__ void compute_physical_kernel(long npts, int Q0, int Q1, int Q2, const double *w0, const double *w1, const double *w2,
                                        const double *jac, double *jinv, double *jw, int *flag)
{
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= npts) return;
  const int q = (int)(g % ((long)Q0 * Q1 * Q2));
  const int k = q % Q2, j = (q / Q2) % Q1, i = q / (Q2 * Q1);
  const double *J = jac + g * 9;
  // dGeomInvert3, include/dohpgeom.h:57-75
  const double b0 = J[4] * J[8] - J[5] * J[7], b3 = J[5] * J[6] - J[3] * J[8], b6 = J[3] * J[7] - J[4] * J[6];
  const double det = J[0] * b0 + J[1] * b3 + J[2] * b6, idet = 1.0 / det;
  double *o = jinv + g * 9;
  o[0] = idet * b0;
  o[1] = idet * (J[2] * J[7] - J[1] * J[8]);
  o[2] = idet * (J[1] * J[5] - J[2] * J[4]);
  o[3] = idet * b3;
  o[4] = idet * (J[0] * J[8] - J[2] * J[6]);
  o[5] = idet * (J[2] * J[3] - J[0] * J[5]);
  o[6] = idet * b6;
  o[7] = idet * (J[1] * J[6] - J[0] * J[7]);
  o[8] = idet * (J[0] * J[4] - J[1] * J[3]);
  jw[g] = w0[i] * w1[j] * w2[k] * det;
  if (!(det > 0)) atomicOr(flag, 1);                      // jacobi.c:397 "Negative Jacobian"
}
}  // namespace dohp

// quadrature-point positions and det check for trilinear elements: one thread per point
__global__ void quad_points_kernel(long nelem, int NQ, const double *q, const double *coords, double *xq, int *flag)
{
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long nq3 = (long)NQ * NQ * NQ;
  if (g >= nelem * nq3) return;
  const long e = g / nq3;
  const int r = (int)(g - e * nq3), c = r % NQ, b = (r / NQ) % NQ, a = r / (NQ * NQ);
  LineGeom lg;
  line_geom(coords + e * 24, q[a], q[b], lg, true);
  double jinv[9], x[3];
  const double det = point_geom(lg, q[c], jinv);
  point_pos(lg, q[c], x);
  if (xq) { xq[g * 3 + 0] = x[0]; xq[g * 3 + 1] = x[1]; xq[g * 3 + 2] = x[2]; }
  if (!(det > 0)) atomicOr(flag, 1);
}
// Is every element a parallelepiped (affine map)?  The four mixed differences of the trilinear map vanish up to the rounding of the
// vertex coordinates.  Affine meshes take the kernels that evaluate the geometry once per element.
__global__ void affine_check_kernel(long nelem, const double *coords, int *flag)
{
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nelem) return;
  const double *X = coords + e * 24;
  bool affine = true;
  double scale = 0;
  for (int d = 0; d < 3; d++) scale += fabs(X[12 + d] - X[d]) + fabs(X[6 + d] - X[d]) + fabs(X[3 + d] - X[d]) + fabs(X[d]);
  for (int d = 0; d < 3; d++) {
    const double x000 = X[d], x001 = X[3 + d], x010 = X[6 + d], x011 = X[9 + d], x100 = X[12 + d], x101 = X[15 + d], x110 = X[18 + d], x111 = X[21 + d];
    const double m1 = x110 - x100 - x010 + x000, m2 = x101 - x100 - x001 + x000, m3 = x011 - x010 - x001 + x000;
    const double m4 = x111 - x110 - x101 - x011 + x100 + x010 + x001 - x000;
    affine = affine && fabs(m1) <= 4e-15 * scale && fabs(m2) <= 4e-15 * scale && fabs(m3) <= 4e-15 * scale && fabs(m4) <= 8e-15 * scale;
  }
  if (!affine) atomicOr(flag, 2);
}
// [e][q=(a,b,c)][nf] (reference order)  <->  planes [f][e][c][a*NQ+b] (kernel order)
__global__ void to_planes_kernel(long nelem, int NQ, int nf, const double *src, double *dst)
{
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long nq3 = (long)NQ * NQ * NQ;
  if (g >= nelem * nq3 * nf) return;
  const int f = (int)(g % nf);
  const long pq = g / nf, e = pq / nq3;
  const int r = (int)(pq - e * nq3), c = r % NQ, ab = r / NQ;
  dst[(f * nelem + e) * nq3 + (long)c * NQ * NQ + ab] = src[g];
}
__global__ void from_planes_kernel(long nelem, int NQ, int nf, const double *src, double *dst)
{
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long nq3 = (long)NQ * NQ * NQ;
  if (g >= nelem * nq3 * nf) return;
  const int f = (int)(g % nf);
  const long pq = g / nf, e = pq / nq3;
  const int r = (int)(pq - e * nq3), c = r % NQ, ab = r / NQ;
  dst[g] = src[(f * nelem + e) * nq3 + (long)c * NQ * NQ + ab];
}
__global__ void gather_idx_kernel(long n, int bs, const int *idx, const double *src, long srcoff, double *dst)
{
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n * bs) return;
  dst[g] = src[((long)idx[g / bs] - srcoff) * bs + g % bs];
}
__global__ void scatter_add_idx_kernel(long n, int bs, const int *idx, const double *src, double *dst)
{
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n * bs) return;
  atomicAdd(dst + (long)idx[g / bs] * bs + g % bs, src[g]);
}
__global__ void scatter_idx_kernel(long n, int bs, const int *idx, long dstoff, const double *src, double *dst)
{
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n * bs) return;
  dst[((long)idx[g / bs] - dstoff) * bs + g % bs] = src[g];
}
__global__ void axpy_kernel(long n, const double *x, double *y)
{
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n) y[g] += x[g];
}
// register-resident DFMA chains: 8 independent accumulators per thread
__global__ void dfma_peak_kernel(double *out, int iters, double a, double b)
{
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int r = 0; r < 16; r++) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  out[(long)blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

static inline int nblk(long n, int t) { return (int)((n + t - 1) / t); }

int dohp_smem_maps(dB200Ctx c, int nq, int sa, int sb, int epb, int threads, int es_base, int *es, const int **d_maps)
{
  const std::array<int, 6> key = {nq, sa, sb, epb, threads, es_base};
  auto it = c->smem_maps.find(key);
  if (it == c->smem_maps.end()) {
    std::vector<int> maps;
    int e;
    CHK(build_smem_maps(nq, sa, sb, epb, threads, es_base, &e, maps));
    int *d;
    CU(cudaMalloc((void **)&d, maps.size() * sizeof(int)));
    CU(cudaMemcpyAsync(d, maps.data(), maps.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    it = c->smem_maps.emplace(key, std::make_pair(e, d)).first;
  }
  *es = it->second.first;
  *d_maps = it->second.second;
  return 0;
}

// ---- context ---------------------------------------------------------------------------------------------------------------
static int ctx_init_common(int device, dB200Ctx c)
{
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return seterr(62, "no CUDA device: dohp_b200 has no CPU path"); }
  if (device < 0 || device >= ndev) return seterr(63, "device %d out of range (%d devices)", device, ndev);
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return seterr(62, "device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major, prop.minor);
  c->device = device;
  CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  c->own_stream = true;
  return 0;
}
extern "C" int dB200Init(int device, dB200Ctx *ctx)
{
  dB200Ctx c = new dB200Ctx_s;
  int err = ctx_init_common(device, c);
  if (err) { delete c; return err; }
  *ctx = c;
  return 0;
}
extern "C" int dB200Destroy(dB200Ctx *ctx)
{
  if (!ctx || !*ctx) return 0;
  dB200Ctx c = *ctx;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (int i = 0; i < 8; i++) if (c->scratch[i]) cudaFree(c->scratch[i]);
  for (auto &kv : c->smem_maps) cudaFree(kv.second.second);
  if (c->comm) { cudaStreamSynchronize(c->comm); cudaStreamDestroy(c->comm); for (int i = 0; i < 4; i++) cudaEventDestroy((cudaEvent_t)c->ev[i]); }
  if (c->h2d) { cudaStreamSynchronize(c->h2d); cudaStreamSynchronize(c->d2h); cudaStreamDestroy(c->h2d); cudaStreamDestroy(c->d2h); }
  if (c->own_stream) cudaStreamDestroy(c->stream);
  dohp_comm_destroy(c);
  delete c;
  *ctx = nullptr;
  return 0;
}
extern "C" int dB200SetStream(dB200Ctx c, void *s)
{
  if (c->own_stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); c->own_stream = false; }
  c->stream = (cudaStream_t)s;
  return 0;
}
// Device-side join: the context's stream waits for everything enqueued so far on the copy streams of dB200OpMultAsync (so that an
// event recorded on the context's stream afterwards covers the downloads too); does not block the host.
extern "C" int dB200StreamJoin(dB200Ctx c)
{
  if (!c->h2d) return 0;
  cudaEvent_t e;
  CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  CU(cudaEventRecord(e, c->h2d)); CU(cudaStreamWaitEvent(c->stream, e, 0));
  CU(cudaEventRecord(e, c->d2h)); CU(cudaStreamWaitEvent(c->stream, e, 0));
  CU(cudaEventDestroy(e));
  return 0;
}
extern "C" int dB200Synchronize(dB200Ctx c)
{
  if (c->h2d) CU(cudaStreamSynchronize(c->h2d));
  CU(cudaStreamSynchronize(c->stream));
  if (c->d2h) CU(cudaStreamSynchronize(c->d2h));
  return 0;
}
extern "C" int dB200NumLaunches(dB200Ctx c, long long *n) { *n = c->nlaunch; return 0; }

static bool is_device_ptr(const void *p)
{
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}
static int scratch_get(dB200Ctx c, int slot, size_t bytes, void **p)
{
  if (c->scratch_bytes[slot] < bytes) {
    if (c->scratch[slot]) { CU(cudaStreamSynchronize(c->stream)); CU(cudaFree(c->scratch[slot])); c->scratch[slot] = nullptr; c->scratch_bytes[slot] = 0; }
    CU(cudaMalloc(&c->scratch[slot], bytes));
    c->scratch_bytes[slot] = bytes;
  }
  *p = c->scratch[slot];
  return 0;
}
// device view of an input buffer
static int stage_in(dB200Ctx c, int slot, const void *p, size_t bytes, const void **dev)
{
  if (!p || is_device_ptr(p)) { *dev = p; return 0; }
  void *s;
  CHK(scratch_get(c, slot, bytes, &s));
  CU(cudaMemcpyAsync(s, p, bytes, cudaMemcpyHostToDevice, c->stream));
  *dev = s;
  return 0;
}
// device view of an output buffer (load=true copies the current host contents in, for ADD semantics)
static int stage_out(dB200Ctx c, int slot, void *p, size_t bytes, bool load, void **dev, bool *host)
{
  *host = false;
  if (!p || is_device_ptr(p)) { *dev = p; return 0; }
  void *s;
  CHK(scratch_get(c, slot, bytes, &s));
  if (load) CU(cudaMemcpyAsync(s, p, bytes, cudaMemcpyHostToDevice, c->stream));
  *dev = s;
  *host = true;
  return 0;
}
static int finish_out(dB200Ctx c, void *p, const void *dev, size_t bytes, bool host)
{
  if (!host) return 0;
  CU(cudaMemcpyAsync(p, dev, bytes, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}
template <class T> static int upload(dB200Ctx c, const T *p, size_t count, T **dev)
{
  CU(cudaMalloc((void **)dev, std::max<size_t>(count, 1) * sizeof(T)));
  if (count) CU(cudaMemcpyAsync(*dev, p, count * sizeof(T), is_device_ptr(p) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

// ---- rule / efs -----------------------------------------------------------------------------------------------------------
extern "C" int dB200RuleCreate(dB200Ctx ctx, const int Q[3], const double *const coord[3], const double *const weight[3], dB200Rule *rule)
{
  dB200Rule r = new dB200Rule_s;
  r->ctx = ctx;
  for (int d = 0; d < 3; d++) {
    if (Q[d] < 1 || Q[d] > 64) { delete r; return seterr(63, "rule size Q[%d]=%d out of range", d, Q[d]); }
    r->Q[d] = Q[d];
    r->coord[d].assign(coord[d], coord[d] + Q[d]);
    r->weight[d].assign(weight[d], weight[d] + Q[d]);
    r->Dt[d].assign((size_t)Q[d] * Q[d], 0.0);
    if (Q[d] > 1) collocation_derivative(Q[d], coord[d], r->Dt[d].data());
  }
  *rule = r;
  return 0;
}
// dTOPO_LINE / dTOPO_QUAD rules (dQuadratureGetRules_Tensor_Private, tensorquad.c:186-195): the same 1-D tables in 1 or 2 directions
extern "C" int dB200RuleCreateDim(dB200Ctx ctx, int dim, const int Q[], const double *const coord[], const double *const weight[], dB200Rule *rule)
{
  if (dim < 1 || dim > 3) return seterr(63, "topological dimension %d", dim);
  static const double zero = 0, one = 1;
  int Q3[3] = {1, 1, 1};
  const double *c3[3] = {&zero, &zero, &zero}, *w3[3] = {&one, &one, &one};
  for (int d = 0; d < dim; d++) { Q3[d] = Q[d]; c3[d] = coord[d]; w3[d] = weight[d]; }
  CHK(dB200RuleCreate(ctx, Q3, c3, w3, rule));
  (*rule)->dim = dim;
  return 0;
}
extern "C" int dB200RuleDestroy(dB200Rule *rule) { if (rule && *rule) { delete *rule; *rule = nullptr; } return 0; }

extern "C" int dB200EFSCreate(dB200Ctx ctx, dB200Rule rule, const int P[3], const double *const interp[3], const double *const deriv[3], dB200EFS *efs)
{
  dB200EFS e = new dB200EFS_s;
  e->ctx = ctx;
  e->rule = rule;
  for (int d = 0; d < 3; d++) {
    const int Q = rule->Q[d];
    if (P[d] < 1 || P[d] > 14) { delete e; return seterr(63, "basis size P[%d]=%d out of range", d, P[d]); }
    e->P[d] = P[d];
    e->interp[d].assign(interp[d], interp[d] + (size_t)Q * P[d]);
    e->deriv[d].assign(deriv[d], deriv[d] + (size_t)Q * P[d]);
    // the kernels differentiate by collocation on the rule points: exact iff deriv == Dt . interp
    double dmax = 0, emax = 0;
    for (int q = 0; q < Q; q++)
      for (int j = 0; j < P[d]; j++) {
        long double s = 0;
        for (int m = 0; m < Q; m++) s += (long double)rule->Dt[d][q * Q + m] * interp[d][m * P[d] + j];
        dmax = std::max(dmax, fabs(deriv[d][q * P[d] + j]));
        emax = std::max(emax, fabs((double)s - deriv[d][q * P[d] + j]));
      }
    // (P > Q, or tables that are not a polynomial basis): only the explicit-table path (apply_generic.cu) applies such a dEFS
    if (emax > 1e-12 * std::max(dmax, 1.0)) e->colloc_ok = false;
  }
  *efs = e;
  return 0;
}
// dEFS on dTOPO_LINE / dTOPO_QUAD (dJacobiGetEFS_Tensor, tensor.c:522-531): 1 or 2 TensorBasis
extern "C" int dB200EFSCreateDim(dB200Ctx ctx, dB200Rule rule, int dim, const int P[], const double *const interp[], const double *const deriv[], dB200EFS *efs)
{
  if (dim < 1 || dim > 3 || dim != rule->dim) return seterr(63, "dEFS dimension %d on a rule of dimension %d", dim, rule->dim);
  static const double zero = 0, one = 1;
  int P3[3] = {1, 1, 1};
  const double *i3[3] = {&one, &one, &one}, *d3[3] = {&zero, &zero, &zero};
  for (int d = 0; d < dim; d++) { P3[d] = P[d]; i3[d] = interp[d]; d3[d] = deriv[d]; }
  CHK(dB200EFSCreate(ctx, rule, P3, i3, d3, efs));
  (*efs)->dim = dim;
  return 0;
}
extern "C" int dB200EFSDestroy(dB200EFS *efs) { if (efs && *efs) { if ((*efs)->d_gtab) cudaFree((*efs)->d_gtab); delete *efs; *efs = nullptr; } return 0; }
extern "C" int dB200EFSGetSizes(dB200EFS efs, int P[3], int Q[3])
{
  for (int d = 0; d < 3; d++) { if (P) P[d] = efs->P[d]; if (Q) Q[d] = efs->rule->Q[d]; }
  return 0;
}

static bool same(const std::vector<double> &a, const std::vector<double> &b)
{
  if (a.size() != b.size()) return false;
  for (size_t i = 0; i < a.size(); i++) if (fabs(a[i] - b[i]) > 1e-14 * (1 + fabs(a[i]))) return false;
  return true;
}
static int efs_tables(dB200EFS e, Tables *t)
{
  dB200Rule r = e->rule;
  if (!(e->P[0] == e->P[1] && e->P[1] == e->P[2] && r->Q[0] == r->Q[1] && r->Q[1] == r->Q[2] && same(e->interp[0], e->interp[1]) &&
        same(e->interp[1], e->interp[2]) && same(r->coord[0], r->coord[1]) && same(r->coord[1], r->coord[2]) && same(r->weight[0], r->weight[1]) &&
        same(r->weight[1], r->weight[2])))
    return seterr(56, "anisotropic dEFS (P=%d,%d,%d Q=%d,%d,%d): the fused operator kernels are instantiated for isotropic hexes only (dB200EFSApplyBatch applies any dEFS)", e->P[0], e->P[1], e->P[2], r->Q[0], r->Q[1], r->Q[2]);
  if (!e->colloc_ok) return seterr(56, "dEFS (P=%d, Q=%d) whose derivative table is not the collocation derivative of its interpolation table: fused operator kernels do not apply", e->P[0], r->Q[0]);
  t->NP = e->P[0]; t->NQ = r->Q[0];
  t->B = e->interp[0].data(); t->Dt = r->Dt[0].data(); t->w = r->weight[0].data(); t->q = r->coord[0].data();
  return 0;
}

static const ApplyFn apply_fns[] = {nullptr, nullptr, launch_apply_2, launch_apply_3, launch_apply_4, launch_apply_5, launch_apply_6,
                                    launch_apply_7, launch_apply_8, launch_apply_9, launch_apply_10, launch_apply_11};
static const VecFn vec_fns[] = {nullptr, nullptr, launch_vec_2, launch_vec_3, launch_vec_4, launch_vec_5, launch_vec_6,
                                launch_vec_7, launch_vec_8, launch_vec_9, launch_vec_10, launch_vec_11};
static const EllipFn ellip_fns[] = {nullptr, nullptr, launch_ellip_2, launch_ellip_3, launch_ellip_4, launch_ellip_5, launch_ellip_6,
                                    launch_ellip_7, launch_ellip_8, launch_ellip_9, launch_ellip_10, launch_ellip_11};

extern "C" int dB200EFSApplyBatch(dB200EFS efs, int nelem, const double *jinv, int D, const double *in, double *out, int amode, int imode)
{
  dB200Ctx c = efs->ctx;
  CU(cudaSetDevice(c->device));
  if (D < 1) return seterr(63, "dofs=%d must be at least 1", D);                           // jacobi.c:495
  if (amode < dB200_APPLY_INTERP || amode > dB200_APPLY_GRAD_TRANSPOSE) return seterr(56, "apply mode %d not implemented", amode);   // efstopo.c:419-420
  if (imode != dB200_INSERT_VALUES && imode != dB200_ADD_VALUES) return seterr(56, "InsertMode %d", imode);
  if (amode == dB200_APPLY_GRAD_TRANSPOSE && !jinv) return seterr(56, "GRAD_TRANSPOSE needs jinv");   // efstopo.c:600
  const int dim = efs->dim;
  if (dim < 3) {                                          // dEFSApply_Tensor_Line / _Quad, efstopo.c:290-366
    if (amode == dB200_APPLY_GRAD_TRANSPOSE) return seterr(56, "invalid dApplyMode %d on a %d-dimensional dEFS", amode, dim);     // efstopo.c:321-325, 361-363
    if (jinv) return seterr(56, "mapped gradients are not implemented for Line / Quad (as upstream, efstopo.c:638,653)");
  }
  ApplyLaunch L;
  // isotropic dEFS with an instantiated (P,Q) pair: templated kernel; anything else dJacobiGetEFS can produce: explicit-table kernel
  bool generic = dim < 3 || efs_tables(efs, &L.tab) != 0;
  if (!generic) generic = L.tab.NQ < DOHP_NMIN || L.tab.NQ > DOHP_NMAX || !(L.tab.NP == L.tab.NQ || (L.tab.NQ >= 3 && L.tab.NP == L.tab.NQ - 1) || (L.tab.NQ >= 4 && L.tab.NP == 2));
  const size_t np3 = (size_t)efs->P[0] * efs->P[1] * efs->P[2], nq3 = (size_t)efs->rule->Q[0] * efs->rule->Q[1] * efs->rule->Q[2];
  const bool fwd = amode == dB200_APPLY_INTERP || amode == dB200_APPLY_GRAD;
  const bool grad = amode == dB200_APPLY_GRAD || amode == dB200_APPLY_GRAD_TRANSPOSE;
  const size_t nin = (size_t)nelem * D * (fwd ? np3 : nq3 * (grad ? 3 : 1)), nout = (size_t)nelem * D * (fwd ? nq3 * (grad ? 3 : 1) : np3);
  const void *din, *djinv;
  void *dout;
  bool host;
  CHK(stage_in(c, 0, in, nin * sizeof(double), &din));
  CHK(stage_in(c, 1, jinv, (size_t)nelem * nq3 * 9 * sizeof(double), &djinv));
  CHK(stage_out(c, 2, out, nout * sizeof(double), imode == dB200_ADD_VALUES, &dout, &host));
  if (nelem == 0) return 0;
  L.nelem = nelem; L.D = D; L.amode = amode; L.add = imode == dB200_ADD_VALUES;
  L.jinv = (const double *)djinv; L.in = (const double *)din; L.out = (double *)dout;
  int err = generic ? launch_apply_generic(efs, nelem, D, amode, L.add, L.jinv, L.in, L.out, c->stream) : apply_fns[L.tab.NQ](L, c->stream);
  if (err == -1) return seterr(56, "no kernel instantiated for (P,Q)=(%d,%d)", L.tab.NP, L.tab.NQ);
  if (err == -2) return seterr(56, "dEFS too large for the shared-memory staging of the generic apply (P=%d,%d,%d Q=%d,%d,%d)", efs->P[0], efs->P[1], efs->P[2], efs->rule->Q[0], efs->rule->Q[1], efs->rule->Q[2]);
  if (err) return seterr(700 + err, "apply kernel launch: %s", cudaGetErrorString((cudaError_t)err));
  LAUNCHED(c);
  return finish_out(c, out, dout, nout * sizeof(double), host);
}

// dQuadratureGetFacetRules_Tensor_Private (tensorquad.c:309-355): the rule of hex facet `facet` is the QUAD rule moved into the hex's
// reference cube by x = jac(:, 0:2) X_ref + translation (the six rotations of tensorquad.c:321-328); weights are the quad's
// (dRuleGetNodeWeight_Transformed, tensorquad.c:225-243).  Host arithmetic, setup only.
extern "C" int dB200FacetRuleNodes(int facet, const int Q[2], const double *const coord[2], const double *const weight[2], double *nodes, double *weights)
{
  static const double jac[6][3][3] = {{{1, 0, 0}, {0, 0, -1}, {0, 1, 0}}, {{0, 0, 1}, {1, 0, 0}, {0, 1, 0}}, {{-1, 0, 0}, {0, 0, 1}, {0, 1, 0}},
                                      {{0, 0, -1}, {-1, 0, 0}, {0, 1, 0}}, {{0, 1, 0}, {1, 0, 0}, {0, 0, -1}}, {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}};
  static const double tr[6][3] = {{0, -1, 0}, {1, 0, 0}, {0, 1, 0}, {-1, 0, 0}, {0, 0, -1}, {0, 0, 1}};
  if (facet < 0 || facet > 5) return seterr(63, "hex facet %d", facet);
  for (int i = 0; i < Q[0]; i++)
    for (int j = 0; j < Q[1]; j++) {
      const double X[3] = {coord[0][i], coord[1][j], 0.0};
      double *x = nodes + 3 * (i * Q[1] + j);
      for (int d = 0; d < 3; d++) x[d] = jac[facet][d][0] * X[0] + jac[facet][d][1] * X[1] + jac[facet][d][2] * X[2] + tr[facet][d];
      if (weights) weights[i * Q[1] + j] = weight[0][i] * weight[1][j];
    }
  return 0;
}

extern "C" int dB200RuleComputePhysicalBatch(dB200Rule rule, int nelem, const double *jac, double *jinv, double *jw)
{
  dB200Ctx c = rule->ctx;
  CU(cudaSetDevice(c->device));
  const long nq = (long)rule->Q[0] * rule->Q[1] * rule->Q[2], npts = nq * nelem;
  const void *djac;
  void *djinv, *djw, *dw;
  bool h1, h2;
  CHK(stage_in(c, 0, jac, npts * 9 * sizeof(double), &djac));
  CHK(stage_out(c, 1, jinv, npts * 9 * sizeof(double), false, &djinv, &h1));
  CHK(stage_out(c, 2, jw, npts * sizeof(double), false, &djw, &h2));
  const int nw = rule->Q[0] + rule->Q[1] + rule->Q[2];
  CHK(scratch_get(c, 3, (nw + 1) * sizeof(double), &dw));
  std::vector<double> w;
  for (int d = 0; d < 3; d++) w.insert(w.end(), rule->weight[d].begin(), rule->weight[d].end());
  CU(cudaMemcpyAsync(dw, w.data(), nw * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  int *flag = (int *)((double *)dw + nw);
  CU(cudaMemsetAsync(flag, 0, sizeof(int), c->stream));
  const double *w0 = (double *)dw, *w1 = w0 + rule->Q[0], *w2 = w1 + rule->Q[1];
  compute_physical_kernel<<<nblk(npts, 256), 256, 0, c->stream>>>(npts, rule->Q[0], rule->Q[1], rule->Q[2], w0, w1, w2, (const double *)djac,
                                                                 (double *)djinv, (double *)djw, flag);
  CU(cudaGetLastError());
  LAUNCHED(c);
  int hflag = 0;
  CU(cudaMemcpyAsync(&hflag, flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  CHK(finish_out(c, jinv, djinv, npts * 9 * sizeof(double), h1));
  CHK(finish_out(c, jw, djw, npts * sizeof(double), h2));
  if (hflag) return seterr(1, "Negative Jacobian");
  return 0;
}

// ---- spaces ------------------------------------------------------------------------------------------------------------------
extern "C" int dB200SpaceCreate(dB200Ctx ctx, int nelem, int P, int bs, int n, int nc, int ngh, const int *e2l, dB200Space *space)
{
  CU(cudaSetDevice(ctx->device));
  if (nelem < 0 || P < 2 || P > 14 || bs < 1 || n < 0 || nc < n || ngh < 0) return seterr(63, "bad space sizes");
  dB200Space s = new dB200Space_s;
  s->ctx = ctx; s->nelem = nelem; s->P = P; s->bs = bs; s->n = n; s->nc = nc; s->ngh = ngh;
  int err = upload(ctx, e2l, (size_t)nelem * P * P * P, &s->d_e2l);
  if (err) { delete s; return err; }
  const size_t naux = (size_t)(nc + ngh - n) * bs;
  CU(cudaMalloc((void **)&s->d_xaux, std::max<size_t>(naux, 1) * sizeof(double)));
  CU(cudaMalloc((void **)&s->d_yaux, std::max<size_t>(naux, 1) * sizeof(double)));
  CU(cudaMemsetAsync(s->d_xaux, 0, std::max<size_t>(naux, 1) * sizeof(double), ctx->stream));
  CU(cudaMemsetAsync(s->d_yaux, 0, std::max<size_t>(naux, 1) * sizeof(double), ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  *space = s;
  return 0;
}
extern "C" int dB200SpaceDestroy(dB200Space *space)
{
  if (!space || !*space) return 0;
  dB200Space s = *space;
  cudaSetDevice(s->ctx->device);
  dohp_p2p_free(s);
  cudaFree(s->d_e2l); cudaFree(s->d_xaux); cudaFree(s->d_yaux);
  if (s->d_xaux_inhomog) cudaFree(s->d_xaux_inhomog);
  if (s->d_sendidx) cudaFree(s->d_sendidx);
  if (s->d_recvidx) cudaFree(s->d_recvidx);
  if (s->d_sendbuf) cudaFree(s->d_sendbuf);
  if (s->d_recvbuf) cudaFree(s->d_recvbuf);
  delete s;
  *space = nullptr;
  return 0;
}
extern "C" int dB200SpaceSetHalo(dB200Space s, int nneigh, const int *ranks, const int *sendcount, const int *const *sendidx,
                                 const int *recvcount, const int *const *recvidx)
{
  dB200Ctx c = s->ctx;
  CU(cudaSetDevice(c->device));
  if (s->d_sendidx || s->d_sendbuf) {                      // called again: drop the previous lists and buffers
    CU(cudaStreamSynchronize(c->stream));
    dohp_p2p_free(s);
    cudaFree(s->d_sendidx); cudaFree(s->d_recvidx); cudaFree(s->d_sendbuf); cudaFree(s->d_recvbuf);
    s->d_sendidx = s->d_recvidx = nullptr; s->d_sendbuf = s->d_recvbuf = nullptr;
  }
  s->nneigh = nneigh;
  s->ranks.assign(ranks, ranks + nneigh);
  s->sendcount.assign(sendcount, sendcount + nneigh);
  s->recvcount.assign(recvcount, recvcount + nneigh);
  s->sendoff.assign(nneigh + 1, 0);
  s->recvoff.assign(nneigh + 1, 0);
  std::vector<int> si, ri;
  for (int r = 0; r < nneigh; r++) {
    s->sendoff[r + 1] = s->sendoff[r] + sendcount[r];
    s->recvoff[r + 1] = s->recvoff[r] + recvcount[r];
    for (int i = 0; i < sendcount[r]; i++) { if (sendidx[r][i] < 0 || sendidx[r][i] >= s->n) return seterr(63, "halo send index %d not an owned node", sendidx[r][i]); si.push_back(sendidx[r][i]); }
    for (int i = 0; i < recvcount[r]; i++) { if (recvidx[r][i] < s->nc || recvidx[r][i] >= s->nc + s->ngh) return seterr(63, "halo recv index %d not a ghost node", recvidx[r][i]); ri.push_back(recvidx[r][i]); }
  }
  s->nsend = (int)si.size(); s->nrecv = (int)ri.size();
  CHK(upload(c, si.data(), si.size(), &s->d_sendidx));
  CHK(upload(c, ri.data(), ri.size(), &s->d_recvidx));
  CU(cudaMalloc((void **)&s->d_sendbuf, std::max<size_t>(std::max(si.size(), ri.size()) * s->bs, 1) * sizeof(double)));
  CU(cudaMalloc((void **)&s->d_recvbuf, std::max<size_t>(std::max(si.size(), ri.size()) * s->bs, 1) * sizeof(double)));
  CHK(dohp_p2p_setup(s));                                  // collective when the context has a communicator; leaves the NCCL path if peers cannot be mapped
  return 0;
}


extern "C" int dB200SpaceHaloMode(dB200Space s, int *mode)
{
  *mode = (s->ngh == 0 && s->nsend == 0) ? 0 : (s->p2p ? 2 : 1);
  return 0;
}

// forward INSERT (fs.c:578-579): owners' x values -> the ghost part of aux (aux index = local - n)
static int halo_forward(dB200Space s, const double *x, double *aux, cudaStream_t st = nullptr)
{
  dB200Ctx c = s->ctx;
  if (!st) st = c->stream;
  if (s->ngh == 0 && s->nsend == 0) return 0;
  if (c->nranks == 1) return seterr(73, "space has ghosts but the context has no communicator (use dB200InitComm)");
  if (s->p2p && (aux == s->d_xaux || aux == s->d_xaux_inhomog)) return dohp_p2p_forward(s, x, aux == s->d_xaux_inhomog, st);      // NVLink peer memory
  if (s->nsend) { gather_idx_kernel<<<nblk((long)s->nsend * s->bs, 256), 256, 0, st>>>(s->nsend, s->bs, s->d_sendidx, x, 0, s->d_sendbuf); LAUNCHED(c); }
  CHK(dohp_exchange(c, s, false, st));
  if (s->nrecv) { scatter_idx_kernel<<<nblk((long)s->nrecv * s->bs, 256), 256, 0, st>>>(s->nrecv, s->bs, s->d_recvidx, s->n, s->d_recvbuf, aux); LAUNCHED(c); }
  CU(cudaGetLastError());
  return 0;
}
// reverse ADD (fs.c:627-628): ghost contributions -> owners' y
static int halo_reverse(dB200Space s, const double *aux, double *y, cudaStream_t st = nullptr)
{
  dB200Ctx c = s->ctx;
  if (!st) st = c->stream;
  if (s->ngh == 0 && s->nsend == 0) return 0;
  if (c->nranks == 1) return seterr(73, "space has ghosts but the context has no communicator (use dB200InitComm)");
  if (s->p2p) return dohp_p2p_reverse(s, aux, y, st);
  if (s->nrecv) { gather_idx_kernel<<<nblk((long)s->nrecv * s->bs, 256), 256, 0, st>>>(s->nrecv, s->bs, s->d_recvidx, aux, s->n, s->d_recvbuf); LAUNCHED(c); }
  CHK(dohp_exchange(c, s, true, st));
  if (s->nsend) { scatter_add_idx_kernel<<<nblk((long)s->nsend * s->bs, 256), 256, 0, st>>>(s->nsend, s->bs, s->d_sendidx, s->d_sendbuf, y); LAUNCHED(c); }
  CU(cudaGetLastError());
  return 0;
}

// ---- overlap of the ghost exchange with the interior elements ---------------------------------------------------------
// A partitioned operator splits its elements into those that touch a ghost node (they read the values the forward exchange brings
// and produce the contributions the reverse exchange ships) and the interior ones.  Per traversal:
//   comm stream:  forward exchange                      | reverse exchange (after the interface elements)
//   main stream:  first half of the interior elements   | interface elements (after the forward exchange) | second half of the interior
// so both VecGhostUpdate legs (fs.c:578-579, 627-628) hide behind element work.
static int ctx_overlap_init(dB200Ctx c)
{
  if (c->comm) return 0;
  CU(cudaStreamCreateWithFlags(&c->comm, cudaStreamNonBlocking));
  for (int i = 0; i < 4; i++) CU(cudaEventCreateWithFlags((cudaEvent_t *)&c->ev[i], cudaEventDisableTiming));
  return 0;
}
__global__ void ghost_touch_kernel(long nelem, int P3, int nc, const int *e2l, int *flag)
{
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nelem * P3) return;
  if (e2l[g] >= nc) flag[g / P3] = 1;
}
static int op_build_eorder(dB200Op o)
{
  dB200Ctx c = o->ctx;
  bool any = false;
  for (int i = 0; i < o->nspaces; i++) any = any || o->space[i]->ngh > 0;
  if (!any || !o->nelem) return 0;
  int *dflag;
  CU(cudaMalloc((void **)&dflag, (size_t)o->nelem * sizeof(int)));
  CU(cudaMemsetAsync(dflag, 0, (size_t)o->nelem * sizeof(int), c->stream));
  for (int i = 0; i < o->nspaces; i++) {
    dB200Space s = o->space[i];
    const int P3 = s->P * s->P * s->P;
    if (s->ngh) { ghost_touch_kernel<<<nblk((long)o->nelem * P3, 256), 256, 0, c->stream>>>(o->nelem, P3, s->nc, s->d_e2l, dflag); LAUNCHED(c); }
  }
  std::vector<int> flag(o->nelem), order;
  CU(cudaMemcpyAsync(flag.data(), dflag, (size_t)o->nelem * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  cudaFree(dflag);
  order.reserve(o->nelem);
  for (int e = 0; e < o->nelem; e++) if (flag[e]) order.push_back(e);
  o->n_bnd = (int)order.size();
  for (int e = 0; e < o->nelem; e++) if (!flag[e]) order.push_back(e);
  CHK(upload(c, order.data(), order.size(), &o->d_eorder));
  return 0;
}
// Runs `launch(elist, count)` over all elements of the operator with the exchanges of the listed spaces overlapped; fwd[i] / rev[i]
// give the vectors of space i taking part (NULL: that leg is not needed for this space).
template <class LaunchFn>
static int run_overlapped(dB200Op o, int nsp, dB200Space *sps, const double *const *fwd_x, double *const *fwd_aux, const double *const *rev_aux, double *const *rev_y, LaunchFn launch)
{
  dB200Ctx c = o->ctx;
  bool comm = false;
  for (int i = 0; i < nsp; i++) comm = comm || ((fwd_x[i] || rev_y[i]) && (sps[i]->ngh || sps[i]->nsend));
  // measured on NVSwitch boxes (profiles/r2y_*): the split into three launches costs what it hides at every degree, at N = 2 and N = 8, so
  // the traversal is one launch between the two exchanges unless DOHP_B200_OVERLAP=1 asks for the overlapped schedule (slower links)
  static const bool no_overlap = !(getenv("DOHP_B200_OVERLAP") && atoi(getenv("DOHP_B200_OVERLAP")) != 0);
  if (!comm || !o->d_eorder || no_overlap) {
    for (int i = 0; i < nsp; i++) if (fwd_x[i]) CHK(halo_forward(sps[i], fwd_x[i], fwd_aux[i]));
    CHK(launch(nullptr, o->nelem, c->stream));
    for (int i = 0; i < nsp; i++) if (rev_y[i]) CHK(halo_reverse(sps[i], rev_aux[i], rev_y[i]));
    return 0;
  }
  CHK(ctx_overlap_init(c));
  cudaEvent_t *ev = (cudaEvent_t *)c->ev;
  const int n_int = o->nelem - o->n_bnd, h1 = n_int / 2;
  CU(cudaEventRecord(ev[0], c->stream));
  CU(cudaStreamWaitEvent(c->comm, ev[0], 0));
  for (int i = 0; i < nsp; i++) if (fwd_x[i]) CHK(halo_forward(sps[i], fwd_x[i], fwd_aux[i], c->comm));
  CU(cudaEventRecord(ev[1], c->comm));
  CHK(launch(o->d_eorder + o->n_bnd, h1, c->stream));
  CU(cudaStreamWaitEvent(c->stream, ev[1], 0));
  CHK(launch(o->d_eorder, o->n_bnd, c->stream));
  CU(cudaEventRecord(ev[2], c->stream));
  CU(cudaStreamWaitEvent(c->comm, ev[2], 0));
  for (int i = 0; i < nsp; i++) if (rev_y[i]) CHK(halo_reverse(sps[i], rev_aux[i], rev_y[i], c->comm));
  CU(cudaEventRecord(ev[3], c->comm));
  CHK(launch(o->d_eorder + o->n_bnd + h1, n_int - h1, c->stream));
  CU(cudaStreamWaitEvent(c->stream, ev[3], 0));
  return 0;
}

int dohp_halo_forward(dB200Space s, const double *x, double *aux) { return halo_forward(s, x, aux); }
int dohp_halo_reverse(dB200Space s, const double *aux, double *y) { return halo_reverse(s, aux, y); }

extern "C" int dB200SpaceGlobalToLocal(dB200Space s, const double *g, const double *dirichlet, int hmode, double *local)
{
  dB200Ctx c = s->ctx;
  CU(cudaSetDevice(c->device));
  const size_t bs = s->bs, nb = (size_t)s->n * bs, ndir = (size_t)(s->nc - s->n) * bs, nloc = (size_t)(s->nc + s->ngh) * bs;
  const void *dg, *dd;
  void *dl;
  bool host;
  CHK(stage_in(c, 0, g, nb * sizeof(double), &dg));
  CHK(stage_in(c, 1, hmode == dB200_INHOMOGENEOUS ? dirichlet : nullptr, ndir * sizeof(double), &dd));
  CHK(stage_out(c, 2, local, nloc * sizeof(double), false, &dl, &host));
  double *L = (double *)dl;
  CU(cudaMemcpyAsync(L, dg, nb * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  if (hmode == dB200_INHOMOGENEOUS && dd) CU(cudaMemcpyAsync(L + nb, dd, ndir * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  else CU(cudaMemsetAsync(L + nb, 0, ndir * sizeof(double), c->stream));                          // fs.c:562-571
  CU(cudaMemsetAsync(L + (size_t)s->nc * bs, 0, (size_t)s->ngh * bs * sizeof(double), c->stream));
  CHK(halo_forward(s, L, L + nb));
  return finish_out(c, local, dl, nloc * sizeof(double), host);
}
extern "C" int dB200SpaceLocalToGlobal(dB200Space s, const double *local, double *g, int imode)
{
  dB200Ctx c = s->ctx;
  CU(cudaSetDevice(c->device));
  const size_t bs = s->bs, nb = (size_t)s->n * bs, nloc = (size_t)(s->nc + s->ngh) * bs;
  const void *dl;
  void *dg;
  bool host;
  CHK(stage_in(c, 0, local, nloc * sizeof(double), &dl));
  CHK(stage_out(c, 2, g, nb * sizeof(double), imode == dB200_ADD_VALUES, &dg, &host));
  if (imode == dB200_ADD_VALUES) { axpy_kernel<<<nblk(nb, 256), 256, 0, c->stream>>>(nb, (const double *)dl, (double *)dg); LAUNCHED(c); }
  else CU(cudaMemcpyAsync(dg, dl, nb * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  CHK(halo_reverse(s, (const double *)dl + nb, (double *)dg));
  return finish_out(c, g, dg, nb * sizeof(double), host);
}

// ---- dFSRotation (slip boundaries) -------------------------------------------------------------------------------------------------
// One thread per rotated node: v <- R v (+ strong dofs) or v <- R^T (v with strong dofs set), the semantics the header of
// src/fs/interface/fsrot.c:4-37 documents.
__global__ void rotation_kernel(int n, int bs, const int *idx, const double *rmat, const int *ns, const int *soff, const double *strong, double *v, int reverse, int inhomog)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double *x = v + (long)idx[i] * bs;
  const double *R = rmat + (long)i * bs * bs, *sv = strong ? strong + soff[i] : nullptr;
  double tmp[16], out[16];
  for (int j = 0; j < bs; j++) tmp[j] = x[j];
  if (!reverse) {
    for (int k = 0; k < bs; k++) { double a = 0; for (int j = 0; j < bs; j++) a = fma(R[k * bs + j], tmp[j], a); out[k] = a; }
    for (int j = 0; j < ns[i]; j++) out[j] = inhomog ? sv[j] : 0.0;
  } else {
    for (int j = 0; j < ns[i]; j++) tmp[j] = inhomog ? sv[j] : 0.0;
    for (int k = 0; k < bs; k++) { double a = 0; for (int j = 0; j < bs; j++) a = fma(R[j * bs + k], tmp[j], a); out[k] = a; }
  }
  for (int j = 0; j < bs; j++) x[j] = out[j];
}
extern "C" int dB200SpaceRotationApply(dB200Space s, int n, const int *idx, const double *rmat, const int *ns, const double *strong, double *local, int reverse, int hmode)
{
  dB200Ctx c = s->ctx;
  CU(cudaSetDevice(c->device));
  const int bs = s->bs;
  if (bs > 16) return seterr(63, "large block size");                                               // fsrot.c:147
  if (n == 0) return 0;
  if (hmode != dB200_HOMOGENEOUS && hmode != dB200_INHOMOGENEOUS) return seterr(63, "Invalid homogeneous mode");
  if (hmode == dB200_INHOMOGENEOUS && !strong) return seterr(85, "INHOMOGENEOUS rotation needs the strong values");
  std::vector<int> hidx(n), hns(n), soff(n + 1, 0);
  CU(cudaMemcpy(hidx.data(), idx, n * sizeof(int), cudaMemcpyDefault));
  CU(cudaMemcpy(hns.data(), ns, n * sizeof(int), cudaMemcpyDefault));
  for (int i = 0; i < n; i++) {
    if (hns[i] < 0 || hns[i] > bs) return seterr(63, "Number of strong dofs must be between 0 and bs=%d (inclusive)", bs);     // fsrot.c:62
    if (hidx[i] < 0 || hidx[i] >= s->nc + s->ngh) return seterr(63, "rotated node %d outside the local vector", hidx[i]);
    soff[i + 1] = soff[i] + hns[i];
  }
  int *d_idx = nullptr, *d_ns = nullptr, *d_soff = nullptr;
  double *d_r = nullptr, *d_s = nullptr;
  struct Tmp { std::vector<void *> p; ~Tmp() { for (void *q : p) cudaFree(q); } } tmp;
  CHK(upload(c, hidx.data(), (size_t)n, &d_idx)); tmp.p.push_back(d_idx);
  CHK(upload(c, hns.data(), (size_t)n, &d_ns)); tmp.p.push_back(d_ns);
  CHK(upload(c, soff.data(), (size_t)n + 1, &d_soff)); tmp.p.push_back(d_soff);
  CHK(upload(c, rmat, (size_t)n * bs * bs, &d_r)); tmp.p.push_back(d_r);
  if (hmode == dB200_INHOMOGENEOUS) { CHK(upload(c, strong, (size_t)soff[n], &d_s)); tmp.p.push_back(d_s); }
  const size_t nloc = (size_t)(s->nc + s->ngh) * bs;
  void *dl;
  bool host;
  CHK(stage_out(c, 2, local, nloc * sizeof(double), true, &dl, &host));
  rotation_kernel<<<nblk(n, 128), 128, 0, c->stream>>>(n, bs, d_idx, d_r, d_ns, d_soff, d_s, (double *)dl, reverse, hmode == dB200_INHOMOGENEOUS);
  LAUNCHED(c);
  CU(cudaGetLastError());
  CHK(finish_out(c, local, dl, nloc * sizeof(double), host));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

// ---- operators ------------------------------------------------------------------------------------------------------------------
extern "C" int dB200OpCreate(dB200Ctx ctx, int kind, const void *param, dB200Rule rule, int nspaces, const dB200Space *spaces,
                             const dB200EFS *efs, const double *coords, dB200Op *op)
{
  CU(cudaSetDevice(ctx->device));
  if (nspaces < 1 || nspaces > DB200_MAXFIELDS) return seterr(63, "nspaces=%d", nspaces);
  if (!param && kind != dB200_OP_PROJECTION) return seterr(85, "operator kind %d needs its parameter struct", kind);
  if (!rule || !spaces || !efs || !coords || !op) return seterr(85, "dB200OpCreate: null argument");
  dB200Op o = new dB200Op_s;
  o->ctx = ctx; o->kind = kind; o->rule = rule; o->nspaces = nspaces;
  for (int i = 0; i < nspaces; i++) {
    o->space[i] = spaces[i]; o->efs[i] = efs[i];
    if (efs[i]->rule != rule) { delete o; return seterr(63, "dEFS %d is tabulated on a different rule", i); }
    if (spaces[i]->nelem != spaces[0]->nelem) { delete o; return seterr(63, "spaces disagree on the element count"); }
    if (spaces[i]->P != efs[i]->P[0]) { delete o; return seterr(63, "space %d has P=%d but its dEFS has P=%d", i, spaces[i]->P, efs[i]->P[0]); }
  }
  o->nelem = spaces[0]->nelem; o->NQ = rule->Q[0];
  switch (kind) {
  case dB200_OP_ELLIP:
    if (nspaces != 1 || spaces[0]->bs != 1) { delete o; return seterr(63, "ellip wants one scalar space"); }
    o->ellip = *(const dB200EllipParam *)param; o->nstash = ELLIP_NSTASH; o->nforcing = 1;
    break;
  case dB200_OP_ELAST:
    if (nspaces != 1 || spaces[0]->bs != 3) { delete o; return seterr(63, "elast wants one bs=3 space"); }
    o->elast = *(const dB200ElastParam *)param; o->nstash = ELAST_NSTASH; o->nforcing = 3;
    break;
  case dB200_OP_STOKES:
    if (nspaces != 2 || spaces[0]->bs != 3 || spaces[1]->bs != 1) { delete o; return seterr(63, "stokes wants a bs=3 velocity space and a bs=1 pressure space"); }
    if (spaces[1]->P != spaces[0]->P - 1 && !(spaces[1]->P == spaces[0]->P - 2 && spaces[1]->P >= 2)) { delete o; return seterr(56, "stokes: -pressure_codim 1 and 2 (pressure degree = velocity degree - 1 or - 2, at least 1) are instantiated"); }
    o->stokes = *(const dB200StokesParam *)param; o->nstash = STOKES_NSTASH; o->nforcing = 4;
    break;
  case dB200_OP_PROJECTION:
    if (nspaces != 1 || spaces[0]->bs != 1) { delete o; return seterr(63, "projection wants one scalar space"); }
    o->ellip = dB200EllipParam{2.0, 1.0, 0.0}; o->nstash = 0; o->nforcing = 1;
    break;
  default: delete o; return seterr(56, "operator kind %d not implemented", kind);
  }
  Tables t;
  int err = efs_tables(efs[0], &t);
  if (err) { delete o; return err; }
  if (t.NQ < DOHP_NMIN || t.NQ > DOHP_NMAX) { delete o; return seterr(56, "no kernel instantiated for Q=%d", t.NQ); }
  err = upload(ctx, coords, (size_t)o->nelem * 24, &o->d_coords);
  if (err) { if (o->d_coords) cudaFree(o->d_coords); delete o; return err; }
  // det J > 0 everywhere (jacobi.c:397)
  void *dq;
  CHK(scratch_get(ctx, 3, (o->NQ + 1) * sizeof(double), &dq));
  CU(cudaMemcpyAsync(dq, rule->coord[0].data(), o->NQ * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  int *flag = (int *)((double *)dq + o->NQ);
  CU(cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
  const long npts = (long)o->nelem * o->NQ * o->NQ * o->NQ;
  quad_points_kernel<<<nblk(npts, 256), 256, 0, ctx->stream>>>(o->nelem, o->NQ, (const double *)dq, o->d_coords, nullptr, flag);
  CU(cudaGetLastError());
  LAUNCHED(ctx);
  if (o->nelem) { affine_check_kernel<<<nblk(o->nelem, 256), 256, 0, ctx->stream>>>(o->nelem, o->d_coords, flag); LAUNCHED(ctx); }
  CU(cudaGetLastError());
  int hflag = 0;
  CU(cudaMemcpyAsync(&hflag, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (hflag & 1) { cudaFree(o->d_coords); delete o; return seterr(1, "Negative Jacobian"); }
  o->affine = !(hflag & 2) && !getenv("DOHP_B200_NO_AFFINE");
  err = op_build_eorder(o);
  if (err) { cudaFree(o->d_coords); delete o; return err; }
  *op = o;
  return 0;
}
extern "C" int dB200OpDestroy(dB200Op *op)
{
  if (!op || !*op) return 0;
  dB200Op o = *op;
  cudaSetDevice(o->ctx->device);
  cudaStreamSynchronize(o->ctx->stream);
  cudaFree(o->d_coords);
  if (o->d_eorder) cudaFree(o->d_eorder);
  dohp_q1_free(o);
  dohp_dense_free(o);
  if (o->d_stash) cudaFree(o->d_stash);
  if (o->d_forcing) cudaFree(o->d_forcing);
  if (o->ev_done) { cudaEventSynchronize((cudaEvent_t)o->ev_done); cudaEventDestroy((cudaEvent_t)o->ev_in); cudaEventDestroy((cudaEvent_t)o->ev_out); cudaEventDestroy((cudaEvent_t)o->ev_done); }
  for (int i = 0; i < DB200_MAXFIELDS; i++) { if (o->d_gx[i]) cudaFree(o->d_gx[i]); if (o->d_gy[i]) cudaFree(o->d_gy[i]); }
  delete o;
  *op = nullptr;
  return 0;
}
extern "C" int dB200OpGetQuadraturePoints(dB200Op o, double *xq)
{
  dB200Ctx c = o->ctx;
  CU(cudaSetDevice(c->device));
  const long npts = (long)o->nelem * o->NQ * o->NQ * o->NQ;
  void *dx, *dq;
  bool host;
  CHK(stage_out(c, 0, xq, npts * 3 * sizeof(double), false, &dx, &host));
  CHK(scratch_get(c, 3, (o->NQ + 1) * sizeof(double), &dq));
  CU(cudaMemcpyAsync(dq, o->rule->coord[0].data(), o->NQ * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  int *flag = (int *)((double *)dq + o->NQ);
  quad_points_kernel<<<nblk(npts, 256), 256, 0, c->stream>>>(o->nelem, o->NQ, (const double *)dq, o->d_coords, (double *)dx, flag);
  CU(cudaGetLastError());
  LAUNCHED(c);
  return finish_out(c, xq, dx, npts * 3 * sizeof(double), host);
}
extern "C" int dB200OpSetForcing(dB200Op o, const double *f)
{
  dB200Ctx c = o->ctx;
  CU(cudaSetDevice(c->device));
  if (!f) { if (o->d_forcing) { CU(cudaStreamSynchronize(c->stream)); cudaFree(o->d_forcing); o->d_forcing = nullptr; } return 0; }
  const long cnt = (long)o->nelem * o->NQ * o->NQ * o->NQ * o->nforcing;
  if (!o->d_forcing) CU(cudaMalloc((void **)&o->d_forcing, cnt * sizeof(double)));
  const void *df;
  CHK(stage_in(c, 0, f, cnt * sizeof(double), &df));
  to_planes_kernel<<<nblk(cnt, 256), 256, 0, c->stream>>>(o->nelem, o->NQ, o->nforcing, (const double *)df, o->d_forcing);
  CU(cudaGetLastError());
  LAUNCHED(c);
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}
extern "C" int dB200OpSetDirichlet(dB200Op o, int ifield, const double *vals)
{
  dB200Ctx c = o->ctx;
  CU(cudaSetDevice(c->device));
  if (ifield < 0 || ifield >= o->nspaces) return seterr(63, "field %d", ifield);
  dB200Space s = o->space[ifield];
  const size_t naux = (size_t)(s->nc + s->ngh - s->n) * s->bs, ndir = (size_t)(s->nc - s->n) * s->bs;
  if (!s->d_xaux_inhomog) {
    CU(cudaMalloc((void **)&s->d_xaux_inhomog, std::max<size_t>(naux, 1) * sizeof(double)));
    CU(cudaMemsetAsync(s->d_xaux_inhomog, 0, std::max<size_t>(naux, 1) * sizeof(double), c->stream));
  }
  if (ndir && vals) CU(cudaMemcpyAsync(s->d_xaux_inhomog, vals, ndir * sizeof(double), is_device_ptr(vals) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

static int ensure_stash(dB200Op o)
{
  if (o->d_stash) return 0;
  const size_t cnt = (size_t)o->nelem * o->NQ * o->NQ * o->NQ * o->nstash;
  CU(cudaMalloc((void **)&o->d_stash, std::max<size_t>(cnt, 1) * sizeof(double)));
  return 0;
}

static int ellip_run(dB200Op o, int mode, const double *x, double *y, int hmode)
{
  dB200Ctx c = o->ctx;
  dB200Space s = o->space[0];
  EllipLaunch L;
  CHK(efs_tables(o->efs[0], &L.tab));
  L.ctx = c; L.affine = o->affine; L.mode = mode; L.nelem = o->nelem; L.n = s->n; L.nc = s->nc; L.e2l = s->d_e2l; L.coords = o->d_coords;
  L.x = x; L.y = y; L.yaux = s->d_yaux;
  if (hmode == dB200_INHOMOGENEOUS) {
    if (!s->d_xaux_inhomog) CHK(dB200OpSetDirichlet(o, 0, nullptr));      // no Dirichlet data committed: zeros
    L.xaux = s->d_xaux_inhomog;
  } else L.xaux = s->d_xaux;
  L.stash = o->d_stash; L.forcing = o->d_forcing;
  L.pm2h = (o->ellip.p - 2) / 2; L.eps2 = o->ellip.epsilon * o->ellip.epsilon; L.lambda = o->ellip.lambda;
  const size_t nb = (size_t)s->n * s->bs;
  CU(cudaMemsetAsync(y, 0, nb * sizeof(double), c->stream));                                    // VecZeroEntries, ellip.c:421,452
  if (s->ngh) CU(cudaMemsetAsync(s->d_yaux + (size_t)(s->nc - s->n) * s->bs, 0, (size_t)s->ngh * s->bs * sizeof(double), c->stream));
  auto launch = [&](const int *elist, int count, cudaStream_t st) -> int {
    if (!count) return 0;
    L.elist = elist; L.count = count;
    int err = ellip_fns[L.tab.NQ](L, st);
    if (err == -1) return seterr(56, "no ellip kernel instantiated for (P,Q)=(%d,%d)", L.tab.NP, L.tab.NQ);
    if (err == -2) return seterr(56, "the 1-D tables are not symmetric under the reflection of the reference interval: no kernel instantiated for such a basis");
    if (err == -3) return 77;
    if (err) return seterr(700 + err, "ellip kernel launch: %s", cudaGetErrorString((cudaError_t)err));
    LAUNCHED(c);
    return 0;
  };
  dB200Space sps[1] = {s};
  const double *fx[1] = {x}, *ra[1] = {s->d_yaux};
  double *fa[1] = {const_cast<double *>(L.xaux)}, *ry[1] = {y};
  return run_overlapped(o, 1, sps, fx, fa, ra, ry, launch);
}

// elast (one space) and stokes (velocity + pressure): XFunction / XShellMatMult / stokes blocks
static int vec_run(dB200Op o, bool function, int block, const double *const x[], double *const y[])
{
  dB200Ctx c = o->ctx;
  const bool stokes = o->kind == dB200_OP_STOKES;
  dB200Space su = o->space[0], sp = stokes ? o->space[1] : nullptr;
  VecLaunch L;
  memset(&L, 0, sizeof L);
  L.ctx = c; L.affine = o->affine;
  CHK(efs_tables(o->efs[0], &L.tab));
  if (stokes) CHK(efs_tables(o->efs[1], &L.tabp)); else L.tabp = L.tab;
  L.kind = stokes ? VEC_STOKES : VEC_ELAST;
  L.mode = function ? VEC_FUNCTION : VEC_JACOBIAN;
  if (!stokes) L.flags = USE_U | OUT_U;
  else if (function || block == dB200_STOKES_J) L.flags = USE_U | USE_P | OUT_U | OUT_P;
  else if (block == dB200_STOKES_A) L.flags = USE_U | OUT_U | BLOCK_A;
  else if (block == dB200_STOKES_B) L.flags = USE_U | OUT_P;
  else if (block == dB200_STOKES_Bt) L.flags = USE_P | OUT_U;
  else return seterr(63, "stokes block %d", block);
  if ((L.flags & USE_U) && !x[0]) return seterr(85, "null velocity input");
  if ((L.flags & OUT_U) && !y[0]) return seterr(85, "null velocity output");
  if ((L.flags & USE_P) && !x[1]) return seterr(85, "null pressure input");
  if ((L.flags & OUT_P) && !y[1]) return seterr(85, "null pressure output");
  const int hmode = function ? dB200_INHOMOGENEOUS : dB200_HOMOGENEOUS;
  dB200Space sps[2] = {su, sp};
  const double *xaux[2] = {nullptr, nullptr};
  for (int f = 0; f < (stokes ? 2 : 1); f++) {
    dB200Space s = sps[f];
    if (hmode == dB200_INHOMOGENEOUS) {
      if (!s->d_xaux_inhomog) CHK(dB200OpSetDirichlet(o, f, nullptr));
      xaux[f] = s->d_xaux_inhomog;
    } else xaux[f] = s->d_xaux;
    const bool in = L.flags & (f ? USE_P : USE_U), out = L.flags & (f ? OUT_P : OUT_U);
    if (out) {
      CU(cudaMemsetAsync(y[f], 0, (size_t)s->n * s->bs * sizeof(double), c->stream));
      if (s->ngh) CU(cudaMemsetAsync(s->d_yaux + (size_t)(s->nc - s->n) * s->bs, 0, (size_t)s->ngh * s->bs * sizeof(double), c->stream));
    }
    (void)in;
  }
  L.nelem = o->nelem; L.n_u = su->n; L.nc_u = su->nc; L.e2l_u = su->d_e2l; L.coords = o->d_coords;
  L.xu = x[0]; L.xuaux = xaux[0]; L.yu = y[0]; L.yuaux = su->d_yaux;
  if (stokes) { L.n_p = sp->n; L.nc_p = sp->nc; L.e2l_p = sp->d_e2l; L.xp = x[1]; L.xpaux = xaux[1]; L.yp = y[1]; L.ypaux = sp->d_yaux; }
  L.stash = o->d_stash; L.forcing = o->d_forcing;
  if (stokes) { L.c0 = o->stokes.B; L.c1 = 0.5 * o->stokes.eps * o->stokes.eps; L.c2 = 0.5 * (o->stokes.p - 2); L.c3 = o->stokes.nu / (1 - 2 * o->stokes.nu); }
  else { L.c0 = o->elast.lambda; L.c1 = o->elast.mu; }
  auto launch = [&](const int *elist, int count, cudaStream_t st) -> int {
    if (!count) return 0;
    L.elist = elist; L.count = count;
    int err = vec_fns[L.tab.NQ](L, st);
    if (err == -1) return seterr(56, "no kernel instantiated for (P,Q)=(%d,%d)", L.tab.NP, L.tab.NQ);
    if (err == -2) return seterr(56, "the 1-D tables are not symmetric under the reflection of the reference interval: no kernel instantiated for such a basis");
    if (err == -3) return 77;
    if (err) return seterr(700 + err, "kernel launch: %s", cudaGetErrorString((cudaError_t)err));
    LAUNCHED(c);
    return 0;
  };
  const int nsp = stokes ? 2 : 1;
  const double *fx[2] = {(L.flags & USE_U) ? x[0] : nullptr, stokes && (L.flags & USE_P) ? x[1] : nullptr};
  double *fa[2] = {const_cast<double *>(xaux[0]), const_cast<double *>(xaux[1])};
  const double *ra[2] = {su->d_yaux, stokes ? sp->d_yaux : nullptr};
  double *ry[2] = {(L.flags & OUT_U) ? y[0] : nullptr, stokes && (L.flags & OUT_P) ? y[1] : nullptr};
  return run_overlapped(o, nsp, sps, fx, fa, ra, ry, launch);
}

// the traversal itself on device vectors (ctx stream)
static int op_compute(dB200Op o, bool function, int block, const void *const dx[], void *const dy[])
{
  switch (o->kind) {
  case dB200_OP_ELLIP: {
    if (!dx[0] || !dy[0]) return seterr(85, "null vector");
    int mode;
    if (function) { CHK(ensure_stash(o)); mode = ELLIP_FUNCTION; o->stash_valid = true; }
    else if (o->ellip.p == 2 && o->ellip.lambda == 0) mode = ELLIP_LINEAR;     // eta=1, s=0, lambda e^u=0 exactly: no stash traffic
    else {
      if (!o->stash_valid) return seterr(73, "matrix-free Jacobian needs the stash of a residual evaluation first (ellip.c:360)");
      mode = ELLIP_JACOBIAN;
    }
    CHK(ellip_run(o, mode, (const double *)dx[0], (double *)dy[0], function ? dB200_INHOMOGENEOUS : dB200_HOMOGENEOUS));
  } break;
  case dB200_OP_PROJECTION:
    if (!dx[0] || !dy[0]) return seterr(85, "null vector");
    CHK(ellip_run(o, function ? ELLIP_PROJ_FUNCTION : ELLIP_PROJ_MULT, (const double *)dx[0], (double *)dy[0], function ? dB200_INHOMOGENEOUS : dB200_HOMOGENEOUS));
    break;
  case dB200_OP_ELAST:
  case dB200_OP_STOKES: {
    if (function) { CHK(ensure_stash(o)); o->stash_valid = true; }
    else if (!o->stash_valid) return seterr(73, "matrix-free Jacobian needs the stash of a residual evaluation first");
    const double *xs[2] = {(const double *)dx[0], (const double *)dx[1]};
    double *ys[2] = {(double *)dy[0], (double *)dy[1]};
    CHK(vec_run(o, function, block, xs, ys));
  } break;
  default: return seterr(56, "operator kind %d", o->kind);
  }
  return 0;
}

static int op_run(dB200Op o, bool function, int block, const double *const gx[], double *const gy[])
{
  dB200Ctx c = o->ctx;
  CU(cudaSetDevice(c->device));
  const void *dx[DB200_MAXFIELDS] = {nullptr};
  void *dy[DB200_MAXFIELDS] = {nullptr};
  bool host[DB200_MAXFIELDS] = {false};
  for (int i = 0; i < o->nspaces; i++) {
    const size_t nb = (size_t)o->space[i]->n * o->space[i]->bs * sizeof(double);
    CHK(stage_in(c, 2 * i, gx[i], nb, &dx[i]));
    CHK(stage_out(c, 2 * i + 1, gy[i], nb, false, &dy[i], &host[i]));
  }
  CHK(op_compute(o, function, block, dx, dy));
  for (int i = 0; i < o->nspaces; i++) CHK(finish_out(c, gy[i], dy[i], (size_t)o->space[i]->n * o->space[i]->bs * sizeof(double), host[i]));
  return 0;
}

// Asynchronous MatMult for host vectors (pinned memory): upload on a copy stream, the traversal on the context's stream,
// download on a second copy stream, ordered by events; returns immediately.  Several operators can be in flight, so the
// PCIe directions and the kernels of independent products overlap.  The vectors must stay valid until dB200Synchronize;
// each operator owns its staging buffers, and a second call on the same operator waits for the first one's download.
extern "C" int dB200OpMultAsync(dB200Op o, const double *const gx[], double *const gy[])
{
  dB200Ctx c = o->ctx;
  CU(cudaSetDevice(c->device));
  if (!c->h2d) { CU(cudaStreamCreateWithFlags(&c->h2d, cudaStreamNonBlocking)); CU(cudaStreamCreateWithFlags(&c->d2h, cudaStreamNonBlocking)); }
  if (!o->ev_in) {
    CU(cudaEventCreateWithFlags((cudaEvent_t *)&o->ev_in, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags((cudaEvent_t *)&o->ev_out, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags((cudaEvent_t *)&o->ev_done, cudaEventDisableTiming));
    CU(cudaEventRecord((cudaEvent_t)o->ev_done, c->d2h));
  }
  const void *dx[DB200_MAXFIELDS] = {nullptr};
  void *dy[DB200_MAXFIELDS] = {nullptr};
  bool host[DB200_MAXFIELDS] = {false};
  CU(cudaStreamWaitEvent(c->h2d, (cudaEvent_t)o->ev_done, 0));                 // the previous product of this operator has left the staging buffers
  for (int i = 0; i < o->nspaces; i++) {
    const size_t nb = (size_t)o->space[i]->n * o->space[i]->bs * sizeof(double);
    if (!gx[i] || is_device_ptr(gx[i])) dx[i] = gx[i];
    else {
      if (!o->d_gx[i]) CU(cudaMalloc((void **)&o->d_gx[i], std::max<size_t>(nb, 8)));
      CU(cudaMemcpyAsync(o->d_gx[i], gx[i], nb, cudaMemcpyHostToDevice, c->h2d));
      dx[i] = o->d_gx[i];
    }
    if (!gy[i] || is_device_ptr(gy[i])) dy[i] = gy[i];
    else {
      if (!o->d_gy[i]) CU(cudaMalloc((void **)&o->d_gy[i], std::max<size_t>(nb, 8)));
      dy[i] = o->d_gy[i];
      host[i] = true;
    }
  }
  CU(cudaEventRecord((cudaEvent_t)o->ev_in, c->h2d));
  CU(cudaStreamWaitEvent(c->stream, (cudaEvent_t)o->ev_in, 0));
  CU(cudaStreamWaitEvent(c->stream, (cudaEvent_t)o->ev_done, 0));              // ... and its result buffer has been downloaded
  CHK(op_compute(o, false, 0, dx, dy));
  CU(cudaEventRecord((cudaEvent_t)o->ev_out, c->stream));
  CU(cudaStreamWaitEvent(c->d2h, (cudaEvent_t)o->ev_out, 0));
  for (int i = 0; i < o->nspaces; i++)
    if (host[i]) CU(cudaMemcpyAsync(gy[i], dy[i], (size_t)o->space[i]->n * o->space[i]->bs * sizeof(double), cudaMemcpyDeviceToHost, c->d2h));
  CU(cudaEventRecord((cudaEvent_t)o->ev_done, c->d2h));
  return 0;
}
extern "C" int dB200OpFunction(dB200Op o, const double *const gx[], double *const gy[]) { return op_run(o, true, 0, gx, gy); }
extern "C" int dB200OpMult(dB200Op o, const double *const gx[], double *const gy[]) { return op_run(o, false, 0, gx, gy); }
extern "C" int dB200OpMultBlock(dB200Op o, int block, const double *const gx[], double *const gy[]) { return op_run(o, false, block, gx, gy); }

extern "C" int dB200OpGetStash(dB200Op o, int *nstash, double *stash)
{
  dB200Ctx c = o->ctx;
  CU(cudaSetDevice(c->device));
  if (nstash) *nstash = o->nstash;
  if (!stash) return 0;
  if (!o->d_stash || !o->stash_valid) return seterr(73, "stash not written yet");
  const long cnt = (long)o->nelem * o->NQ * o->NQ * o->NQ * o->nstash;
  void *ds;
  bool host;
  CHK(stage_out(c, 0, stash, cnt * sizeof(double), false, &ds, &host));
  from_planes_kernel<<<nblk(cnt, 256), 256, 0, c->stream>>>(o->nelem, o->NQ, o->nstash, o->d_stash, (double *)ds);
  CU(cudaGetLastError());
  LAUNCHED(c);
  return finish_out(c, stash, ds, cnt * sizeof(double), host);
}


extern "C" int dB200MeasureFP64Peak(dB200Ctx c, double *tflops)
{
  CU(cudaSetDevice(c->device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, c->device));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
  void *out;
  CHK(scratch_get(c, 7, (size_t)blocks * threads * sizeof(double), &out));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
  double best = 0;
  for (int rep = 0; rep < 4; rep++) {
    CU(cudaEventRecord(e0, c->stream));
    dfma_peak_kernel<<<blocks, threads, 0, c->stream>>>((double *)out, iters, 0.999999, 1e-9);
    CU(cudaEventRecord(e1, c->stream));
    CU(cudaEventSynchronize(e1));
    LAUNCHED(c);
    float ms;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 8 * 16 * (double)iters * blocks * threads;
    if (rep) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *tflops = best;
  return 0;
}
