// dEFSApply for ANY tensor-product hex dEFS the reference's dJacobiGetEFS_Tensor can hand out (src/jacobi/impls/tensor/tensor.c:503-545):
// per-direction basis sizes P[3] and rule sizes Q[3] chosen at run time (anisotropic degree, P > Q, the (P,Q) pairs of
// src/jacobi/tests/ex1.c:285-286), explicit interp / deriv tables (no collocation assumption).  Same layouts and semantics as
// dEFSApply_Tensor_Hex (src/jacobi/impls/tensor/efstopo.c:368-423) with TensorMult_Hex_nounroll (inlinetmulthex.h:19-100),
// TensorRuleMapping / NoMapping / MappingTranspose (efstopo.c:518-623).  This is the parity surface for the shapes the templated
// apply_kernel does not instantiate; one block per element, one component at a time, the three 1-D contractions staged through
// shared memory.  Not a bench line.
#include "internal.h"
#include "launch.h"
#include <vector>
#ifndef DI
#define DI __device__ __forceinline__
#endif

namespace dohp {

struct GenApplyArgs {
  int P[3], Q[3], D, nelem, amode, add, dim;
  const double *tab;             // device: per direction interp[Q][P] then deriv[Q][P]
  int toff[3];                   // offset of direction d's interp table in tab (deriv follows at + Q*P)
  const double *jinv, *in;
  double *out;
  int szF, szA, szB;             // shared-memory buffer sizes (doubles)
};

// one 1-D contraction: out[(l, r)] = sum_i M(l,i) in[(i, r)] along the direction with stride `stride` and `nin` entries; the other two
// directions are flattened by (hi, lo): index = (h * n + x) * stride + lo.  M(l,i) = T ? A[i*ncol + l] : A[l*ncol + i]
template <class Emit>
DI void contract(const double *A, bool T, int ncol, int nhi, int nin, int nout, int stride, const double *src, Emit emit)
{
  const int total = nhi * nout * stride;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int lo = idx % stride, l = (idx / stride) % nout, h = idx / (stride * nout);
    const double *s = src + (size_t)h * nin * stride + lo;
    double acc = 0;
    for (int i = 0; i < nin; i++) acc = fma(T ? A[i * ncol + l] : A[l * ncol + i], s[(size_t)i * stride], acc);
    emit(idx, acc);
  }
}

__global__ void __launch_bounds__(256) apply_generic_kernel(const GenApplyArgs G)
{
  extern __shared__ double sm[];
  const int P0 = G.P[0], P1 = G.P[1], P2 = G.P[2], Q0 = G.Q[0], Q1 = G.Q[1], Q2 = G.Q[2], D = G.D;
  const int ntab = G.toff[2] + 2 * Q2 * P2;
  double *tab = sm, *F = tab + ntab, *A = F + G.szF, *B = A + G.szA;
  for (int i = threadIdx.x; i < ntab; i += blockDim.x) tab[i] = G.tab[i];
  const long e = blockIdx.x;
  const int np = P0 * P1 * P2, nq = Q0 * Q1 * Q2;
  const bool fwd = G.amode == dB200_APPLY_INTERP || G.amode == dB200_APPLY_GRAD;
  const bool grad = G.amode == dB200_APPLY_GRAD || G.amode == dB200_APPLY_GRAD_TRANSPOSE;
  const double *jinv = G.jinv ? G.jinv + e * nq * 9 : nullptr;
  for (int d = 0; d < D; d++) {
    for (int k = 0; k < (grad ? 3 : 1); k++) {
      const double *M0 = tab + G.toff[0] + (grad && k == 0 ? Q0 * P0 : 0);
      const double *M1 = tab + G.toff[1] + (grad && k == 1 ? Q1 * P1 : 0);
      const double *M2 = tab + G.toff[2] + (grad && k == 2 ? Q2 * P2 : 0);
      __syncthreads();
      if (fwd) {
        const double *in = G.in + e * np * D;
        for (int i = threadIdx.x; i < np; i += blockDim.x) F[i] = in[(size_t)i * D + d];
        __syncthreads();
        contract(M0, false, P0, 1, P0, Q0, P1 * P2, F, [&](int idx, double v) { A[idx] = v; });
        __syncthreads();
        contract(M1, false, P1, Q0, P1, Q1, P2, A, [&](int idx, double v) { B[idx] = v; });
        __syncthreads();
        double *out = G.out + e * nq * D * (grad ? 3 : 1);
        contract(M2, false, P2, Q0 * Q1, P2, Q2, 1, B, [&](int q, double v) {
          if (!grad) {
            double *o = out + (size_t)q * D + d;
            *o = G.add ? *o + v : v;
          } else {
            double *o = out + ((size_t)q * D + d) * 3;
            if (jinv) {                                   // TensorRuleMapping, efstopo.c:557-591: v[q][d][e] = sum_k u[k][q][d] jinv[q][k][e]
#pragma unroll
              for (int x = 0; x < 3; x++) o[x] = ((k == 0 && !G.add) ? 0.0 : o[x]) + v * jinv[(size_t)q * 9 + k * 3 + x];
            } else o[k] = G.add ? o[k] + v : v;            // TensorRuleNoMapping, efstopo.c:518-546
          }
        });
      } else {
        const double *in = G.in + e * nq * D * (grad ? 3 : 1);
        for (int q = threadIdx.x; q < nq; q += blockDim.x) {
          double v;
          if (grad) {                                     // TensorRuleMappingTranspose, efstopo.c:593-623: u[k][q][d] = sum_e in[q][d][e] jinv[q][k][e]
            const double *g = in + ((size_t)q * D + d) * 3, *j = jinv + (size_t)q * 9 + k * 3;
            v = g[0] * j[0] + g[1] * j[1] + g[2] * j[2];
          } else v = in[(size_t)q * D + d];
          F[q] = v;
        }
        __syncthreads();
        contract(M0, true, P0, 1, Q0, P0, Q1 * Q2, F, [&](int idx, double v) { A[idx] = v; });
        __syncthreads();
        contract(M1, true, P1, P0, Q1, P1, Q2, A, [&](int idx, double v) { B[idx] = v; });
        __syncthreads();
        double *out = G.out + e * np * D;
        contract(M2, true, P2, P0 * P1, Q2, P2, 1, B, [&](int p, double v) {
          double *o = out + (size_t)p * D + d;
          *o = ((k == 0 && !G.add) ? 0.0 : *o) + v;
        });
      }
    }
  }
}

int launch_apply_generic(dB200EFS efs, int nelem, int D, int amode, int add, const double *jinv, const double *in, double *out, cudaStream_t s)
{
  GenApplyArgs G;
  dB200Rule r = efs->rule;
  int off = 0;
  std::vector<double> h;
  for (int d = 0; d < 3; d++) {
    G.P[d] = efs->P[d]; G.Q[d] = r->Q[d];
    G.toff[d] = off;
    h.insert(h.end(), efs->interp[d].begin(), efs->interp[d].end());
    h.insert(h.end(), efs->deriv[d].begin(), efs->deriv[d].end());
    off += 2 * G.P[d] * G.Q[d];
  }
  const int P0 = G.P[0], P1 = G.P[1], P2 = G.P[2], Q0 = G.Q[0], Q1 = G.Q[1], Q2 = G.Q[2];
  const bool fwd = amode == dB200_APPLY_INTERP || amode == dB200_APPLY_GRAD;
  G.szF = fwd ? P0 * P1 * P2 : Q0 * Q1 * Q2;
  G.szA = fwd ? Q0 * P1 * P2 : P0 * Q1 * Q2;
  G.szB = fwd ? Q0 * Q1 * P2 : P0 * P1 * Q2;
  const size_t smem = sizeof(double) * ((size_t)off + G.szF + G.szA + G.szB);
  if (smem > 227 * 1024) return -2;
  if (!efs->d_gtab) {
    if (cudaMalloc(&efs->d_gtab, h.size() * sizeof(double)) != cudaSuccess) return (int)cudaGetLastError();
    cudaError_t ce = cudaMemcpyAsync(efs->d_gtab, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, s);
    if (ce != cudaSuccess) return (int)ce;
    ce = cudaStreamSynchronize(s);                    // h is a local
    if (ce != cudaSuccess) return (int)ce;
  }
  G.tab = efs->d_gtab;
  G.dim = efs->dim;
  G.D = D; G.nelem = nelem; G.amode = amode; G.add = add; G.jinv = jinv; G.in = in; G.out = out;
  cudaError_t ce = cudaFuncSetAttribute(apply_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (ce != cudaSuccess) return (int)ce;
  apply_generic_kernel<<<nelem, 256, smem, s>>>(G);
  return (int)cudaGetLastError();
}

}  // namespace dohp
