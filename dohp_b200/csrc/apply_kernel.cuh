// dEFSApply over a batch of elements (src/jacobi/interface/jacobi.c:489-506 -> dEFSApply_Tensor_Hex,
// src/jacobi/impls/tensor/efstopo.c:368-423), same layouts as the reference: in/out element-contiguous,
// component innermost, gradients [q][d][e].  Built from the same stages as the fused operator kernels so that the
// dEFS-level parity tests exercise exactly the code the operators run.
#pragma once
#include "tensor_stages.cuh"
#include "../../include/dohp_b200.h"

namespace dohp {

template <int NP, int NQ> struct ApplyArgs {
  BasisTab<NP, NQ> bt;
  RuleTab<NQ> rt;
  int nelem, D, amode, add;      // amode: dB200_APPLY_*, add: ADD_VALUES
  const double *jinv;            // [nelem][NQ^3][9] or NULL
  const double *in;
  double *out;
};

template <int NQ> struct ApplyCfg {
  static constexpr int T = NQ * NQ;
  static constexpr int EPB = (T >= 128) ? 1 : (128 / T);
  static constexpr int THREADS = ((EPB * T + 31) / 32) * 32;
  static constexpr int SMEM_DOUBLES = EPB * 3 * Lay<NQ>::CS;
};

template <int NP, int NQ>
__global__ void __launch_bounds__(ApplyCfg<NQ>::THREADS) apply_kernel(const __grid_constant__ ApplyArgs<NP, NQ> P)
{
  using L = Lay<NQ>;
  using C = ApplyCfg<NQ>;
  constexpr int T = C::T, NP3 = NP * NP * NP, NQ3 = NQ * NQ * NQ;
  extern __shared__ double smem[];
  const int le = threadIdx.x / T, t = threadIdx.x - le * T;
  const long e = (long)blockIdx.x * C::EPB + le;
  const bool live = le < C::EPB && e < P.nelem;
  double *A = smem + (live ? le : 0) * 3 * L::CS, *Gx = A + L::CS, *Gy = Gx + L::CS;
  const int ta = t / NQ, tb = t - ta * NQ;
  const int D = P.D;
  const bool fwd = P.amode == dB200_APPLY_INTERP || P.amode == dB200_APPLY_GRAD;
  const bool grad = P.amode == dB200_APPLY_GRAD || P.amode == dB200_APPLY_GRAD_TRANSPOSE;

  for (int d = 0; d < D; d++) {
    if (fwd) {
      if (live && t < NP * NP) {
        const int j = t / NP, k = t - j * NP;
        double r[NP];
#pragma unroll
        for (int i = 0; i < NP; i++) r[i] = P.in[(e * NP3 + i * NP * NP + t) * D + d];
        fwd_x<NP, NQ>(P.bt.B, r, A, j, k);
      }
      __syncthreads();
      if (live && t < NQ * NP) { const int a = t / NP, k = t - a * NP; fwd_y<NP, NQ>(P.bt.B, A, a, k); }
      __syncthreads();
      double u[NQ], gz[NQ];
      if (live) {
        fwd_z<NP, NQ>(P.bt.B, A, ta, tb, u);
        if (grad) { der_z<NQ>(P.rt.Dt, u, gz); put_z<NQ>(A, ta, tb, u); }
      }
      if (grad) {
        __syncthreads();
        if (live) { der_x<NQ>(P.rt.Dt, A, Gx, ta, tb); der_y<NQ>(P.rt.Dt, A, Gy, ta, tb); }
        __syncthreads();
      }
      if (live) {
#pragma unroll
        for (int c = 0; c < NQ; c++) {
          const long q = e * NQ3 + (ta * NQ + tb) * NQ + c;
          if (!grad) {
            double *o = P.out + q * D + d;
            *o = P.add ? *o + u[c] : u[c];
          } else {
            const double g[3] = {Gx[L::at(ta, tb, c)], Gy[L::at(ta, tb, c)], gz[c]};
            double *o = P.out + (q * D + d) * 3;
#pragma unroll
            for (int x = 0; x < 3; x++) {
              double v;
              if (P.jinv) { const double *ji = P.jinv + q * 9; v = g[0] * ji[x] + g[1] * ji[3 + x] + g[2] * ji[6 + x]; }
              else v = g[x];                                  // TensorRuleNoMapping, efstopo.c:518-546
              o[x] = P.add ? o[x] + v : v;
            }
          }
        }
      }
      __syncthreads();
    } else {
      double h[NQ];
      if (live) {
        if (!grad) {
#pragma unroll
          for (int c = 0; c < NQ; c++) h[c] = P.in[(e * NQ3 + (ta * NQ + tb) * NQ + c) * D + d];
        } else {
          double gz[NQ];
#pragma unroll
          for (int c = 0; c < NQ; c++) {
            const long q = e * NQ3 + (ta * NQ + tb) * NQ + c;
            const double *ji = P.jinv + q * 9, *dv = P.in + (q * D + d) * 3;
            Gx[L::at(ta, tb, c)] = dv[0] * ji[0] + dv[1] * ji[1] + dv[2] * ji[2];       // efstopo.c:606
            Gy[L::at(ta, tb, c)] = dv[0] * ji[3] + dv[1] * ji[4] + dv[2] * ji[5];
            gz[c] = dv[0] * ji[6] + dv[1] * ji[7] + dv[2] * ji[8];
            h[c] = 0;
          }
          der_zT_add<NQ>(P.rt.Dt, gz, h);
        }
      }
      if (grad) {
        __syncthreads();
        if (live) { der_xT<NQ>(P.rt.Dt, Gx, ta, tb); der_yT<NQ>(P.rt.Dt, Gy, ta, tb); }
        __syncthreads();
        if (live) {
#pragma unroll
          for (int c = 0; c < NQ; c++) h[c] += Gx[L::at(ta, tb, c)] + Gy[L::at(ta, tb, c)];
        }
      }
      if (live) bwd_z<NP, NQ>(P.bt.B, h, A, ta, tb);
      __syncthreads();
      if (live && t < NQ * NP) { const int a = t / NP, k = t - a * NP; bwd_y<NP, NQ>(P.bt.B, A, a, k); }
      __syncthreads();
      if (live && t < NP * NP) {
        const int j = t / NP, k = t - j * NP;
        double o[NP];
        bwd_x<NP, NQ>(P.bt.B, A, j, k, o);
#pragma unroll
        for (int i = 0; i < NP; i++) {
          double *dst = P.out + (e * NP3 + i * NP * NP + t) * D + d;
          *dst = P.add ? *dst + o[i] : o[i];
        }
      }
      __syncthreads();
    }
  }
}

// dRuleComputePhysical over a batch (src/jacobi/interface/jacobi.c:383-414): one thread per quadrature point
__global__ void compute_physical_kernel(long npts, int Q0, int Q1, int Q2, const double *w0, const double *w1, const double *w2,
                                        const double *jac, double *jinv, double *jw, int *flag);

}  // namespace dohp
