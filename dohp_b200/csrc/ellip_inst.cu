// Instantiations of the fused elliptic kernel for NP = NQ = DOHP_N (FAST Gauss rule: 1 + 2p/2 = p+1 points,
// src/fs/interface/fsruleset.c:32-37, tensorquad.c:81-86).
#include "ellip_kernel.cuh"
#include "launch.h"
#ifndef DOHP_N
#error "compile with -DDOHP_N=<quadrature points per direction>"
#endif
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

namespace dohp {

template <int NP, int NQ, int MODE> static int run_ellip(const EllipLaunch &L, cudaStream_t s)
{
  using C = EllipCfg<NQ>;
  EllipArgs<NP, NQ> a;
  fill_basis(a.bt, L.tab, NP, NQ);
  fill_rule(a.rt, L.tab, NQ);
  a.nelem = L.nelem; a.n = L.n; a.nc = L.nc; a.e2l = L.e2l; a.coords = L.coords; a.x = L.x; a.xaux = L.xaux; a.y = L.y; a.yaux = L.yaux;
  a.stash = L.stash; a.forcing = L.forcing; a.pm2h = L.pm2h; a.eps2 = L.eps2; a.lambda = L.lambda;
  const size_t smem = sizeof(double) * C::SMEM_DOUBLES;
  cudaError_t err = cudaFuncSetAttribute(ellip_kernel<NP, NQ, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return (int)err;
  err = cudaFuncSetAttribute(ellip_kernel<NP, NQ, MODE>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (err != cudaSuccess) return (int)err;
  const int grid = (L.nelem + C::EPB - 1) / C::EPB;
  ellip_kernel<NP, NQ, MODE><<<grid, C::THREADS, smem, s>>>(a);
  return (int)cudaGetLastError();
}

int CAT(launch_ellip_, DOHP_N)(const EllipLaunch &L, cudaStream_t s)
{
  constexpr int N = DOHP_N;
  if (L.tab.NQ != N || L.tab.NP != N) return -1;
  switch (L.mode) {
  case ELLIP_LINEAR: return run_ellip<N, N, ELLIP_LINEAR>(L, s);
  case ELLIP_JACOBIAN: return run_ellip<N, N, ELLIP_JACOBIAN>(L, s);
  case ELLIP_FUNCTION: return run_ellip<N, N, ELLIP_FUNCTION>(L, s);
  }
  return -1;
}

}  // namespace dohp
