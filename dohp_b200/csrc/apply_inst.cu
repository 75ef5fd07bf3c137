// Instantiations of apply_kernel for quadrature size NQ = DOHP_N and basis sizes NP in {NQ, NQ-1, 2}.
#include "apply_kernel.cuh"
#include "launch.h"
#ifndef DOHP_N
#error "compile with -DDOHP_N=<quadrature points per direction>"
#endif
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

namespace dohp {

template <int NP, int NQ> static int run_apply(const ApplyLaunch &L, cudaStream_t s)
{
  using C = ApplyCfg<NQ>;
  ApplyArgs<NP, NQ> a;
  fill_basis(a.bt, L.tab, NP, NQ);
  fill_rule(a.rt, L.tab, NQ);
  a.nelem = L.nelem; a.D = L.D; a.amode = L.amode; a.add = L.add; a.jinv = L.jinv; a.in = L.in; a.out = L.out;
  const size_t smem = sizeof(double) * C::SMEM_DOUBLES;
  cudaError_t err = cudaFuncSetAttribute(apply_kernel<NP, NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return (int)err;
  const int grid = (L.nelem + C::EPB - 1) / C::EPB;
  apply_kernel<NP, NQ><<<grid, C::THREADS, smem, s>>>(a);
  return (int)cudaGetLastError();
}

int CAT(launch_apply_, DOHP_N)(const ApplyLaunch &L, cudaStream_t s)
{
  constexpr int N = DOHP_N;
  if (L.tab.NQ != N) return -1;
  if (L.tab.NP == N) return run_apply<N, N>(L, s);
  if constexpr (N >= 3) { if (L.tab.NP == N - 1) return run_apply<N - 1, N>(L, s); }
  if constexpr (N >= 4) { if (L.tab.NP == 2) return run_apply<2, N>(L, s); }
  return -1;
}

}  // namespace dohp
