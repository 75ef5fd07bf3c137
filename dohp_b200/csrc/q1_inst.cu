// Instantiations of the embedded-Q1 assembly kernel for degree P = DOHP_N - 1 (one translation unit per degree).
#include "q1_elast_kernel.cuh"
#ifndef DOHP_N
#error "compile with -DDOHP_N=<basis nodes per direction>"
#endif
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

namespace dohp {

template <class K> static int run_q1(K kernel, const Q1Launch &L, int epb, int threads, size_t smem, cudaStream_t s)
{
  cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return (int)err;
  err = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (err != cudaSuccess) return (int)err;
  kernel<<<(L.nelem + epb - 1) / epb, threads, smem, s>>>(L);
  return (int)cudaGetLastError();
}

template <int P> static int run_q1_elast(const Q1Launch &L, cudaStream_t s)
{
  {
    using CE = Q1ECfg<P>;
    if (L.kind == Q1_STOKES_A) return run_q1(q1_elast_kernel<P, Q1_STOKES_A>, L, CE::EPB, CE::THREADS, sizeof(double) * CE::SMEM_DOUBLES, s);
    return run_q1(q1_elast_kernel<P, Q1_ELAST>, L, CE::EPB, CE::THREADS, sizeof(double) * CE::SMEM_DOUBLES, s);
  }
}

int CAT(launch_q1_, DOHP_N)(const Q1Launch &L, cudaStream_t s)
{
  constexpr int P = DOHP_N - 1;
  using C = Q1Cfg<P>;
  if (L.P != P) return -1;
  const size_t smem = sizeof(double) * C::SMEM_DOUBLES;
  if (L.kind == Q1_ELLIP) {
    if (L.c2 != 0) return run_q1(q1_ellip_kernel<P, true>, L, C::EPB, C::THREADS, smem, s);
    return run_q1(q1_ellip_kernel<P, false>, L, C::EPB, C::THREADS, smem, s);
  }
  if (L.kind == Q1_ELAST || L.kind == Q1_STOKES_A) return run_q1_elast<P>(L, s);
  return -1;
}

}  // namespace dohp
