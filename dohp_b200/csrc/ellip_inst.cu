// Instantiations of the fused elliptic kernel for NP = NQ = DOHP_N (FAST Gauss rule: 1 + 2p/2 = p+1 points,
// src/fs/interface/fsruleset.c:32-37, tensorquad.c:81-86).
#include "internal.h"
#include "ellip_kernel.cuh"
#include "launch.h"
#ifndef DOHP_N
#error "compile with -DDOHP_N=<quadrature points per direction>"
#endif
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

namespace dohp {

template <int NP, int NQ, int MODE, bool AFFINE> static int run_ellip(const EllipLaunch &L, cudaStream_t s)
{
  using C = EllipCfg<NQ, MODE, AFFINE>;
  EllipArgs<NP, NQ> a;
  // the even-odd contractions need tables that are (anti)symmetric under the reflection of the reference interval
  const double v1 = mat1d_fill(a.bt, L.tab.B), v2 = mat1d_fill(a.rt.D, L.tab.Dt);
  if (v1 > 1e-12 || v2 > 1e-11) return -2;
  for (int i = 0; i < NQ; i++) { a.rt.w[i] = L.tab.w[i]; a.rt.q[i] = L.tab.q[i]; }
  int err = dohp_smem_maps(L.ctx, NQ, Lay2<NQ>::SA, Lay2<NQ>::SB, C::EPB, C::THREADS, C::ES_BASE, &a.ES, &a.maps);
  if (err) return -3;
  a.nelem = L.nelem; a.count = L.elist ? L.count : L.nelem; a.elist = L.elist; a.n = L.n; a.nc = L.nc; a.e2l = L.e2l; a.coords = L.coords; a.x = L.x; a.xaux = L.xaux; a.y = L.y; a.yaux = L.yaux;
  a.stash = L.stash; a.forcing = L.forcing; a.pm2h = L.pm2h; a.eps2 = L.eps2; a.lambda = L.lambda;
  a.nstage = L.lambda != 0 ? 5 : 4;
  const size_t smem = sizeof(double) * ((size_t)C::EPB * a.ES + (MODE == ELLIP_JACOBIAN ? (size_t)C::stage_doubles(a.nstage) : 0));
  auto kern = ellip_kernel<NP, NQ, MODE, AFFINE>;
  cudaError_t cerr = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (cerr != cudaSuccess) return (int)cerr;
  cerr = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (cerr != cudaSuccess) return (int)cerr;
  const int grid = (a.count + C::EPB - 1) / C::EPB;
  if (grid == 0) return 0;
  kern<<<grid, C::THREADS, smem, s>>>(a);
  return (int)cudaGetLastError();
}

template <int N, bool AFFINE> static int pick_mode(const EllipLaunch &L, cudaStream_t s)
{
  switch (L.mode) {
  case ELLIP_LINEAR: return run_ellip<N, N, ELLIP_LINEAR, AFFINE>(L, s);
  case ELLIP_JACOBIAN: return run_ellip<N, N, ELLIP_JACOBIAN, AFFINE>(L, s);
  case ELLIP_FUNCTION: return run_ellip<N, N, ELLIP_FUNCTION, AFFINE>(L, s);
  case ELLIP_PROJ_FUNCTION: return run_ellip<N, N, ELLIP_PROJ_FUNCTION, AFFINE>(L, s);
  case ELLIP_PROJ_MULT: return run_ellip<N, N, ELLIP_PROJ_MULT, AFFINE>(L, s);
  }
  return -1;
}

int CAT(launch_ellip_, DOHP_N)(const EllipLaunch &L, cudaStream_t s)
{
  constexpr int N = DOHP_N;
  if (L.tab.NQ != N || L.tab.NP != N) return -1;
  return L.affine ? pick_mode<N, true>(L, s) : pick_mode<N, false>(L, s);
}

}  // namespace dohp
