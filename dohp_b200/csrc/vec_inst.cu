// Instantiations of the elasticity / Stokes kernel for NP = NQ = DOHP_N (pressure NPP = NP-1, -pressure_codim 1).
#include "internal.h"
#include "vec_kernel.cuh"
#include "launch.h"
#include <algorithm>
#include <cstring>
#ifndef DOHP_N
#error "compile with -DDOHP_N=<quadrature points per direction>"
#endif
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

namespace dohp {

template <int NP, int NPP, int NQ, int KIND, int MODE, bool AFFINE> static int run_vec(const VecLaunch &L, cudaStream_t s)
{
  using C = VecCfg<NQ, KIND>;
  VecArgs<NP, NPP, NQ> a;
  double viol = std::max(mat1d_fill(a.bt, L.tab.B), 0.1 * mat1d_fill(a.rt.D, L.tab.Dt));
  for (int i = 0; i < NQ; i++) { a.rt.w[i] = L.tab.w[i]; a.rt.q[i] = L.tab.q[i]; }
  if (KIND == VEC_STOKES) viol = std::max(viol, mat1d_fill(a.bp, L.tabp.B));
  else memset(&a.bp, 0, sizeof a.bp);
  if (viol > 1e-12) return -2;
  if (dohp_smem_maps(L.ctx, NQ, Lay2<NQ>::SA, Lay2<NQ>::SB, C::EPB, C::THREADS, C::ES_BASE, &a.ES, &a.maps)) return -3;
  a.nelem = L.nelem; a.count = L.elist ? L.count : L.nelem; a.elist = L.elist; a.flags = L.flags; a.n_u = L.n_u; a.nc_u = L.nc_u; a.n_p = L.n_p; a.nc_p = L.nc_p;
  a.e2l_u = L.e2l_u; a.e2l_p = L.e2l_p; a.coords = L.coords; a.xu = L.xu; a.xuaux = L.xuaux; a.xp = L.xp; a.xpaux = L.xpaux;
  a.yu = L.yu; a.yuaux = L.yuaux; a.yp = L.yp; a.ypaux = L.ypaux; a.stash = L.stash; a.forcing = L.forcing;
  a.c0 = L.c0; a.c1 = L.c1; a.c2 = L.c2; a.c3 = L.c3;
  const size_t smem = sizeof(double) * ((size_t)C::EPB * a.ES + (MODE == VEC_JACOBIAN ? (size_t)C::STAGE_DOUBLES : 0));
  auto kern = vec_kernel<NP, NPP, NQ, KIND, MODE, AFFINE>;
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return (int)err;
  err = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (err != cudaSuccess) return (int)err;
  const int grid = (a.count + C::EPB - 1) / C::EPB;
  if (grid == 0) return 0;
  kern<<<grid, C::THREADS, smem, s>>>(a);
  return (int)cudaGetLastError();
}

template <int N, int NPP, bool AFFINE> static int pick_elast(const VecLaunch &L, cudaStream_t s)
{
  if (L.mode == VEC_FUNCTION) return run_vec<N, NPP, N, VEC_ELAST, VEC_FUNCTION, AFFINE>(L, s);
  return run_vec<N, NPP, N, VEC_ELAST, VEC_JACOBIAN, AFFINE>(L, s);
}
template <int N, int NPP, bool AFFINE> static int pick_stokes(const VecLaunch &L, cudaStream_t s)
{
  if (L.mode == VEC_FUNCTION) return run_vec<N, NPP, N, VEC_STOKES, VEC_FUNCTION, AFFINE>(L, s);
  return run_vec<N, NPP, N, VEC_STOKES, VEC_JACOBIAN, AFFINE>(L, s);
}

int CAT(launch_vec_, DOHP_N)(const VecLaunch &L, cudaStream_t s)
{
  constexpr int N = DOHP_N, NP1 = N > 2 ? N - 1 : 2, NP2 = N > 3 ? N - 2 : 2;
  if (L.tab.NQ != N || L.tab.NP != N) return -1;
  if (L.kind == VEC_ELAST) return L.affine ? pick_elast<N, NP1, true>(L, s) : pick_elast<N, NP1, false>(L, s);
  if (L.tabp.NQ != N) return -1;
  // -pressure_codim 1 (every test of the reference) and -pressure_codim 2 (the default of stokes.c:102-129)
  if (N >= 3 && L.tabp.NP == N - 1) return L.affine ? pick_stokes<N, NP1, true>(L, s) : pick_stokes<N, NP1, false>(L, s);
  if (N >= 4 && L.tabp.NP == N - 2) return L.affine ? pick_stokes<N, NP2, true>(L, s) : pick_stokes<N, NP2, false>(L, s);
  return -1;
}

}  // namespace dohp
