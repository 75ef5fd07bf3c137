// Synthetic structured hex boxes with the index arrays the kernels consume (setup only; nothing here runs per apply).
//
// Restates WHAT the reference's block generator and continuous-space builder produce for a conforming isotropic-degree
// box:
//   * genblock: vertices (i*n+j)*p+k, hexes, faces z-/y-/x-normal, edges z-/y-/x-running, boundary sets 100/200/300,
//     block partition part = (I*N+J)*P+K     (src/fs/mesh/interface/genblock.c:208-420)
//   * dFSBuildSpace_Cont with -dfs_ordering_type natural: entities in MOAB handle order (type-major, creation order),
//     interior nodes of an entity contiguous, local layout [explicit | dirichlet | ghost]
//     (src/fs/impls/cont/cont.c:93-178,293-385; src/fs/interface/fsbuild.c:55-146; src/jacobi/impls/tensor/tensor.c:360-468)
// The same numbering is produced independently (vectorised numpy with a sort) by dohp_b200/box.py, which is checked
// against the reference's golden src/fs/tests/refout/bu-0.refout; tests/test_box.py compares the two.
// Here the order is generated directly: walking entity kinds in creation order and handing out counters per
// category is a single O(nodes) pass, which is what makes 64^3 x p=10 (2.6e8 nodes) a matter of seconds.
#include "internal.h"
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <chrono>
#include <cstdio>
static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

using namespace dohp;

namespace {

struct Box1 { int e0, e1; };                    // element range of a part in one direction
inline Box1 part_range(int nel, int nparts, int i) { return {(int)((long)i * nel / nparts), (int)((long)(i + 1) * nel / nparts)}; }   // genblock.c:286-303
// membership in the partition SETS (inverse of part_range); the per-region tag formula genblock.c:260-266 disagrees when nel % nparts != 0
inline int part_of_element(int e, int nel, int nparts) { return (int)((((long)e + 1) * nparts - 1) / nel); }

struct Grid {
  int ne[3], p, procs[3], mask;
  int NN[3];                                    // global nodes per direction
  bool diri(int gx, int gy, int gz) const
  {
    bool d = false;
    if (mask & 1) d |= gx == 0 || gx == NN[0] - 1 || gy == 0 || gy == NN[1] - 1;
    if (mask & 2) d |= gz == NN[2] - 1;
    if (mask & 4) d |= gz == 0;
    return d;
  }
  int low(int g, int d) const { int e = (g - 1) / p; if (g == 0) e = 0; e = std::min(std::max(e, 0), ne[d] - 1); return part_of_element(e, ne[d], procs[d]); }
  int owner(int gx, int gy, int gz) const { return (low(gx, 0) * procs[1] + low(gy, 1)) * procs[2] + low(gz, 2); }
};

struct Local {                                  // node box and numbering of one rank
  Box1 b[3];
  int g0[3], nn[3];                             // first global node, node count per direction
  long nnodes;
  int n, nc, ngh;
  std::vector<int> local_of_lex;                // local index of every node of the box, lexicographic (x slowest)
  long lex(int gx, int gy, int gz) const { return ((long)(gx - g0[0]) * nn[1] + (gy - g0[1])) * nn[2] + (gz - g0[2]); }
};

// One pass over the entities of rank r's box in natural order (type-major, creation order within a type,
// interior nodes lexicographic in the entity's own axes = global lexicographic restricted to the entity).
void number_rank(const Grid &G, int rank, Local &L)
{
  const int p = G.p, N1 = G.procs[1], N2 = G.procs[2];
  const int rI = rank / (N1 * N2), rJ = (rank / N2) % N1, rK = rank % N2;
  const int rr[3] = {rI, rJ, rK};
  for (int d = 0; d < 3; d++) {
    L.b[d] = part_range(G.ne[d], G.procs[d], rr[d]);
    L.g0[d] = L.b[d].e0 * p;
    L.nn[d] = (L.b[d].e1 - L.b[d].e0) * p + 1;
  }
  L.nnodes = (long)L.nn[0] * L.nn[1] * L.nn[2];
  L.local_of_lex.assign(L.nnodes, -1);
  // the Dirichlet and ownership predicates are separable: tabulate them per direction
  std::vector<signed char> dirf[3], ghf[3];
  for (int d = 0; d < 3; d++) {
    dirf[d].resize(L.nn[d]); ghf[d].resize(L.nn[d]);
    for (int i = 0; i < L.nn[d]; i++) {
      const int g = L.g0[d] + i;
      bool dir;
      if (d < 2) dir = (G.mask & 1) && (g == 0 || g == G.NN[d] - 1);
      else dir = ((G.mask & 2) && g == G.NN[2] - 1) || ((G.mask & 4) && g == 0);
      dirf[d][i] = dir; ghf[d][i] = G.low(g, d) != rr[d];
    }
  }
  // kinds: free[d]=1 when the entity extends along direction d (index there is an element index, nodes 1..p-1)
  static const int kinds[8][3] = {{0, 0, 0},                       // vertices
                                  {0, 0, 1}, {0, 1, 0}, {1, 0, 0},   // edges running z, y, x (genblock.c:352-390)
                                  {1, 1, 0}, {1, 0, 1}, {0, 1, 1},   // faces normal z, y, x (genblock.c:309-347)
                                  {1, 1, 1}};                      // regions
  // Work unit = (kind, I): pass 0 counts the categories of each unit, a serial prefix sum turns the counts into
  // starting indices (natural order = kinds in order, I ascending), pass 1 hands out the indices.  Units are independent.
  struct Unit { int kd, I; long start[3]; };
  std::vector<Unit> units;
  for (int kd = 0; kd < 8; kd++) {
    const int *fr = kinds[kd];
    if ((fr[0] || fr[1] || fr[2]) && p == 1) continue;           // p=1: no interior nodes on edges/faces/regions
    for (int I = L.b[0].e0; I < (fr[0] ? L.b[0].e1 : L.b[0].e1 + 1); I++) units.push_back({kd, I, {0, 0, 0}});
  }
  for (int pass = 0; pass < 2; pass++) {
#pragma omp parallel for schedule(dynamic, 4)
    for (long ui = 0; ui < (long)units.size(); ui++) {
      Unit &U = units[ui];
      const int *fr = kinds[U.kd];
      const int w0 = fr[0] ? p - 1 : 1, w1 = fr[1] ? p - 1 : 1, w2 = fr[2] ? p - 1 : 1, I = U.I;
      long cnt[3] = {U.start[0], U.start[1], U.start[2]};
      if (pass == 0) cnt[0] = cnt[1] = cnt[2] = 0;
      for (int J = L.b[1].e0; J < (fr[1] ? L.b[1].e1 : L.b[1].e1 + 1); J++)
        for (int K = L.b[2].e0; K < (fr[2] ? L.b[2].e1 : L.b[2].e1 + 1); K++)
          for (int a = 0; a < w0; a++) {
            const int ix = I * p + (fr[0] ? a + 1 : 0) - L.g0[0];
            for (int b = 0; b < w1; b++) {
              const int iy = J * p + (fr[1] ? b + 1 : 0) - L.g0[1];
              const int dxy = dirf[0][ix] | dirf[1][iy], gxy = ghf[0][ix] | ghf[1][iy];
              const long row = ((long)ix * L.nn[1] + iy) * L.nn[2];
              for (int c = 0; c < w2; c++) {
                const int iz = K * p + (fr[2] ? c + 1 : 0) - L.g0[2];
                const int ct = (dxy | dirf[2][iz]) ? 1 : ((gxy | ghf[2][iz]) ? 2 : 0);
                if (pass) L.local_of_lex[row + iz] = (int)cnt[ct];
                cnt[ct]++;
              }
            }
          }
      if (pass == 0) { U.start[0] = cnt[0]; U.start[1] = cnt[1]; U.start[2] = cnt[2]; }
    }
    if (pass == 0) {
      long tot[3] = {0, 0, 0};
      for (Unit &U : units) for (int c = 0; c < 3; c++) { const long k = U.start[c]; U.start[c] = tot[c]; tot[c] += k; }
      L.n = (int)tot[0]; L.nc = (int)(tot[0] + tot[1]); L.ngh = (int)tot[2];
      for (Unit &U : units) { U.start[1] += L.n; U.start[2] += L.nc; }       // layout [explicit | dirichlet | ghost]
    }
  }
}

// owned (explicit, non-ghost) node count of rank r in closed form: the ownership and Dirichlet predicates are
// products of 1-D predicates
long count_owned(const Grid &G, int rank)
{
  const int N1 = G.procs[1], N2 = G.procs[2];
  const int rr[3] = {rank / (N1 * N2), (rank / N2) % N1, rank % N2};
  long prod = 1;
  for (int d = 0; d < 3; d++) {
    const Box1 b = part_range(G.ne[d], G.procs[d], rr[d]);
    long c = 0;
    for (int g = b.e0 * G.p; g <= b.e1 * G.p; g++) {
      if (G.low(g, d) != rr[d]) continue;
      bool dir = false;
      if (d < 2) dir = (G.mask & 1) && (g == 0 || g == G.NN[d] - 1);
      else dir = ((G.mask & 2) && g == G.NN[2] - 1) || ((G.mask & 4) && g == 0);
      if (!dir) c++;
    }
    prod *= c;
  }
  return prod;
}

void morph_point(double morph, double twist, double *x)      // Morph, src/fs/tests/ellip.c:173-181
{
  const double X = x[0], Y = x[1], Z = x[2];
  x[0] = cos(twist * Z) * X + sin(twist * Z) * Y;
  x[1] = -sin(twist * Z) * X + cos(twist * Z) * Y;
  x[2] = -1 + (Z + 1) * (1 + morph * ((X + 1) * (X + 1) + (Y + 1) * (Y + 1)));
}

template <class T> T *dup(const std::vector<T> &v)
{
  T *p = (T *)malloc(std::max<size_t>(v.size(), 1) * sizeof(T));
  if (p && !v.empty()) memcpy(p, v.data(), v.size() * sizeof(T));
  return p;
}

}  // namespace

extern "C" int dB200BoxMeshCreate(const dB200BoxSpec *S, dB200BoxMesh **out)
{
  Grid G;
  for (int d = 0; d < 3; d++) {
    G.ne[d] = S->ne[d]; G.procs[d] = S->procs[d] > 0 ? S->procs[d] : 1;
    if (G.ne[d] < 1 || G.procs[d] > G.ne[d]) return seterr(63, "box: ne[%d]=%d procs=%d", d, G.ne[d], G.procs[d]);
  }
  G.p = S->p; G.mask = S->dirichlet_mask;
  if (G.p < 1 || G.p > 13) return seterr(63, "box: degree %d out of range", G.p);
  for (int d = 0; d < 3; d++) G.NN[d] = G.ne[d] * G.p + 1;
  const int nranks = G.procs[0] * G.procs[1] * G.procs[2], rank = S->rank;
  if (rank < 0 || rank >= nranks) return seterr(63, "box: rank %d of %d", rank, nranks);
  const int p = G.p, d1 = p + 1;

  const bool timing = getenv("DOHP_BOX_TIMING") != nullptr;
  double t0 = now_s();
  Local L;
  number_rank(G, rank, L);
  if (timing) { fprintf(stderr, "[box] numbering %.2fs\n", now_s() - t0); t0 = now_s(); }
  if ((double)(L.b[0].e1 - L.b[0].e0) * (L.b[1].e1 - L.b[1].e0) * (L.b[2].e1 - L.b[2].e0) * d1 * d1 * d1 > 2.1e9)
    return seterr(63, "box: more than 2^31 expanded nodes on one rank");
  dB200BoxMesh *M = (dB200BoxMesh *)calloc(1, sizeof(dB200BoxMesh));
  const int le[3] = {L.b[0].e1 - L.b[0].e0, L.b[1].e1 - L.b[1].e0, L.b[2].e1 - L.b[2].e0};
  M->nelem = le[0] * le[1] * le[2];
  M->n = L.n; M->nc = L.nc; M->ngh = L.ngh;
  std::vector<long> counts(nranks);
  long rstart = 0, nglobal = 0;
  for (int r = 0; r < nranks; r++) { counts[r] = r == rank ? L.n : count_owned(G, r); if (r < rank) rstart += counts[r]; nglobal += counts[r]; }   // MPI_Scan, cont.c:326-336
  M->rstart = rstart; M->nglobal = nglobal;

  // ---- e2l: element (ei,ej,ek) lexicographic (k fastest, genblock.c:235-250), node (i,j,k) -> (i*d+j)*d+k (tensor.c:372-373,409)
  const size_t np3 = (size_t)d1 * d1 * d1;
  M->e2l = (int *)malloc(std::max<size_t>(M->nelem * np3, 1) * sizeof(int));
  M->coords = (double *)malloc(std::max<size_t>((size_t)M->nelem * 24, 1) * sizeof(double));
  if (!M->e2l || !M->coords) return seterr(55, "box: out of memory");
#pragma omp parallel for schedule(static)
  for (int e = 0; e < M->nelem; e++) {
    const int ek = e % le[2], ej = (e / le[2]) % le[1], ei = e / (le[2] * le[1]);
    int *dst = M->e2l + (size_t)e * np3;
    for (int i = 0; i < d1; i++)
      for (int j = 0; j < d1; j++) {
        const long row = ((long)(ei * p + i) * L.nn[1] + (ej * p + j)) * L.nn[2] + ek * p;
        for (int k = 0; k < d1; k++) dst[(i * d1 + j) * d1 + k] = L.local_of_lex[row + k];
      }
    // expanded Q1 geometry in tensor order (fsgeom.c:150-152): vertex (di,dj,dk) -> slot 4di+2dj+dk
    for (int di = 0; di < 2; di++)
      for (int dj = 0; dj < 2; dj++)
        for (int dk = 0; dk < 2; dk++) {
          const int vi = L.b[0].e0 + ei + di, vj = L.b[1].e0 + ej + dj, vk = L.b[2].e0 + ek + dk;
          double x[3] = {S->box[0] + (S->box[1] - S->box[0]) * ((double)vi / G.ne[0]),      // genblock.c:219-221
                         S->box[2] + (S->box[3] - S->box[2]) * ((double)vj / G.ne[1]),
                         S->box[4] + (S->box[5] - S->box[4]) * ((double)vk / G.ne[2])};
          if (S->morph != 0 || S->twist != 0) morph_point(S->morph, S->twist, x);
          double *c = M->coords + (size_t)e * 24 + (4 * di + 2 * dj + dk) * 3;
          c[0] = x[0]; c[1] = x[1]; c[2] = x[2];
        }
  }

  if (timing) { fprintf(stderr, "[box] e2l+coords %.2fs\n", now_s() - t0); t0 = now_s(); }
  // ---- nodal coordinates: Q1 geometry at the GLL nodes (dFSGetNodalCoordinates, fsgeom.c:20-106) ------------------------------
  const size_t nloc = (size_t)L.nc + L.ngh;
  M->node_xyz = (S->flags & 1) ? nullptr : (double *)malloc(std::max<size_t>(nloc * 3, 1) * sizeof(double));
  if (M->node_xyz) {
    std::vector<double> z(d1), w(d1);
    gauss_lobatto(d1, z.data(), w.data());
#pragma omp parallel for schedule(static)
    for (int e = 0; e < M->nelem; e++) {
      const double *X = M->coords + (size_t)e * 24;
      const int *idx = M->e2l + (size_t)e * np3;
      for (int i = 0; i < d1; i++)
        for (int j = 0; j < d1; j++)
          for (int k = 0; k < d1; k++) {
            const double a[2] = {0.5 * (1 - z[i]), 0.5 * (1 + z[i])}, b[2] = {0.5 * (1 - z[j]), 0.5 * (1 + z[j])};
            const double c[2] = {0.5 * (1 - z[k]), 0.5 * (1 + z[k])};
            double x[3] = {0, 0, 0};
            for (int di = 0; di < 2; di++)
              for (int dj = 0; dj < 2; dj++)
                for (int dk = 0; dk < 2; dk++)
                  for (int d = 0; d < 3; d++) x[d] += a[di] * b[dj] * c[dk] * X[(4 * di + 2 * dj + dk) * 3 + d];
            double *dst = M->node_xyz + (size_t)idx[(i * d1 + j) * d1 + k] * 3;
            dst[0] = x[0]; dst[1] = x[1]; dst[2] = x[2];           // shared nodes: every element writes the same value up to rounding
          }
    }
  }

  // ---- l2g and halos ----------------------------------------------------------------------------------------------------------
  // l2g is dInt = 32-bit like the reference's (dohptype.h:30-34).  A global numbering that does not fit is an error for anything that
  // assembles; the matrix-free benches (flags bit 1: no ghost numbers wanted) get no l2g at all, so no wrapped index can ever be used.
  const bool l2g_fits = nglobal <= 2147483647L;
  if (!l2g_fits && !(S->flags & 2)) return seterr(63, "box: %ld global nodes do not fit the 32-bit local-to-global map (dInt); matrix-free runs may pass flags bit 1 to skip it", nglobal);
  M->l2g = l2g_fits ? (int *)malloc(std::max<size_t>(nloc, 1) * sizeof(int)) : nullptr;
  if (M->l2g) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)nloc; i++) M->l2g[i] = i < (long)L.n ? (int)(rstart + i) : -1;     // fsbuild.c:27-37
  }
  std::vector<int> nr;
  std::vector<std::vector<int>> sidx, ridx;
  if (nranks > 1) {
    for (int r = 0; r < nranks; r++) {
      if (r == rank) continue;
      Local R;                                   // only its box is needed first
      const int N1 = G.procs[1], N2 = G.procs[2];
      const int rr[3] = {r / (N1 * N2), (r / N2) % N1, r % N2};
      int lo[3], hi[3];
      bool overlap = true;
      for (int d = 0; d < 3; d++) {
        R.b[d] = part_range(G.ne[d], G.procs[d], rr[d]);
        lo[d] = std::max(L.g0[d], R.b[d].e0 * p);
        hi[d] = std::min(L.g0[d] + L.nn[d] - 1, R.b[d].e1 * p);
        if (lo[d] > hi[d]) overlap = false;
      }
      if (!overlap) continue;
      // shared nodes in global lexicographic order: the ones I own go to r (send), the ones r owns come to me (recv)
      std::vector<int> snd, rcv;
      std::vector<long> rcv_lex;
      for (int gx = lo[0]; gx <= hi[0]; gx++)
        for (int gy = lo[1]; gy <= hi[1]; gy++)
          for (int gz = lo[2]; gz <= hi[2]; gz++) {
            if (G.diri(gx, gy, gz)) continue;
            const int ow = G.owner(gx, gy, gz);
            if (ow == rank) snd.push_back(L.local_of_lex[L.lex(gx, gy, gz)]);
            else if (ow == r) { rcv.push_back(L.local_of_lex[L.lex(gx, gy, gz)]); rcv_lex.push_back(((long)gx * G.NN[1] + gy) * G.NN[2] + gz); }
          }
      if (snd.empty() && rcv.empty()) continue;
      if (!rcv.empty() && !(S->flags & 2) && M->l2g) {     // ghost columns: owner's rstart + owner's local index
        number_rank(G, r, R);
        long rs = 0;
        for (int q = 0; q < r; q++) rs += counts[q];
        for (size_t i = 0; i < rcv.size(); i++) {
          const long gl = rcv_lex[i];
          const int gz = (int)(gl % G.NN[2]), gy = (int)((gl / G.NN[2]) % G.NN[1]), gx = (int)(gl / ((long)G.NN[2] * G.NN[1]));
          M->l2g[rcv[i]] = (int)(rs + R.local_of_lex[R.lex(gx, gy, gz)]);
        }
      }
      nr.push_back(r); sidx.push_back(snd); ridx.push_back(rcv);
    }
  }
  M->nneigh = (int)nr.size();
  M->neigh_rank = dup(nr);
  M->sendcount = (int *)malloc(std::max<size_t>(nr.size(), 1) * sizeof(int));
  M->recvcount = (int *)malloc(std::max<size_t>(nr.size(), 1) * sizeof(int));
  M->sendidx = (int **)malloc(std::max<size_t>(nr.size(), 1) * sizeof(int *));
  M->recvidx = (int **)malloc(std::max<size_t>(nr.size(), 1) * sizeof(int *));
  for (size_t i = 0; i < nr.size(); i++) {
    M->sendcount[i] = (int)sidx[i].size(); M->recvcount[i] = (int)ridx[i].size();
    M->sendidx[i] = dup(sidx[i]); M->recvidx[i] = dup(ridx[i]);
  }
  *out = M;
  return 0;
}

extern "C" int dB200BoxMeshDestroy(dB200BoxMesh **pm)
{
  if (!pm || !*pm) return 0;
  dB200BoxMesh *M = *pm;
  free(M->e2l); free(M->l2g); free(M->coords); free(M->node_xyz);
  for (int i = 0; i < M->nneigh; i++) { free(M->sendidx[i]); free(M->recvidx[i]); }
  free(M->neigh_rank); free(M->sendcount); free(M->recvcount); free(M->sendidx); free(M->recvidx);
  free(M);
  *pm = nullptr;
  return 0;
}
