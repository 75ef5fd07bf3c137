// Embedded-Q1 assembly kernel, sm_100a, FP64.
//
// Reference path per element (XJacobian, src/fs/tests/ellip.c:474-536 / elast.c:343-407): for each of the p^3 Q1 patches
// on the SPARSE rule (tensorquad.c:120-169) build the 8-function patch basis (dEFSGetExplicitSparse_Tensor_Hex,
// src/jacobi/impls/tensor/efstopo.c:660-717), K[i][j] += interp.v + deriv.dv with the pointwise Jacobian applied to unit
// trial functions (ellip.c:495-525), dFSMatSetValuesBlockedExpanded(ADD) (src/fs/interface/fs.c:826-930).
//
// Here, per element (several small elements per block):
//   B  the state is interpolated to the SPARSE points by sum factorisation (x, y through shared memory; z per point);
//   C  one thread per (patch, unit) forms a per-point 3x3 tensor T_q (the pointwise Jacobian folded with the inverse
//      Jacobian the reference maps the patch basis with, quirk (ii)) and accumulates K[i][j] = sum_q d_i^T T_q d_j
//      (- c_q phi_i phi_j) in registers; the reference-space patch basis d_b(q), phi_b(q) is built from interp1 =
//      {0.788..,0.211..} and deriv1 = +-1 (quirk (i), efstopo.c:665-666).  ellip: 1 unit, symmetric, 36 entries.
//   G  the patch matrices of a slab of patch layers are reduced inside the element through shared memory: one thread per
//      row node adds the entries of the <= 8 patches around it into 27 neighbour slots (registers); the 32 x 27 sums of a
//      warp go through a per-warp staging area together with the rows' records so that 8 consecutive lanes write 8
//      consecutive CSR entries of one row: ONE RED.ADD.F64 per CSR entry and (element, slab).  The positions come from a
//      per-(element, node) record: 27 one-byte offsets into the row + the row's CSR start (what MatSetPreallocationCOO
//      precomputes).  (A variant that used plain stores for the entries no other element touches was measured: no gain.)
// HBM traffic per element: 8 B per CSR entry touched (+ the zeroing pass), 32 B record + 4 B index + 8 B state per node.
#pragma once
#include "tensor_stages.cuh"

namespace dohp {

enum { Q1_ELLIP = 0, Q1_ELAST = 1, Q1_STOKES_A = 2 };

struct Q1Launch {
  int P, kind;
  int nelem, n_own, nc;          // elements; owned explicit nodes; closure size (local nodes >= nc are ghosts)
  const int *e2l;
  const double *coords, *x, *xaux;
  const double *tab;             // device: Bs[2P][n], Ds[2P][n], qs[2P], ws[2P] (SPARSE rule, basis tabulated on it)
  const int *rowptr;
  const unsigned char *noff;     // [nelem*n^3][32]: 27 offsets into the row (0xFF = dropped), pad, int32 CSR start of the row (-1: no row)
  double *vals;
  double c0, c1, c2, c3;         // ellip: (p-2)/2, eps^2, lambda ; elast: lambda, mu ; stokes: B, eps^2/2, (p-2)/2, nu/(1-2nu)
  int reduced;                   // -dfs_assemble_reduced (fs.c:905-917): only the bs diagonal of every block is assembled
  int dbg;                       // measurement ablations (DOHP_Q1_DBG): 1 no stores, 2 no patch loop, 4 no gather
  int mass_only;                 // Q1_ELLIP: no stiffness term and a state-independent mass coefficient c2 (the projection test's Q1 mass matrix)
};

// patch layers per slab / elements per block; the census in q1.cu (q1_census_kernel) uses the same slab boundaries
__host__ __device__ constexpr int q1_slab(int P) { return P <= 5 ? P : (P == 7 ? 1 : 2); }      // measured per degree (tools/q1_bench.py)
__host__ __device__ constexpr int q1_epb(int P) { return P == 1 ? 32 : P == 2 ? 16 : P == 3 ? 4 : 1; }

template <int P> struct Q1Cfg {
  static constexpr int N = P + 1, QS = 2 * P, N3 = N * N * N;
  static constexpr int SL = q1_slab(P), EPB = q1_epb(P), QL = 2 * SL;
  static constexpr int NKP = 37;                                  // 36 symmetric entries, odd stride
  static constexpr int THREADS = ((EPB * SL * P * P + 31) / 32) * 32 < 64 ? 64 : ((EPB * SL * P * P + 31) / 32) * 32;
  static constexpr int MINB = 384 / THREADS < 1 ? 1 : 384 / THREADS;           // caps the patch phase at about 168 registers
  static constexpr int NW = THREADS / 32;
  static constexpr int TAB = 2 * QS * N + 2 * QS;                 // Bs, Ds, qs, ws
  // block layout (doubles): tables | per element: 8 vertices, nodal state | per element: x / y contraction stages (dead once
  // the patch phase is over: the gather stages its row sums there) | per element: packed patch matrices of the slab
  static constexpr int XU = 24 + N3, DS = 2 * QL * N * N + 3 * QL * QS * N, KS = SL * P * P * NKP;
  static constexpr int GR = 31, GS = 16 * GR;                     // gather staging per warp: 16 rows x (27 sums + the 32 B row record)
  static constexpr int BX = TAB, BD = BX + EPB * XU, BK = BD + (EPB * DS > NW * GS ? EPB * DS : NW * GS);
  static constexpr int SMEM_DOUBLES = BK + EPB * KS;
};

// upper-triangular packing of the symmetric 8x8 patch matrix
__host__ __device__ constexpr int q1_sym(int r, int c)
{
  const int lo = r < c ? r : c, hi = r < c ? c : r;
  return lo * 8 - lo * (lo - 1) / 2 + (hi - lo);
}
// the eight packed indices of row r as 6-bit fields of one constant (lane-dependent column, compile-time row)
__host__ __device__ constexpr unsigned long long q1_sym_row(int r)
{
  unsigned long long w = 0;
  for (int c = 0; c < 8; c++) w |= (unsigned long long)q1_sym(r, c) << (6 * c);
  return w;
}

// RED.ADD.F64 unless the record byte says the entry is dropped (0xFF): predicated, no branch
DI void q1_red_unless_dropped(double *addr, double v, unsigned rel)
{
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 255;\n\t@p red.global.add.f64 [%0], %1;\n\t}" ::"l"(addr), "d"(v), "r"(rel) : "memory");
}

#ifndef Q1_QUNROLL
#define Q1_QUNROLL 1                      // points of a patch unrolled (8 = immediate patch-basis operands; measured: spills at 168 registers, 1 kept)
#endif
constexpr int Q1_QU = Q1_QUNROLL;

DI bool elem_is_affine(const double *X)
{
  bool a = true;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const double ex = X[(4) * 3 + d] - X[d], ey = X[(2) * 3 + d] - X[d], ez = X[(1) * 3 + d] - X[d];
#pragma unroll
    for (int v = 1; v < 4; v++) {
      // the other three edges of each direction: vertex index bits (i<<2 | j<<1 | k)
      const int bx = ((v & 2) << 0) | (v & 1), by = ((v & 2) << 1) | (v & 1), bz = v << 1;     // (j,k) / (i,k) / (i,j) combinations
      a = a && (X[(bx | 4) * 3 + d] - X[bx * 3 + d] == ex) && (X[(by | 2) * 3 + d] - X[by * 3 + d] == ey) && (X[(bz | 1) * 3 + d] - X[bz * 3 + d] == ez);
    }
  }
  return a;
}

template <int P, bool MASS> __global__ void __launch_bounds__(Q1Cfg<P>::THREADS, Q1Cfg<P>::MINB) q1_ellip_kernel(const Q1Launch A)
{
  using C = Q1Cfg<P>;
  constexpr int N = C::N, QS = C::QS, N3 = C::N3, SL = C::SL, QL = C::QL, EPB = C::EPB, NKP = C::NKP, T = C::THREADS;
  extern __shared__ double sm[];
  const double *Bs = sm, *Ds = Bs + QS * N, *qs = Ds + QS * N, *ws = qs + QS;
  const int tid = threadIdx.x;
  const long e0 = (long)blockIdx.x * EPB;
  for (int i = tid; i < C::TAB; i += T) sm[i] = A.tab[i];
  for (int i = tid; i < EPB * 24; i += T) {
    const int el = i / 24, r = i - el * 24;
    sm[C::BX + el * C::XU + r] = e0 + el < A.nelem ? A.coords[(e0 + el) * 24 + r] : (double)((r / 3 >> (2 - r % 3)) & 1);   // padding elements: unit cube
  }
  for (int i = tid; i < EPB * N3; i += T) {              // dFSGlobalToExpanded(INHOMOGENEOUS), fs.c:543-586
    const int el = i / N3, nd = i - el * N3;
    double v = 0;
    if (e0 + el < A.nelem) {
      const long gnode = (e0 + el) * N3 + nd;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(A.noff + gnode * 32));                    // the row record is read in the gather phase
      const int idx = A.e2l[gnode];
      v = idx < A.n_own ? A.x[idx] : A.xaux[idx - A.n_own];
    }
    sm[C::BX + el * C::XU + 24 + nd] = v;
  }
  __syncthreads();

  const int lane = tid & 31, warp = tid >> 5;

#pragma unroll 1
  for (int s0 = 0; s0 < P; s0 += SL) {
    const int nl = SL < P - s0 ? SL : P - s0;            // patch layers in this slab
    // ---- B: x then y contraction for the 2*nl quadrature layers of the slab -------------------------------------------------
    for (int it = tid; it < EPB * 2 * nl * N * N; it += T) {
      const int el = it / (2 * nl * N * N), r = it - el * (2 * nl * N * N), ql = r / (N * N), jk = r - ql * (N * N), qa = 2 * s0 + ql;
      const double *U = sm + C::BX + el * C::XU + 24;
      double a0 = 0, a1 = 0;
#pragma unroll
      for (int i = 0; i < N; i++) { const double u = U[i * N * N + jk]; a0 = fma(Bs[qa * N + i], u, a0); a1 = fma(Ds[qa * N + i], u, a1); }
      sm[C::BD + el * C::DS + r] = a0;
      sm[C::BD + el * C::DS + QL * N * N + r] = a1;
    }
    __syncthreads();
    for (int it = tid; it < EPB * 2 * nl * QS * N; it += T) {
      const int el = it / (2 * nl * QS * N), r = it - el * (2 * nl * QS * N), ql = r / (QS * N), r2 = r - ql * (QS * N), qb = r2 / N, k = r2 - qb * N;
      const double *X0 = sm + C::BD + el * C::DS + ql * N * N, *X1 = X0 + QL * N * N;
      double v = 0, vx = 0, vy = 0;
#pragma unroll
      for (int j = 0; j < N; j++) {
        const double b = Bs[qb * N + j], x0 = X0[j * N + k];
        v = fma(b, x0, v); vx = fma(b, X1[j * N + k], vx); vy = fma(Ds[qb * N + j], x0, vy);
      }
      double *Vd = sm + C::BD + el * C::DS + 2 * QL * N * N;
      Vd[r] = v;
      Vd[QL * QS * N + r] = vx;
      Vd[2 * QL * QS * N + r] = vy;
    }
    __syncthreads();

    // ---- C: one thread per patch ----------------------------------------------------------------------------------------------
    for (int it = tid; it < EPB * nl * P * P; it += T) {
      const int el = it / (nl * P * P), r = it - el * (nl * P * P), pl = r / (P * P), r2 = r - pl * (P * P), pj = r2 / P, pk = r2 - pj * P;
      if (e0 + el >= A.nelem) continue;
      const int pi = s0 + pl, pt = (pi * P + pj) * P + pk;
      const double *X = sm + C::BX + el * C::XU, *Vd = sm + C::BD + el * C::DS + 2 * QL * N * N;
      const bool aff = elem_is_affine(X);
      PointGeom pgA;
      if (aff) { LineGeom lg; line_geom(X, 0.0, 0.0, lg, false); point_geom_adj(lg, 0.0, pgA); }
      double K[36];
#pragma unroll
      for (int i = 0; i < 36; i++) K[i] = 0;
      if (!(A.dbg & 2))
#pragma unroll Q1_QU
      for (int q = 0; q < 8; q++) {
        constexpr double IA = 0.78867513459481287, IB = 0.21132486540518713;       // interp1, efstopo.c:665
        const int q0 = q >> 2, q1 = (q >> 1) & 1, q2 = q & 1;
        const int ql = 2 * pl + q0, qa = 2 * pi + q0, qb = 2 * pj + q1, qc = 2 * pk + q2;      // patchind, tensorquad.c:144-158
        // state at the point: z contraction
        double u = 0, g0 = 0, g1 = 0, g2 = 0;
        {
          const double *v = Vd + (ql * QS + qb) * N, *vx = v + QL * QS * N, *vy = vx + QL * QS * N;
#pragma unroll
          for (int k = 0; k < N; k++) {
            const double b = Bs[qc * N + k], vv = v[k];
            u = fma(b, vv, u); g0 = fma(b, vx[k], g0); g1 = fma(b, vy[k], g1); g2 = fma(Ds[qc * N + k], vv, g2);
          }
        }
        // geometry at the true point (jacobi.c:383-414) and, quirk (ii), the inverse Jacobian of element-natural point
        // number pt*8+q that the reference maps the patch basis with (efstopo.c:692, fsrulesetit.c:677-683)
        PointGeom pg;
        double jq[9];
        if (aff) {
          pg = pgA;
#pragma unroll
          for (int i = 0; i < 9; i++) jq[i] = pgA.adj[i] * pgA.rdet;
        } else {
          LineGeom lg;
          line_geom(X, qs[qa], qs[qb], lg, false);
          point_geom_adj(lg, qs[qc], pg);
          const int m = pt * 8 + q, mc = m % QS, mb = (m / QS) % QS, ma = m / (QS * QS);
          PointGeom pq;
          line_geom(X, qs[ma], qs[mb], lg, false);
          point_geom_adj(lg, qs[mc], pq);
#pragma unroll
          for (int i = 0; i < 9; i++) jq[i] = pq.adj[i] * pq.rdet;
        }
        const double jw = ws[qa] * ws[qb] * ws[qc] * pg.det;
        double Du[3];
#pragma unroll
        for (int x = 0; x < 3; x++) Du[x] = pg.rdet * (g0 * pg.adj[x] + g1 * pg.adj[3 + x] + g2 * pg.adj[6 + x]);
        // EllipPointwiseComputeStore (ellip.c:333-342) + EllipPointwiseJacobian (ellip.c:369-377) on unit trial functions:
        // T = jw (eta jq jq^T - t t^T), t = jq s, c = jw lambda e^u
        double eta = A.mass_only ? 0.0 : 1.0, sq = 0;
        if (A.c0 != 0) {
          const double gamma = 0.5 * (Du[0] * Du[0] + Du[1] * Du[1] + Du[2] * Du[2]), espg = A.c1 + gamma;
          eta = exp(A.c0 * log(espg));                     // = pow(espg, c0) to a few ulp for the O(1) exponents of the p-Laplacian
          sq = sqrt(-A.c0 * eta / espg);
        }
        const double s0_ = sq * Du[0], s1_ = sq * Du[1], s2_ = sq * Du[2];
        const double t0 = jq[0] * s0_ + jq[1] * s1_ + jq[2] * s2_, t1 = jq[3] * s0_ + jq[4] * s1_ + jq[5] * s2_, t2 = jq[6] * s0_ + jq[7] * s1_ + jq[8] * s2_;
        const double etaw = eta * jw;
        const double T00 = etaw * (jq[0] * jq[0] + jq[1] * jq[1] + jq[2] * jq[2]) - jw * t0 * t0;
        const double T11 = etaw * (jq[3] * jq[3] + jq[4] * jq[4] + jq[5] * jq[5]) - jw * t1 * t1;
        const double T22 = etaw * (jq[6] * jq[6] + jq[7] * jq[7] + jq[8] * jq[8]) - jw * t2 * t2;
        const double T01 = etaw * (jq[0] * jq[3] + jq[1] * jq[4] + jq[2] * jq[5]) - jw * t0 * t1;
        const double T02 = etaw * (jq[0] * jq[6] + jq[1] * jq[7] + jq[2] * jq[8]) - jw * t0 * t2;
        const double T12 = etaw * (jq[3] * jq[6] + jq[4] * jq[7] + jq[5] * jq[8]) - jw * t1 * t2;
        const double cq = MASS ? jw * A.c2 * (A.mass_only ? 1.0 : exp(u)) : 0.0;
        // reference-space patch basis at this point: interp1[q][b] = IA if b is on the point's side else IB, deriv1 = -1,+1
        double yz[4], xz[4], xy[4], xyz[8];
        {
          const double ax[2] = {q0 ? IB : IA, q0 ? IA : IB}, ay[2] = {q1 ? IB : IA, q1 ? IA : IB}, az[2] = {q2 ? IB : IA, q2 ? IA : IB};
#pragma unroll
          for (int m = 0; m < 4; m++) { yz[m] = ay[m >> 1] * az[m & 1]; xz[m] = ax[m >> 1] * az[m & 1]; xy[m] = ax[m >> 1] * ay[m & 1]; }
#pragma unroll
          for (int m = 0; m < 8; m++) xyz[m] = xy[m >> 1] * az[m & 1];
        }
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const int j0 = j >> 2, j1 = (j >> 1) & 1, j2 = j & 1;
          const double dj0 = j0 ? yz[j & 3] : -yz[j & 3], dj1 = j1 ? xz[j0 * 2 + j2] : -xz[j0 * 2 + j2], dj2 = j2 ? xy[j >> 1] : -xy[j >> 1];
          const double w0 = T00 * dj0 + T01 * dj1 + T02 * dj2, w1 = T01 * dj0 + T11 * dj1 + T12 * dj2, w2 = T02 * dj0 + T12 * dj1 + T22 * dj2;
          const double cphi = cq * xyz[j];
#pragma unroll
          for (int i = 0; i <= j; i++) {
            const int i0 = i >> 2, i1 = (i >> 1) & 1, i2 = i & 1;
            const double di0 = i0 ? yz[i & 3] : -yz[i & 3], di1 = i1 ? xz[i0 * 2 + i2] : -xz[i0 * 2 + i2], di2 = i2 ? xy[i >> 1] : -xy[i >> 1];
            double acc = K[q1_sym(i, j)];
            acc = fma(di0, w0, acc); acc = fma(di1, w1, acc); acc = fma(di2, w2, acc);
            if (MASS) acc = fma(-xyz[i], cphi, acc);
            K[q1_sym(i, j)] = acc;
          }
        }
      }
      double *Kp = sm + C::BK + el * C::KS + r * NKP;
#pragma unroll
      for (int i = 0; i < 36; i++) Kp[i] = K[i];
    }
    __syncthreads();

    // ---- G: rows of node layers s0 .. s0+nl (dFSMatSetValuesBlockedExpanded, fs.c:826-930) ------------------------------------
    // G1: one thread per row node adds the entries of the <= 8 patches around it into the 27 neighbour slots (registers).
    // G2: the sums go through a per-warp staging area so that 8 consecutive lanes write 8 consecutive CSR entries of one row.
    double *S = sm + C::BD + warp * C::GS;
    for (int base = warp * 32; base < ((A.dbg & 4) ? 0 : EPB * (SL + 1) * N * N); base += C::NW * 32) {
      const int idx = base + lane;
      const int el = idx / ((SL + 1) * N * N), r = idx - el * ((SL + 1) * N * N), il = r / (N * N), r2 = r - il * (N * N), j = r2 / N, k = r2 - j * N;
      uint4 lo = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu), hi = lo;      // no row: rp = -1
      if (idx < EPB * (SL + 1) * N * N && il <= nl && e0 + el < A.nelem) {
        const long gnode = (e0 + el) * N3 + ((s0 + il) * N + j) * N + k;
        lo = reinterpret_cast<const uint4 *>(A.noff)[gnode * 2];
        hi = reinterpret_cast<const uint4 *>(A.noff)[gnode * 2 + 1];
      }
      const int rp = (int)hi.w;                            // CSR start of the node's row, -1 for Dirichlet nodes (fsbuild.c:31-34)
      double sums[27];
#pragma unroll
      for (int t = 0; t < 27; t++) sums[t] = 0;
      if (rp >= 0) {
        const double *Kel = sm + C::BK + el * C::KS + ((il * P + j) * P + k) * NKP;
#pragma unroll
        for (int abc = 0; abc < 8; abc++) {
          const int a = abc >> 2, b = (abc >> 1) & 1, c = abc & 1;          // the node is corner (a,b,c) of patch (il-a, j-b, k-c)
          if ((unsigned)(il - a) < (unsigned)nl && (unsigned)(j - b) < (unsigned)P && (unsigned)(k - c) < (unsigned)P) {
            const double *Kp = Kel - ((a * P + b) * P + c) * NKP;
#pragma unroll
            for (int col = 0; col < 8; col++) {
              const int ca = col >> 2, cb = (col >> 1) & 1, cc = col & 1;
              sums[(ca - a + 1) * 9 + (cb - b + 1) * 3 + (cc - c + 1)] += Kp[q1_sym(abc, col)];
            }
          }
        }
      }
#pragma unroll
      for (int h = 0; h < 2; h++) {
        if ((lane >> 4) == h) {
          double *row = S + (lane & 15) * C::GR;
#pragma unroll
          for (int t = 0; t < 27; t++) row[t] = sums[t];
          unsigned long long *r64 = reinterpret_cast<unsigned long long *>(row + 27);      // the staging rows are only 8-byte aligned
          r64[0] = lo.x | (unsigned long long)lo.y << 32; r64[1] = lo.z | (unsigned long long)lo.w << 32;
          r64[2] = hi.x | (unsigned long long)hi.y << 32; r64[3] = hi.z | (unsigned long long)hi.w << 32;
        }
        __syncwarp();
#pragma unroll
        for (int itr = 0; itr < 4; itr++) {
          const int ml = 4 * itr + (lane >> 3), sub = lane & 7;
          const double *row = S + ml * C::GR;
          const unsigned char *rec = reinterpret_cast<const unsigned char *>(row + 27);
          const int rpm = *reinterpret_cast<const int *>(rec + 28);
          int ilm = 0;
          if (SL < P) { const int im = base + 16 * h + ml; ilm = (im % ((SL + 1) * N * N)) / (N * N); }
#pragma unroll
          for (int sI = 0; sI < 4; sI++) {
            const int t = sub + 8 * sI;
            unsigned rel = t < 27 ? rec[t] : 0xFFu;
            if ((A.dbg & 1) || rpm < 0 || (SL < P && (unsigned)(ilm + t / 9 - 1) > (unsigned)nl)) rel = 0xFFu;
            q1_red_unless_dropped(A.vals + (rpm + (long)rel), row[t], rel);    // RED.ADD.F64: 8 consecutive lanes, 8 consecutive entries
          }
        }
        __syncwarp();
      }
    }
    __syncthreads();
  }
}

typedef int (*Q1Fn)(const Q1Launch &, cudaStream_t);
#define DOHP_Q1_DECL(N) int launch_q1_##N(const Q1Launch &, cudaStream_t);
DOHP_Q1_DECL(2) DOHP_Q1_DECL(3) DOHP_Q1_DECL(4) DOHP_Q1_DECL(5) DOHP_Q1_DECL(6) DOHP_Q1_DECL(7) DOHP_Q1_DECL(8) DOHP_Q1_DECL(9) DOHP_Q1_DECL(10) DOHP_Q1_DECL(11)
#undef DOHP_Q1_DECL

}  // namespace dohp
