// Internal declarations shared by the host translation units of libdohp_b200.so.
#pragma once
#include "../../include/dohp_b200.h"
#include <cstddef>
#include <array>
#include <map>
#include <string>
#include <vector>

#ifdef __CUDACC__
#include <cuda_runtime.h>
#else
typedef struct CUstream_st *cudaStream_t;
#endif

namespace dohp {

int seterr(int code, const char *fmt, ...) __attribute__((format(printf, 2, 3)));

void gauss_legendre(int n, double *z, double *w);
void gauss_lobatto(int n, double *z, double *w);
void collocation_derivative(int P, const double *z, double *D);
int rule_1d(int method, int family, int order, int *size, double *coord, double *weight);
int basis_1d(int P, int Q, const double *qcoord, double *interp, double *deriv, double *node, double *weight);
int build_smem_maps(int nq, int sa, int sb, int epb, int threads, int es_base, int *es_out, std::vector<int> &maps);

}  // namespace dohp

#define DB200_MAXFIELDS 2

struct dB200Ctx_s {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  long long nlaunch = 0;
  int rank = 0, nranks = 1;
  void *nccl = nullptr;                 // ncclComm_t
  cudaStream_t h2d = nullptr, d2h = nullptr;   // copy streams of the asynchronous host-pointer calls (dB200OpMultAsync)
  cudaStream_t comm = nullptr;          // ghost exchanges of partitioned operators run here, overlapped with the interior elements
  void *ev[4] = {nullptr};              // cudaEvent_t: vectors ready / ghosts arrived / interface elements done / owners updated
  void *scratch[8] = {nullptr};         // growable device staging buffers
  std::map<std::array<int, 6>, std::pair<int, int *>> smem_maps;   // (nq, sa, sb, epb, threads, es_base) -> (element stride, device thread maps), smem_map.cpp
  size_t scratch_bytes[8] = {0};
};

struct dB200Rule_s {
  dB200Ctx ctx;
  int Q[3];
  int dim = 3;                          // tensor directions (1: dTOPO_LINE, 2: dTOPO_QUAD, 3: dTOPO_HEX); the unused ones have Q = 1
  std::vector<double> coord[3], weight[3], Dt[3];     // Dt = collocation derivative on the rule's points
};

struct dB200EFS_s {
  dB200Ctx ctx;
  dB200Rule rule;
  int P[3];
  int dim = 3;
  std::vector<double> interp[3], deriv[3];
  bool colloc_ok = true;                // deriv == (collocation derivative on the rule points) . interp in every direction: the fused kernels apply
  double *d_gtab = nullptr;             // device copy of the explicit tables (apply_generic.cu), created on first use
};

struct dB200Space_s {
  dB200Ctx ctx;
  int nelem, P, bs, n, nc, ngh;
  int *d_e2l = nullptr;
  double *d_xaux = nullptr, *d_xaux_inhomog = nullptr, *d_yaux = nullptr;   // [(nc+ngh-n)*bs]
  // halo
  int nneigh = 0;
  std::vector<int> ranks, sendcount, recvcount, sendoff, recvoff;
  int *d_sendidx = nullptr, *d_recvidx = nullptr;       // concatenated, local node indices
  double *d_sendbuf = nullptr, *d_recvbuf = nullptr;
  int nsend = 0, nrecv = 0;
  void *p2p = nullptr;                  // p2p.cu: neighbours' buffers mapped over NVLink (NULL: the NCCL path of comm.cu)
};

struct dB200Op_s {
  dB200Ctx ctx;
  int kind;
  dB200Rule rule;
  int nspaces;
  dB200Space space[DB200_MAXFIELDS];
  dB200EFS efs[DB200_MAXFIELDS];
  int nelem, NQ;
  double *d_coords = nullptr;
  double *d_stash = nullptr;
  int nstash = 0;
  double *d_forcing = nullptr;
  int nforcing = 0;
  bool stash_valid = false;
  bool affine = false;                  // every element is a parallelepiped (checked at creation)
  int *d_eorder = nullptr;              // partitioned spaces: elements touching ghost nodes first (n_bnd of them), then the interior ones
  int n_bnd = 0;
  dB200EllipParam ellip;
  dB200ElastParam elast;
  dB200StokesParam stokes;
  double *d_gx[DB200_MAXFIELDS] = {nullptr}, *d_gy[DB200_MAXFIELDS] = {nullptr};   // per-operator staging of dB200OpMultAsync
  void *ev_in = nullptr, *ev_out = nullptr, *ev_done = nullptr;                      // cudaEvent_t: upload done / kernels done / download done
  void *q1cache = nullptr;              // q1.cu: cached block-CSR pattern
  int q1reduced = 0;                    // -dfs_assemble_reduced
  double *d_q1tab = nullptr;            // q1.cu: SPARSE-rule tables
  void *densecache[DB200_MAXFIELDS] = {nullptr};      // dense.cu: element-dense CSR pattern per field
  void *q1owner = nullptr;              // q1_owner.cu: merged pattern + position maps of the owner-side reduction
};

// api.cu: conflict-free shared-memory scheme of one kernel configuration (cached per context)
int dohp_smem_maps(dB200Ctx c, int nq, int sa, int sb, int epb, int threads, int es_base, int *es, const int **d_maps);
// comm.cu
int dohp_exchange(dB200Ctx c, dB200Space s, bool reverse, cudaStream_t stream);
void dohp_comm_destroy(dB200Ctx c);
// api.cu: forward INSERT of the owners' values into the ghost part of aux
int dohp_halo_forward(dB200Space s, const double *x, double *aux);
int dohp_halo_reverse(dB200Space s, const double *aux, double *y);     // reverse ADD of the ghost part of aux into the owners' y
// p2p.cu: ghost exchange by direct stores into the neighbours' memory
int dohp_p2p_setup(dB200Space s);
void dohp_p2p_free(dB200Space s);
int dohp_p2p_forward(dB200Space s, const double *x, int which, cudaStream_t st);
int dohp_p2p_reverse(dB200Space s, const double *aux, double *y, cudaStream_t st);
// q1.cu
void dohp_q1_free(dB200Op o);
// dense.cu
void dohp_dense_free(dB200Op o);
// q1_owner.cu
void dohp_q1owner_free(dB200Op o);
