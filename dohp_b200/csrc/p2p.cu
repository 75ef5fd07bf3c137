// Ghost exchange over NVLink peer memory, without NCCL on the hot path: the VecGhostUpdate pair of dFSGlobalToExpanded /
// dFSExpandedToGlobal (src/fs/interface/fs.c:578-579 forward INSERT, :627-628 reverse ADD) as direct stores into the neighbours' HBM.
//
// Setup (once per space, inside dB200SpaceSetHalo; collective): every rank exports its ghost buffers (the two `aux` arrays the
// kernels gather ghost values from, the buffer reverse contributions land in, a flag array) with cudaIpcGetMemHandle, the handles and
// the slot lists travel to the neighbours over NCCL, cudaIpcOpenMemHandle maps the neighbours' buffers.  Per exchange, on the
// caller's stream, three small launches and no host synchronisation:
//   sync(READY)   every rank tells its neighbours that its stream has reached this exchange (so their buffers may be overwritten:
//                 everything that read the previous contents has completed) and waits for the same word from each of them;
//   push          forward: x[sendidx[j]] -> neighbour's aux[slot[j]] (pack and transfer in one kernel, no unpack on the other side);
//                 reverse: my ghost contributions -> the owner's landing buffer, segment by segment;
//   sync(DONE)    flags "my stores are complete" (kernel boundary + __threadfence_system) and waits for the neighbours' flags;
// the reverse leg ends with the owner's local scatter-add.  Flags are monotone epochs (never reset), one word per (kind, rank).
// Against the grouped ncclSend/ncclRecv of comm.cu this removes the pack -> NCCL kernel -> unpack chain (about 40 us of latency per
// exchange at low degree, about 110 us for the 10 MB halo of p = 10 on 8 GPUs): 8 GPUs, 64^3 elements each, p = 2..10 sweep 449.9 (NCCL) ->
// 464.8 GDoF/s, p = 2 MatMult 0.214 -> 0.167 ms, p = 10 4.106 -> 3.980 ms (profiles/r2y_bench_n8_*.json).
// Falls back to the NCCL path when the peers cannot be mapped (other node, no peer access) or DOHP_B200_P2P=0.
#include "internal.h"
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace dohp;

#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return seterr(700 + (int)e_, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); } while (0)
#define CHK(call) do { int e_ = (call); if (e_) return e_; } while (0)
#define LAUNCHED(ctx) ((ctx)->nlaunch++)

int dohp_exchange_raw(dB200Ctx c, dB200Space s, int elem_bytes, const void *src, const std::vector<long long> &so, void *dst, const std::vector<long long> &ro, cudaStream_t stream);
int dohp_allreduce_min_int(dB200Ctx c, int *d_val, cudaStream_t stream);

enum { P2P_MAXN = 26, P2P_READY = 0, P2P_FWD = 1, P2P_REV = 2, P2P_KINDS = 3 };

struct P2PTab {                    // kernel parameter: the neighbours of one space
  int nn, myrank, nranks;
  int rank[P2P_MAXN];
  unsigned *peer_flags[P2P_MAXN];  // neighbour's flag array [P2P_KINDS][nranks]
  double *dst[P2P_MAXN];           // destination buffer at the neighbour for the current push
  int off[P2P_MAXN + 1];           // entry ranges of the neighbours in the list being pushed
  int dst_off[P2P_MAXN];           // reverse: first entry of my segment in the neighbour's landing buffer
};

struct P2PHalo {
  P2PTab tab;
  double *peer_xaux[P2P_MAXN], *peer_xaux_in[P2P_MAXN], *peer_land[P2P_MAXN];
  void *mapped[P2P_MAXN][4];
  unsigned *d_flags = nullptr;     // mine
  int *d_fwd_slot = nullptr;       // [nsend] aux node offset at the neighbour for every forward entry
  int *h_err = nullptr, *d_err = nullptr;      // pinned, mapped: set by a kernel whose wait timed out
  unsigned epoch[P2P_KINDS] = {0, 0, 0};
};

// signal `epoch` of `kind` to every neighbour, then wait for theirs
__global__ void p2p_sync_kernel(const P2PTab T, int kind, unsigned epoch, unsigned *myflags, int *err)
{
  const int t = threadIdx.x;
  if (t >= T.nn) return;
  __threadfence_system();
  *(volatile unsigned *)(T.peer_flags[t] + kind * T.nranks + T.myrank) = epoch;
  const volatile unsigned *f = myflags + kind * T.nranks + T.rank[t];
  const long long t0 = clock64();
  while ((int)(*f - epoch) < 0) {
    if (clock64() - t0 > 600000000000LL) { *err = 1 + kind; break; }     // about 5 min: a lost neighbour must not hang the GPU for ever
    __nanosleep(200);
  }
  __threadfence_system();
}
// forward: pack and transfer in one kernel
__global__ void p2p_push_fwd_kernel(const P2PTab T, long n, int bs, const int *sendidx, const int *slot, const double *x)
{
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n * bs) return;
  const int j = (int)(g / bs), b = (int)(g - (long)j * bs);
  int r = 0;
  while (j >= T.off[r + 1]) r++;
  T.dst[r][(long)slot[j] * bs + b] = x[(long)sendidx[j] * bs + b];
}
// reverse: my ghost contributions (aux index = local - n) -> the owner's landing buffer
__global__ void p2p_push_rev_kernel(const P2PTab T, long n, int bs, const int *recvidx, int nown, const double *aux)
{
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n * bs) return;
  const int j = (int)(g / bs), b = (int)(g - (long)j * bs);
  int r = 0;
  while (j >= T.off[r + 1]) r++;
  T.dst[r][(long)(T.dst_off[r] + (j - T.off[r])) * bs + b] = aux[(long)(recvidx[j] - nown) * bs + b];
}
__global__ void p2p_scatter_add_kernel(long n, int bs, const int *idx, const double *src, double *dst)
{
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n * bs) return;
  atomicAdd(dst + (long)idx[g / bs] * bs + g % bs, src[g]);
}

void dohp_p2p_free(dB200Space s)
{
  P2PHalo *h = (P2PHalo *)s->p2p;
  if (!h) return;
  for (int r = 0; r < h->tab.nn; r++)
    for (int k = 0; k < 4; k++) if (h->mapped[r][k]) cudaIpcCloseMemHandle(h->mapped[r][k]);
  cudaFree(h->d_flags); cudaFree(h->d_fwd_slot);
  if (h->h_err) cudaFreeHost(h->h_err);
  delete h;
  s->p2p = nullptr;
}

struct P2PBlob { cudaIpcMemHandle_t h[4]; int land_off; int ok; };

// collective over the ranks of the communicator (every rank calls dB200SpaceSetHalo for the corresponding space)
int dohp_p2p_setup(dB200Space s)
{
  dB200Ctx c = s->ctx;
  const char *env = getenv("DOHP_B200_P2P");
  if (env && atoi(env) == 0) return 0;
  if (!c->nccl || s->nneigh == 0 || s->nneigh > P2P_MAXN) return 0;
  const int nn = s->nneigh;
  const size_t naux = std::max<size_t>((size_t)(s->nc + s->ngh - s->n) * s->bs, 1);
  if (!s->d_xaux_inhomog) {
    CU(cudaMalloc((void **)&s->d_xaux_inhomog, naux * sizeof(double)));
    CU(cudaMemsetAsync(s->d_xaux_inhomog, 0, naux * sizeof(double), c->stream));
  }
  P2PHalo *h = new P2PHalo;
  s->p2p = h;
  memset(&h->tab, 0, sizeof h->tab);
  memset(h->mapped, 0, sizeof h->mapped);
  h->tab.nn = nn; h->tab.myrank = c->rank; h->tab.nranks = c->nranks;
  for (int r = 0; r < nn; r++) h->tab.rank[r] = s->ranks[r];
  CU(cudaMalloc((void **)&h->d_flags, (size_t)P2P_KINDS * c->nranks * sizeof(unsigned)));
  CU(cudaMemsetAsync(h->d_flags, 0, (size_t)P2P_KINDS * c->nranks * sizeof(unsigned), c->stream));
  CU(cudaHostAlloc((void **)&h->h_err, sizeof(int), cudaHostAllocMapped));
  *h->h_err = 0;
  CU(cudaHostGetDevicePointer((void **)&h->d_err, h->h_err, 0));
  // ---- handles + landing offsets ----------------------------------------------------------------------------------------------------
  std::vector<P2PBlob> out(nn), in(nn);
  int ok = 1;
  cudaIpcMemHandle_t mine[4];
  void *bufs[4] = {s->d_xaux, s->d_xaux_inhomog, s->d_sendbuf, h->d_flags};
  for (int k = 0; k < 4; k++) if (cudaIpcGetMemHandle(&mine[k], bufs[k]) != cudaSuccess) { cudaGetLastError(); ok = 0; }
  for (int r = 0; r < nn; r++) { memcpy(out[r].h, mine, sizeof mine); out[r].land_off = s->sendoff[r]; out[r].ok = ok; }
  const int words = (int)(sizeof(P2PBlob) / 4);
  std::vector<long long> bo(nn + 1);
  for (int r = 0; r <= nn; r++) bo[r] = (long long)r * words;
  P2PBlob *d_out = nullptr, *d_in = nullptr;
  CU(cudaMalloc((void **)&d_out, nn * sizeof(P2PBlob))); CU(cudaMalloc((void **)&d_in, nn * sizeof(P2PBlob)));
  CU(cudaMemcpyAsync(d_out, out.data(), nn * sizeof(P2PBlob), cudaMemcpyHostToDevice, c->stream));
  int err = dohp_exchange_raw(c, s, 4, d_out, bo, d_in, bo, c->stream);
  if (!err) { CU(cudaMemcpyAsync(in.data(), d_in, nn * sizeof(P2PBlob), cudaMemcpyDeviceToHost, c->stream)); CU(cudaStreamSynchronize(c->stream)); }
  cudaFree(d_out); cudaFree(d_in);
  if (err) { dohp_p2p_free(s); return err; }
  // ---- slot lists: where each of my forward entries goes in the neighbour's aux = the neighbour's receive list for me -----------------
  std::vector<int> ridx((size_t)s->nrecv), slot((size_t)s->nsend);
  if (s->nrecv) CU(cudaMemcpy(ridx.data(), s->d_recvidx, ridx.size() * sizeof(int), cudaMemcpyDeviceToHost));
  for (auto &v : ridx) v -= s->n;
  std::vector<long long> so(nn + 1), ro(nn + 1);
  for (int r = 0; r <= nn; r++) { so[r] = s->recvoff[r]; ro[r] = s->sendoff[r]; }
  int *d_r = nullptr;
  CU(cudaMalloc((void **)&d_r, std::max<size_t>(ridx.size(), 1) * sizeof(int)));
  CU(cudaMalloc((void **)&h->d_fwd_slot, std::max<size_t>(slot.size(), 1) * sizeof(int)));
  if (!ridx.empty()) CU(cudaMemcpyAsync(d_r, ridx.data(), ridx.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  err = dohp_exchange_raw(c, s, 4, d_r, so, h->d_fwd_slot, ro, c->stream);
  CU(cudaStreamSynchronize(c->stream));
  cudaFree(d_r);
  if (err) { dohp_p2p_free(s); return err; }
  // ---- map the neighbours' buffers ---------------------------------------------------------------------------------------------------------
  for (int r = 0; r < nn && ok; r++) {
    if (!in[r].ok) { ok = 0; break; }
    for (int k = 0; k < 4; k++) {
      if (cudaIpcOpenMemHandle(&h->mapped[r][k], in[r].h[k], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); h->mapped[r][k] = nullptr; ok = 0; break; }
    }
    h->peer_xaux[r] = (double *)h->mapped[r][0]; h->peer_xaux_in[r] = (double *)h->mapped[r][1]; h->peer_land[r] = (double *)h->mapped[r][2];
    h->tab.peer_flags[r] = (unsigned *)h->mapped[r][3];
    h->tab.dst_off[r] = in[r].land_off;
  }
  // every rank of the communicator must take the same path (a rank on NCCL and its neighbour on peer memory would wait forever)
  {
    int *d_a = nullptr;
    CU(cudaMalloc((void **)&d_a, sizeof(int)));
    CU(cudaMemcpyAsync(d_a, &ok, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    err = dohp_allreduce_min_int(c, d_a, c->stream);
    if (!err) { CU(cudaMemcpyAsync(&ok, d_a, sizeof(int), cudaMemcpyDeviceToHost, c->stream)); CU(cudaStreamSynchronize(c->stream)); }
    cudaFree(d_a);
    if (err) { dohp_p2p_free(s); return err; }
  }
  if (!ok) dohp_p2p_free(s);
  return 0;
}

static int p2p_check(dB200Space s)
{
  P2PHalo *h = (P2PHalo *)s->p2p;
  if (*(volatile int *)h->h_err) return seterr(74, "peer-memory ghost exchange: a neighbour did not arrive within 5 minutes (kind %d)", *h->h_err - 1);
  return 0;
}

// forward INSERT into the neighbours' aux; `which` = 0: d_xaux, 1: d_xaux_inhomog (the target every rank passes)
int dohp_p2p_forward(dB200Space s, const double *x, int which, cudaStream_t st)
{
  dB200Ctx c = s->ctx;
  P2PHalo *h = (P2PHalo *)s->p2p;
  CHK(p2p_check(s));
  P2PTab T = h->tab;
  for (int r = 0; r < T.nn; r++) { T.dst[r] = which ? h->peer_xaux_in[r] : h->peer_xaux[r]; T.off[r] = s->sendoff[r]; }
  T.off[T.nn] = s->sendoff[T.nn];
  p2p_sync_kernel<<<1, 32, 0, st>>>(T, P2P_READY, ++h->epoch[P2P_READY], h->d_flags, h->d_err); LAUNCHED(c);
  if (s->nsend) {
    const long m = (long)s->nsend * s->bs;
    p2p_push_fwd_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(T, s->nsend, s->bs, s->d_sendidx, h->d_fwd_slot, x); LAUNCHED(c);
  }
  p2p_sync_kernel<<<1, 32, 0, st>>>(T, P2P_FWD, ++h->epoch[P2P_FWD], h->d_flags, h->d_err); LAUNCHED(c);
  CU(cudaGetLastError());
  return 0;
}

// reverse ADD: ghost part of `aux` -> the owners' landing buffers -> local scatter-add into y
int dohp_p2p_reverse(dB200Space s, const double *aux, double *y, cudaStream_t st)
{
  dB200Ctx c = s->ctx;
  P2PHalo *h = (P2PHalo *)s->p2p;
  CHK(p2p_check(s));
  P2PTab T = h->tab;
  for (int r = 0; r < T.nn; r++) { T.dst[r] = h->peer_land[r]; T.off[r] = s->recvoff[r]; }
  T.off[T.nn] = s->recvoff[T.nn];
  p2p_sync_kernel<<<1, 32, 0, st>>>(T, P2P_READY, ++h->epoch[P2P_READY], h->d_flags, h->d_err); LAUNCHED(c);
  if (s->nrecv) {
    const long m = (long)s->nrecv * s->bs;
    p2p_push_rev_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(T, s->nrecv, s->bs, s->d_recvidx, s->n, aux); LAUNCHED(c);
  }
  p2p_sync_kernel<<<1, 32, 0, st>>>(T, P2P_REV, ++h->epoch[P2P_REV], h->d_flags, h->d_err); LAUNCHED(c);
  if (s->nsend) {
    const long m = (long)s->nsend * s->bs;
    p2p_scatter_add_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(s->nsend, s->bs, s->d_sendidx, s->d_sendbuf, y); LAUNCHED(c);
  }
  CU(cudaGetLastError());
  return 0;
}
