// Owner-side reduction of the embedded-Q1 rows a rank assembled for nodes it does not own: the MatAssemblyBegin/End pair of the
// reference (src/fs/tests/ellip.c:528-529, elast.c:401-402) after dFSMatSetValuesBlockedExpanded stashed the off-process rows
// (MatSetValuesBlockedLocal, src/fs/interface/fs.c:919).  dB200OpAssembleQ1 leaves those contributions in the rows [n, n+ngh) of the
// local block CSR; here every such row travels to the rank that owns its node (the halo lists of the space run backwards, as in the
// reverse ghost update fs.c:627-628) and is ADDed into the owner's row.  An owner's final row also couples to nodes that only its
// neighbours' elements touch, so the pattern is reduced first (once): OwnerPattern exchanges the ghost rows' column lists over NCCL,
// merges them into the owned rows and keeps position maps; OwnerReduce is then pack -> grouped ncclSend/ncclRecv -> one add kernel per
// neighbour in neighbour order (deterministic sums).  Result: the n owned (block) rows with global columns, exactly the rows of the
// global matrix this rank holds after MatAssemblyEnd.
#include "internal.h"
#include <algorithm>
#include <vector>

using namespace dohp;

#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return seterr(700 + (int)e_, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); } while (0)
#define CHK(call) do { int e_ = (call); if (e_) return e_; } while (0)
#define LAUNCHED(ctx) ((ctx)->nlaunch++)

int dohp_exchange_raw(dB200Ctx c, dB200Space s, int elem_bytes, const void *src, const std::vector<long long> &so, void *dst, const std::vector<long long> &ro, cudaStream_t stream);

struct Q1Owner {
  int n = 0, bs2 = 1;
  long long nnz_own = 0, onnz = 0, nsend = 0, nrecv = 0;
  std::vector<int> orowptr, ocolidx;                 // merged pattern of the owned rows (host copy for the caller)
  std::vector<long long> soff, roff;                  // per-neighbour entry offsets (send = my ghost rows, recv = rows I own)
  int *d_own_pos = nullptr, *d_send_src = nullptr, *d_recv_pos = nullptr;
  double *d_sendbuf = nullptr, *d_recvbuf = nullptr;
};

void dohp_q1owner_free(dB200Op o)
{
  Q1Owner *w = (Q1Owner *)o->q1owner;
  if (!w) return;
  cudaFree(w->d_own_pos); cudaFree(w->d_send_src); cudaFree(w->d_recv_pos); cudaFree(w->d_sendbuf); cudaFree(w->d_recvbuf);
  delete w;
  o->q1owner = nullptr;
}

__global__ void q1own_scatter_kernel(long long nent, int bs2, const int *pos, const double *src, double *dst, int add)
{
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nent * bs2) return;
  const long long k = g / bs2;
  const int t = (int)(g - k * bs2);
  double *d = dst + (long long)pos[k] * bs2 + t;
  *d = add ? *d + src[g] : src[g];
}
__global__ void q1own_pack_kernel(long long nent, int bs2, const int *srcidx, const double *src, double *dst)
{
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nent * bs2) return;
  const long long k = g / bs2;
  dst[g] = src[(long long)srcidx[k] * bs2 + (g - k * bs2)];
}

template <class T> static int upload(const std::vector<T> &h, T **d, cudaStream_t s)
{
  CU(cudaMalloc((void **)d, std::max<size_t>(h.size(), 1) * sizeof(T)));
  if (!h.empty()) CU(cudaMemcpyAsync(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s));
  return 0;
}

extern "C" int dB200OpQ1OwnerPattern(dB200Op o, const int *rowptr, const int *colidx, long long *nnz, int *orowptr, int *ocolidx)
{
  dB200Ctx c = o->ctx;
  CU(cudaSetDevice(c->device));
  dB200Space s = o->space[0];
  Q1Owner *w = (Q1Owner *)o->q1owner;
  if (!w) {
    if (!rowptr || !colidx) return seterr(85, "Q1 owner pattern: the local pattern (dB200OpQ1Pattern) is needed on the first call");
    const int n = s->n, ngh = s->ngh, nrows = n + ngh;
    std::vector<int> rp((size_t)nrows + 1);
    CU(cudaMemcpy(rp.data(), rowptr, rp.size() * sizeof(int), cudaMemcpyDefault));
    std::vector<int> ci((size_t)rp[nrows]);
    if (!ci.empty()) CU(cudaMemcpy(ci.data(), colidx, ci.size() * sizeof(int), cudaMemcpyDefault));
    // halo lists: forward send = my owned nodes the neighbour ghosts, forward recv = my ghost nodes (local index nc + g)
    std::vector<int> sidx((size_t)s->nsend), ridx((size_t)s->nrecv);
    if (s->nsend) CU(cudaMemcpy(sidx.data(), s->d_sendidx, sidx.size() * sizeof(int), cudaMemcpyDeviceToHost));
    if (s->nrecv) CU(cudaMemcpy(ridx.data(), s->d_recvidx, ridx.size() * sizeof(int), cudaMemcpyDeviceToHost));
    const int nn = s->nneigh;
    w = new Q1Owner;
    w->n = n; w->bs2 = s->bs * s->bs; w->nnz_own = rp[n];
    // ---- 1. row lengths of my ghost rows -> owners ---------------------------------------------------------------------------
    std::vector<long long> lso(nn + 1, 0), lro(nn + 1, 0);         // the reduction runs backwards: I send per forward-recv entry
    for (int r = 0; r < nn; r++) { lso[r + 1] = lso[r] + s->recvcount[r]; lro[r + 1] = lro[r] + s->sendcount[r]; }
    std::vector<int> slen((size_t)lso[nn]), rlen((size_t)lro[nn]);
    w->soff.assign(nn + 1, 0);
    std::vector<int> send_src, scol;
    for (int r = 0; r < nn; r++) {
      for (long long j = lso[r]; j < lso[r + 1]; j++) {
        const int row = n + (ridx[(size_t)j] - s->nc);
        if (row < n || row >= nrows) { delete w; return seterr(63, "halo receive index %d is not a ghost node", ridx[(size_t)j]); }
        slen[(size_t)j] = rp[row + 1] - rp[row];
        for (int k = rp[row]; k < rp[row + 1]; k++) { send_src.push_back(k); scol.push_back(ci[k]); }
      }
      w->soff[r + 1] = (long long)send_src.size();
    }
    w->nsend = (long long)send_src.size();
    int *d_slen = nullptr, *d_rlen = nullptr, *d_scol = nullptr, *d_rcol = nullptr;
    struct Tmp { std::vector<void *> p; ~Tmp() { for (void *q : p) cudaFree(q); } } tmp;
    if (nn && c->nccl) {
      CHK(upload(slen, &d_slen, c->stream)); tmp.p.push_back(d_slen);
      CU(cudaMalloc((void **)&d_rlen, std::max<size_t>(rlen.size(), 1) * sizeof(int))); tmp.p.push_back(d_rlen);
      int err = dohp_exchange_raw(c, s, 4, d_slen, lso, d_rlen, lro, c->stream);
      if (err) { delete w; return err; }
      if (!rlen.empty()) CU(cudaMemcpyAsync(rlen.data(), d_rlen, rlen.size() * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
      CU(cudaStreamSynchronize(c->stream));
    } else if (nn) { delete w; return seterr(73, "Q1 owner pattern: the space has neighbours but the context has no communicator"); }
    // ---- 2. column lists ------------------------------------------------------------------------------------------------------
    w->roff.assign(nn + 1, 0);
    for (int r = 0; r < nn; r++) {
      long long t = 0;
      for (long long j = lro[r]; j < lro[r + 1]; j++) t += rlen[(size_t)j];
      w->roff[r + 1] = w->roff[r] + t;
    }
    w->nrecv = w->roff[nn];
    std::vector<int> rcol((size_t)w->nrecv);
    if (nn) {
      CHK(upload(scol, &d_scol, c->stream)); tmp.p.push_back(d_scol);
      CU(cudaMalloc((void **)&d_rcol, std::max<size_t>(rcol.size(), 1) * sizeof(int))); tmp.p.push_back(d_rcol);
      int err = dohp_exchange_raw(c, s, 4, d_scol, w->soff, d_rcol, w->roff, c->stream);
      if (err) { delete w; return err; }
      if (!rcol.empty()) CU(cudaMemcpyAsync(rcol.data(), d_rcol, rcol.size() * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
      CU(cudaStreamSynchronize(c->stream));
    }
    // ---- 3. merge: final row = sorted union of my own columns and of every received list for that row -------------------------
    std::vector<std::vector<int>> extra;                              // per touched owned row
    std::vector<int> slot((size_t)n, -1);
    {
      long long e = 0;
      for (int r = 0; r < nn; r++)
        for (long long j = lro[r]; j < lro[r + 1]; j++) {
          const int row = sidx[(size_t)j];
          if (row < 0 || row >= n) { delete w; return seterr(63, "halo send index %d is not an owned node", row); }
          if (slot[row] < 0) { slot[row] = (int)extra.size(); extra.emplace_back(); }
          std::vector<int> &x = extra[slot[row]];
          x.insert(x.end(), rcol.begin() + e, rcol.begin() + e + rlen[(size_t)j]);
          e += rlen[(size_t)j];
        }
    }
    w->orowptr.assign((size_t)n + 1, 0);
    for (int i = 0; i < n; i++) {
      int len = rp[i + 1] - rp[i];
      if (slot[i] >= 0) {
        std::vector<int> &x = extra[slot[i]];
        x.insert(x.end(), ci.begin() + rp[i], ci.begin() + rp[i + 1]);
        std::sort(x.begin(), x.end());
        x.erase(std::unique(x.begin(), x.end()), x.end());
        len = (int)x.size();
      }
      const long long next = (long long)w->orowptr[i] + len;
      if (next >= (1LL << 31)) { delete w; return seterr(63, "Q1 owner pattern: more than 2^31 blocks"); }
      w->orowptr[i + 1] = (int)next;
    }
    w->onnz = w->orowptr[n];
    w->ocolidx.resize((size_t)w->onnz);
    std::vector<int> own_pos((size_t)w->nnz_own), recv_pos((size_t)w->nrecv);
    for (int i = 0; i < n; i++) {
      int *dst = w->ocolidx.data() + w->orowptr[i];
      if (slot[i] < 0) {
        std::copy(ci.begin() + rp[i], ci.begin() + rp[i + 1], dst);
        for (int k = rp[i]; k < rp[i + 1]; k++) own_pos[k] = w->orowptr[i] + (k - rp[i]);
      } else {
        const std::vector<int> &x = extra[slot[i]];
        std::copy(x.begin(), x.end(), dst);
        int q = 0;
        for (int k = rp[i]; k < rp[i + 1]; k++) { while (x[q] != ci[k]) q++; own_pos[k] = w->orowptr[i] + q; }     // both sorted
      }
    }
    {
      long long e = 0;
      for (int r = 0; r < nn; r++)
        for (long long j = lro[r]; j < lro[r + 1]; j++) {
          const int row = sidx[(size_t)j];
          const int *b = w->ocolidx.data() + w->orowptr[row], *en = w->ocolidx.data() + w->orowptr[row + 1];
          for (int t = 0; t < rlen[(size_t)j]; t++, e++) recv_pos[(size_t)e] = (int)(std::lower_bound(b, en, rcol[(size_t)e]) - w->ocolidx.data());
        }
    }
    CHK(upload(own_pos, &w->d_own_pos, c->stream));
    CHK(upload(send_src, &w->d_send_src, c->stream));
    CHK(upload(recv_pos, &w->d_recv_pos, c->stream));
    CU(cudaMalloc((void **)&w->d_sendbuf, std::max<size_t>((size_t)w->nsend * w->bs2, 1) * sizeof(double)));
    CU(cudaMalloc((void **)&w->d_recvbuf, std::max<size_t>((size_t)w->nrecv * w->bs2, 1) * sizeof(double)));
    CU(cudaStreamSynchronize(c->stream));
    o->q1owner = w;
  }
  if (nnz) *nnz = w->onnz;
  if (orowptr) CU(cudaMemcpy(orowptr, w->orowptr.data(), w->orowptr.size() * sizeof(int), cudaMemcpyDefault));
  if (ocolidx && w->onnz) CU(cudaMemcpy(ocolidx, w->ocolidx.data(), w->ocolidx.size() * sizeof(int), cudaMemcpyDefault));
  return 0;
}

extern "C" int dB200OpQ1OwnerReduce(dB200Op o, const double *vals, double *ovals)
{
  dB200Ctx c = o->ctx;
  CU(cudaSetDevice(c->device));
  dB200Space s = o->space[0];
  Q1Owner *w = (Q1Owner *)o->q1owner;
  if (!w) return seterr(73, "Q1 owner reduce: call dB200OpQ1OwnerPattern first");
  if (!vals || !ovals) return seterr(85, "Q1 owner reduce: null argument");
  cudaPointerAttributes a;
  const bool vdev = cudaPointerGetAttributes(&a, vals) == cudaSuccess && (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged);
  cudaGetLastError();
  const bool odev = cudaPointerGetAttributes(&a, ovals) == cudaSuccess && (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged);
  cudaGetLastError();
  const size_t nloc = (size_t)(w->nnz_own + w->nsend) * w->bs2, nout = (size_t)w->onnz * w->bs2;
  const double *dv = vals;
  double *dvo = ovals, *t1 = nullptr, *t2 = nullptr;
  struct Tmp { std::vector<void *> p; ~Tmp() { for (void *q : p) cudaFree(q); } } tmp;
  if (!vdev) { CU(cudaMalloc((void **)&t1, std::max<size_t>(nloc, 1) * sizeof(double))); tmp.p.push_back(t1); CU(cudaMemcpyAsync(t1, vals, nloc * sizeof(double), cudaMemcpyHostToDevice, c->stream)); dv = t1; }
  if (!odev) { CU(cudaMalloc((void **)&t2, std::max<size_t>(nout, 1) * sizeof(double))); tmp.p.push_back(t2); dvo = t2; }
  const int T = 256;
  auto blocks = [&](long long m) { return (unsigned)((m * w->bs2 + T - 1) / T); };
  CU(cudaMemsetAsync(dvo, 0, nout * sizeof(double), c->stream));
  if (w->nnz_own) { q1own_scatter_kernel<<<blocks(w->nnz_own), T, 0, c->stream>>>(w->nnz_own, w->bs2, w->d_own_pos, dv, dvo, 0); LAUNCHED(c); }
  if (s->nneigh) {
    if (w->nsend) { q1own_pack_kernel<<<blocks(w->nsend), T, 0, c->stream>>>(w->nsend, w->bs2, w->d_send_src, dv, w->d_sendbuf); LAUNCHED(c); }
    std::vector<long long> so(w->soff), ro(w->roff);
    for (auto &x : so) x *= w->bs2;
    for (auto &x : ro) x *= w->bs2;
    CHK(dohp_exchange_raw(c, s, 8, w->d_sendbuf, so, w->d_recvbuf, ro, c->stream));
    for (int r = 0; r < s->nneigh; r++) {                    // neighbour by neighbour: no two adds of one launch hit the same entry
      const long long m = w->roff[r + 1] - w->roff[r];
      if (m) { q1own_scatter_kernel<<<blocks(m), T, 0, c->stream>>>(m, w->bs2, w->d_recv_pos + w->roff[r], w->d_recvbuf + w->roff[r] * w->bs2, dvo, 1); LAUNCHED(c); }
    }
  }
  CU(cudaGetLastError());
  if (!odev) CU(cudaMemcpyAsync(ovals, t2, nout * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}
