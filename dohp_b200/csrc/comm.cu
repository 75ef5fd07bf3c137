// Ghost exchange of nodal vectors over NCCL (NVLink 5 / NVSwitch): the VecGhostUpdateBegin/End pair of
// dFSGlobalToExpanded / dFSExpandedToGlobal (src/fs/interface/fs.c:578-579 forward INSERT, :627-628 reverse ADD).
// One grouped ncclSend/ncclRecv per neighbour on the context's stream; pack/unpack kernels are in api.cu.
// NCCL is bound at run time (dlopen) so that the library has no link-time dependency on a particular libnccl:
// inside a torch process the copy torch already loaded is reused.
#include "internal.h"
#include <dlfcn.h>
#include <cstdlib>
#include <cstring>

using namespace dohp;

typedef struct { char internal[128]; } ncclUniqueId_t;
typedef void *ncclComm_p;
enum { NCCL_DOUBLE = 8 };
struct NcclApi {
  void *h = nullptr;
  int (*GetUniqueId)(ncclUniqueId_t *);
  int (*CommInitRank)(ncclComm_p *, int, ncclUniqueId_t, int);
  int (*CommDestroy)(ncclComm_p);
  int (*GroupStart)();
  int (*GroupEnd)();
  int (*Send)(const void *, size_t, int, int, ncclComm_p, cudaStream_t);
  int (*Recv)(void *, size_t, int, int, ncclComm_p, cudaStream_t);
  int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_p, cudaStream_t);
  const char *(*GetErrorString)(int);
};
static NcclApi g_nccl;

static int nccl_load()
{
  if (g_nccl.h) return 0;
  // by soname only (inside a torch process the copy torch already mapped is found first); DOHP_B200_NCCL names another library
  const char *env = getenv("DOHP_B200_NCCL");
  const char *names[] = {env ? env : "libnccl.so.2", "libnccl.so.2", "libnccl.so"};
  void *h = nullptr;
  for (const char *n : names) { h = dlopen(n, RTLD_NOW | RTLD_NOLOAD); if (h) break; }
  if (!h) for (const char *n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
  if (!h) return seterr(62, "cannot load libnccl.so.2: %s", dlerror());
#define SYM(field, name) do { *(void **)&g_nccl.field = dlsym(h, name); if (!g_nccl.field) return seterr(62, "libnccl lacks %s", name); } while (0)
  SYM(GetUniqueId, "ncclGetUniqueId"); SYM(CommInitRank, "ncclCommInitRank"); SYM(CommDestroy, "ncclCommDestroy");
  SYM(GroupStart, "ncclGroupStart"); SYM(GroupEnd, "ncclGroupEnd"); SYM(Send, "ncclSend"); SYM(Recv, "ncclRecv"); SYM(AllReduce, "ncclAllReduce");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  g_nccl.h = h;
  return 0;
}
#define NC(call) do { int e_ = (call); if (e_) return seterr(800 + e_, "%s: %s", #call, g_nccl.GetErrorString(e_)); } while (0)

extern "C" int dB200NcclUniqueId(void *id128)
{
  int err = nccl_load();
  if (err) return err;
  ncclUniqueId_t id;
  NC(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, 128);
  return 0;
}

extern "C" int dB200InitComm(int device, int rank, int nranks, const void *uid, dB200Ctx *ctx)
{
  int err = dB200Init(device, ctx);
  if (err) return err;
  (*ctx)->rank = rank; (*ctx)->nranks = nranks;
  if (nranks == 1) return 0;
  err = nccl_load();
  if (err) return err;
  ncclUniqueId_t id;
  memcpy(&id, uid, 128);
  ncclComm_p comm;
  NC(g_nccl.CommInitRank(&comm, nranks, id, rank));
  (*ctx)->nccl = comm;
  return 0;
}

void dohp_comm_destroy(dB200Ctx c)
{
  if (c->nccl && g_nccl.h) g_nccl.CommDestroy((ncclComm_p)c->nccl);
  c->nccl = nullptr;
}

// forward: sendbuf (packed by sendoff) -> neighbours' recvbuf (by recvoff); reverse swaps the roles of the two lists
int dohp_exchange(dB200Ctx c, dB200Space s, bool reverse, cudaStream_t stream)
{
  if (!c->nccl) return seterr(73, "no communicator");
  const std::vector<int> &so = reverse ? s->recvoff : s->sendoff, &ro = reverse ? s->sendoff : s->recvoff;
  // the packed data is always in d_recvbuf for reverse (gathered ghosts) and d_sendbuf for forward
  const double *src = reverse ? s->d_recvbuf : s->d_sendbuf;
  double *dst = reverse ? s->d_sendbuf : s->d_recvbuf;
  NC(g_nccl.GroupStart());
  for (int r = 0; r < s->nneigh; r++) {
    const size_t ns = (size_t)(so[r + 1] - so[r]) * s->bs, nr = (size_t)(ro[r + 1] - ro[r]) * s->bs;
    if (ns) NC(g_nccl.Send(src + (size_t)so[r] * s->bs, ns, NCCL_DOUBLE, s->ranks[r], (ncclComm_p)c->nccl, stream));
    if (nr) NC(g_nccl.Recv(dst + (size_t)ro[r] * s->bs, nr, NCCL_DOUBLE, s->ranks[r], (ncclComm_p)c->nccl, stream));
  }
  NC(g_nccl.GroupEnd());
  return 0;
}

// grouped point-to-point exchange of raw buffers with the neighbours of a space (setup and matrix-row traffic of q1_owner.cu):
// elem_bytes 4 (ncclInt32) or 8 (ncclFloat64); send/recv offsets in elements, one segment per neighbour
int dohp_exchange_raw(dB200Ctx c, dB200Space s, int elem_bytes, const void *src, const std::vector<long long> &so, void *dst, const std::vector<long long> &ro, cudaStream_t stream)
{
  if (!c->nccl) return seterr(73, "no communicator");
  const int type = elem_bytes == 4 ? 2 /* ncclInt32 */ : NCCL_DOUBLE;
  NC(g_nccl.GroupStart());
  for (int r = 0; r < s->nneigh; r++) {
    const size_t ns = (size_t)(so[r + 1] - so[r]), nr = (size_t)(ro[r + 1] - ro[r]);
    if (ns) NC(g_nccl.Send((const char *)src + (size_t)so[r] * elem_bytes, ns, type, s->ranks[r], (ncclComm_p)c->nccl, stream));
    if (nr) NC(g_nccl.Recv((char *)dst + (size_t)ro[r] * elem_bytes, nr, type, s->ranks[r], (ncclComm_p)c->nccl, stream));
  }
  NC(g_nccl.GroupEnd());
  return 0;
}

// minimum of one int over all ranks of the communicator (setup decisions every rank must take alike); d_val: device int, in place
int dohp_allreduce_min_int(dB200Ctx c, int *d_val, cudaStream_t stream)
{
  if (!c->nccl) return 0;
  NC(g_nccl.AllReduce(d_val, d_val, 1, 2 /* ncclInt32 */, 3 /* ncclMin */, (ncclComm_p)c->nccl, stream));
  return 0;
}
