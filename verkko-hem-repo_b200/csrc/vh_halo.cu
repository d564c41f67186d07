// Ghost-DoF halo exchange and scalar all-reduce over NCCL (NVLink 5 / NVSwitch inside one B200 box).
//
// Replaces the communication deal.II/Trilinos hide behind
//   ghosted assignment  locally_relevant_* = distributed_*   (solve.cc:183, iteration.cc:145,183,210)  -> halo exchange
//   Vector::l2_norm / operator* / add_and_dot                (solve.cc:158, SolverFGMRES)                -> all-reduce
// Matrix/rhs export (compress(add), assemble.cc:369-370) does not exist here: every rank assembles complete
// owned rows from its ghost-cell layer.
//
// NCCL is bound with dlopen so the library has no link-time NCCL dependency: inside a torchrun worker the
// already-loaded torch-bundled libnccl.so.2 is reused, in a plain C++ host the system one is loaded.
#include "vh_internal.h"

#include <dlfcn.h>

#include <cstdlib>
#include <cstring>

namespace
{
typedef struct
{
  char internal[VH_NCCL_UNIQUE_ID_BYTES];
} nccl_uid;
typedef void *nccl_comm;
enum
{
  NCCL_FLOAT64 = 8,
  NCCL_SUM     = 0
};

struct NcclApi
{
  void *handle = nullptr;
  int (*GetUniqueId)(nccl_uid *);
  int (*CommInitRank)(nccl_comm *, int, nccl_uid, int);
  int (*CommDestroy)(nccl_comm);
  int (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm, cudaStream_t);
  int (*Send)(const void *, size_t, int, int, nccl_comm, cudaStream_t);
  int (*Recv)(void *, size_t, int, int, nccl_comm, cudaStream_t);
  int (*GroupStart)();
  int (*GroupEnd)();
  const char *(*GetErrorString)(int);
  std::string err;
};

NcclApi &api()
{
  static NcclApi a;
  return a;
}

bool load_nccl()
{
  NcclApi &a = api();
  if (a.handle)
    return true;
  const char *names[] = {getenv("VH_NCCL_LIB"), "libnccl.so.2", "libnccl.so", "/usr/lib/x86_64-linux-gnu/libnccl.so.2"};
  for (const char *n : names)
    {
      if (!n || !*n)
        continue;
      a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (a.handle)
        break;
    }
  if (!a.handle)
    {
      a.err = std::string("cannot dlopen libnccl: ") + (dlerror() ? dlerror() : "?");
      return false;
    }
#define SYM(field, name)                                             \
  *(void **)(&a.field) = dlsym(a.handle, name);                      \
  if (!a.field)                                                      \
    {                                                                \
      a.err    = std::string("NCCL symbol missing: ") + name;        \
      a.handle = nullptr;                                            \
      return false;                                                  \
    }
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(AllReduce, "ncclAllReduce")
  SYM(Send, "ncclSend")
  SYM(Recv, "ncclRecv")
  SYM(GroupStart, "ncclGroupStart")
  SYM(GroupEnd, "ncclGroupEnd")
  SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  return true;
}

__global__ void k_pack(int64_t n, const int32_t *__restrict__ nodes, const double *__restrict__ x, double *__restrict__ buf)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    buf[i] = x[18 * (int64_t)nodes[i / 18] + i % 18];
}
__global__ void k_unpack(int64_t n, const int32_t *__restrict__ nodes, const double *__restrict__ buf, double *__restrict__ x)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    x[18 * (int64_t)nodes[i / 18] + i % 18] = buf[i];
}
} // namespace

#define VH_NCCL(call)                                                                                          \
  do                                                                                                           \
    {                                                                                                          \
      int r__ = (call);                                                                                        \
      if (r__ != 0)                                                                                            \
        return vh_fail(ctx, VH_ERR_NCCL, std::string(#call) + ": " + api().GetErrorString(r__));               \
    }                                                                                                          \
  while (0)

extern "C" int vh_nccl_unique_id(void *id_out)
{
  vh_ctx *ctx = nullptr;
  if (!id_out)
    return vh_fail(ctx, VH_ERR_ARG, "vh_nccl_unique_id: null");
  if (!load_nccl())
    return vh_fail(ctx, VH_ERR_NCCL, api().err);
  nccl_uid id;
  VH_NCCL(api().GetUniqueId(&id));
  std::memcpy(id_out, &id, sizeof(id));
  return VH_OK;
}

extern "C" int vh_comm_init(vh_ctx *ctx, int rank, int n_ranks, const void *unique_id)
{
  if (!ctx || n_ranks < 1 || rank < 0 || rank >= n_ranks)
    return vh_fail(ctx, VH_ERR_ARG, "vh_comm_init: bad rank / n_ranks");
  ctx->rank    = rank;
  ctx->n_ranks = n_ranks;
  if (n_ranks == 1)
    return VH_OK;
  if (!unique_id)
    return vh_fail(ctx, VH_ERR_ARG, "vh_comm_init: unique_id is null");
  for (int p : ctx->peer_rank)
    if (p < 0 || p >= n_ranks || p == rank)
      return vh_fail(ctx, VH_ERR_ARG, "vh_comm_init: halo plan names a peer outside the communicator");
  if (!load_nccl())
    return vh_fail(ctx, VH_ERR_NCCL, api().err);
  VH_CUDA(cudaSetDevice(ctx->device));
  nccl_uid id;
  std::memcpy(&id, unique_id, sizeof(id));
  nccl_comm comm = nullptr;
  VH_NCCL(api().CommInitRank(&comm, n_ranks, id, rank));
  ctx->nccl_comm = comm;
  return VH_OK;
}

void vh_comm_destroy(vh_ctx *ctx)
{
  if (ctx->nccl_comm && api().handle)
    api().CommDestroy((nccl_comm)ctx->nccl_comm);
  ctx->nccl_comm = nullptr;
}

int vhk_allreduce_sum(vh_ctx *ctx, double *dev, int n)
{
  if (ctx->n_ranks == 1)
    return VH_OK;
  if (!ctx->nccl_comm)
    return vh_fail(ctx, VH_ERR_STATE, "multi-rank context without vh_comm_init");
  VH_NCCL(api().AllReduce(dev, dev, (size_t)n, NCCL_FLOAT64, NCCL_SUM, (nccl_comm)ctx->nccl_comm, ctx->stream));
  return VH_OK;
}

int vhk_halo_exchange(vh_ctx *ctx, double *x_local)
{
  if (ctx->n_ranks == 1 || ctx->peer_rank.empty())
    {
      if (ctx->n_ghost > 0 && ctx->n_ranks == 1 && !ctx->peer_rank.empty())
        return vh_fail(ctx, VH_ERR_STATE, "ghost nodes present but vh_comm_init was not called");
      return VH_OK;
    }
  if (!ctx->nccl_comm)
    return vh_fail(ctx, VH_ERR_STATE, "multi-rank context without vh_comm_init");
  const int64_t ns = 18 * ctx->n_send, nr = 18 * ctx->n_recv;
  if (ns > 0)
    {
      k_pack<<<(unsigned)((ns + 255) / 256), 256, 0, ctx->stream>>>(ns, ctx->send_nodes, x_local, ctx->send_buf);
      VH_LAUNCH_CHECK();
    }
  VH_NCCL(api().GroupStart());
  for (size_t p = 0; p < ctx->peer_rank.size(); ++p)
    {
      const size_t cs = 18 * (size_t)(ctx->send_ptr[p + 1] - ctx->send_ptr[p]);
      const size_t cr = 18 * (size_t)(ctx->recv_ptr[p + 1] - ctx->recv_ptr[p]);
      if (cs)
        VH_NCCL(api().Send(ctx->send_buf + 18 * (size_t)ctx->send_ptr[p], cs, NCCL_FLOAT64, ctx->peer_rank[p],
                           (nccl_comm)ctx->nccl_comm, ctx->stream));
      if (cr)
        VH_NCCL(api().Recv(ctx->recv_buf + 18 * (size_t)ctx->recv_ptr[p], cr, NCCL_FLOAT64, ctx->peer_rank[p],
                           (nccl_comm)ctx->nccl_comm, ctx->stream));
    }
  VH_NCCL(api().GroupEnd());
  if (nr > 0)
    {
      k_unpack<<<(unsigned)((nr + 255) / 256), 256, 0, ctx->stream>>>(nr, ctx->recv_nodes, ctx->recv_buf, x_local);
      VH_LAUNCH_CHECK();
    }
  return VH_OK;
}
