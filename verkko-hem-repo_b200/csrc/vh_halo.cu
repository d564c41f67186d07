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
#include "vh_p2p.cuh"

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
  int (*AllGather)(const void *, void *, size_t, int, nccl_comm, cudaStream_t);
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
  SYM(AllGather, "ncclAllGather")
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
// all-reduce of n scalars in place, one block of 32 threads per rank
__global__ void k_p2p_allreduce(VhP2P P, unsigned long long seq0, double *__restrict__ v, int n)
{
  for (int i = 0; i < n; ++i)
    {
      const double s = vh_p2p_allreduce_warp(P, seq0 + i, v[i]);
      if (threadIdx.x == 0)
        v[i] = s;
    }
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

// Fused ghost push (GMRES): map every neighbour's zbuf through CUDA IPC and learn, for each of our send nodes, the ghost
// slot it occupies in that neighbour's numbering (the neighbour's recv list, exchanged once with ncclSend/ncclRecv).
// Opt-in (VH_HALO_PUSH=1): the round's GPU budget ran out before this path could be re-measured on 2/8 GPUs, so the
// measured NCCL ghost refresh stays the default.
static int setup_ghost_push(vh_ctx *ctx, nccl_comm comm)
{
  const char *e = getenv("VH_HALO_PUSH");
  const int   n_ranks = ctx->n_ranks, rank = ctx->rank, np = (int)ctx->peer_rank.size();
  int         ok = (e && e[0] == '1') && np > 0 && np <= VH_P2P_MAX_RANKS && ctx->n_owned > 0;
  // IPC handles of zbuf, all-gathered
  cudaIpcMemHandle_t mine;
  std::memset(&mine, 0, sizeof(mine));
  if (ok && cudaIpcGetMemHandle(&mine, ctx->zbuf) != cudaSuccess)
    {
      cudaGetLastError();
      ok = 0;
    }
  const size_t      rec = sizeof(cudaIpcMemHandle_t) + 8;
  char             *d_all = nullptr;
  std::vector<char> h_all(rec * n_ranks, 0);
  VH_CUDA(cudaMalloc((void **)&d_all, rec * n_ranks));
  std::memcpy(&h_all[rec * rank], &mine, sizeof(mine));
  h_all[rec * rank + sizeof(mine)] = (char)ok;
  VH_CUDA(cudaMemcpy(d_all + rec * rank, &h_all[rec * rank], rec, cudaMemcpyHostToDevice));
  VH_NCCL(api().AllGather(d_all + rec * rank, d_all, rec, /*ncclChar*/ 0, comm, ctx->stream));
  VH_CUDA(cudaStreamSynchronize(ctx->stream));
  VH_CUDA(cudaMemcpy(h_all.data(), d_all, rec * n_ranks, cudaMemcpyDeviceToHost));
  cudaFree(d_all);
  for (int r = 0; r < n_ranks; ++r)
    ok = ok && h_all[rec * r + sizeof(mine)];
  // every rank now has the same `ok` (all flags were gathered); the exchanges below are collective only if ok
  if (!ok)
    return VH_OK;
  int opened = 1;
  for (int p = 0; p < np && opened; ++p)
    {
      cudaIpcMemHandle_t h;
      std::memcpy(&h, &h_all[rec * ctx->peer_rank[p]], sizeof(h));
      if (cudaIpcOpenMemHandle(&ctx->zpush_open[p], h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
        {
          cudaGetLastError();
          ctx->zpush_open[p] = nullptr;
          opened             = 0;
        }
    }
  // destination slots: the neighbour's recv list for us, in the order of our send list
  int32_t *d_dst = nullptr;
  VH_CUDA(cudaMalloc((void **)&d_dst, sizeof(int32_t) * (size_t)std::max<int64_t>(ctx->n_send, 1)));
  VH_NCCL(api().GroupStart());
  for (int p = 0; p < np; ++p)
    {
      const size_t cs = (size_t)(ctx->send_ptr[p + 1] - ctx->send_ptr[p]), cr = (size_t)(ctx->recv_ptr[p + 1] - ctx->recv_ptr[p]);
      if (cr)
        VH_NCCL(api().Send(ctx->recv_nodes + ctx->recv_ptr[p], cr, /*ncclInt32*/ 2, ctx->peer_rank[p], comm, ctx->stream));
      if (cs)
        VH_NCCL(api().Recv(d_dst + ctx->send_ptr[p], cs, /*ncclInt32*/ 2, ctx->peer_rank[p], comm, ctx->stream));
    }
  VH_NCCL(api().GroupEnd());
  VH_CUDA(cudaStreamSynchronize(ctx->stream));
  std::vector<int32_t> h_dst((size_t)ctx->n_send);
  if (ctx->n_send)
    VH_CUDA(cudaMemcpy(h_dst.data(), d_dst, sizeof(int32_t) * (size_t)ctx->n_send, cudaMemcpyDeviceToHost));
  cudaFree(d_dst);
  // agree on success of the mappings
  double *d_flag = nullptr, h_flag = opened ? 1.0 : 0.0;
  VH_CUDA(cudaMalloc((void **)&d_flag, sizeof(double)));
  VH_CUDA(cudaMemcpy(d_flag, &h_flag, sizeof(double), cudaMemcpyHostToDevice));
  VH_NCCL(api().AllReduce(d_flag, d_flag, 1, NCCL_FLOAT64, /*ncclMin*/ 3, comm, ctx->stream));
  VH_CUDA(cudaStreamSynchronize(ctx->stream));
  VH_CUDA(cudaMemcpy(&h_flag, d_flag, sizeof(double), cudaMemcpyDeviceToHost));
  cudaFree(d_flag);
  if (h_flag != 1.0)
    return VH_OK;
  // CSR over the owned nodes: (peer slot, destination ghost node)
  std::vector<int32_t> pptr(ctx->n_owned + 1, 0), pdst((size_t)ctx->n_send);
  std::vector<int8_t>  ppeer((size_t)ctx->n_send);
  for (int64_t i = 0; i < ctx->n_send; ++i)
    pptr[ctx->h_send_nodes[i] + 1]++;
  for (int i = 0; i < ctx->n_owned; ++i)
    pptr[i + 1] += pptr[i];
  std::vector<int32_t> fill(pptr.begin(), pptr.end() - 1);
  for (int p = 0; p < np; ++p)
    for (int i = ctx->send_ptr[p]; i < ctx->send_ptr[p + 1]; ++i)
      {
        const int k = fill[ctx->h_send_nodes[i]]++;
        pdst[k]     = h_dst[i];
        ppeer[k]    = (int8_t)p;
      }
  VH_TRY(vh_dev_upload(ctx, &ctx->push_ptr, pptr.data(), pptr.size()));
  VH_TRY(vh_dev_upload(ctx, &ctx->push_dst, pdst.data(), pdst.size()));
  VH_TRY(vh_dev_upload(ctx, &ctx->push_peer, ppeer.data(), ppeer.size()));
  VH_CUDA(cudaMalloc((void **)&ctx->push_ticket, sizeof(unsigned int)));
  VH_CUDA(cudaMemset(ctx->push_ticket, 0, sizeof(unsigned int)));
  VhPush &H = ctx->zpush_dev;
  H         = VhPush();
  for (int p = 0; p < np; ++p)
    {
      const int r  = ctx->peer_rank[p];
      H.zpeer[p]    = static_cast<double *>(ctx->zpush_open[p]);
      H.flag_dst[p] = ctx->p2p_dev.peer[r] + VH_P2P_SLOTS * n_ranks + rank; // my cell in the neighbour's mailbox
      H.flag_src[p] = ctx->p2p_dev.peer[rank] + VH_P2P_SLOTS * n_ranks + r; // the neighbour's cell in mine
    }
  H.push_ptr  = ctx->push_ptr;
  H.push_dst  = ctx->push_dst;
  H.push_peer = ctx->push_peer;
  H.ticket    = ctx->push_ticket;
  H.n_peers   = np;
  H.err       = ctx->p2p_err;
  ctx->zpush_seq = 0;
  ctx->zpush     = true;
  return VH_OK;
}

static int comm_setup(vh_ctx *ctx, int rank, int n_ranks, nccl_comm comm);

static int comm_args_ok(vh_ctx *ctx, int rank, int n_ranks)
{
  if (!ctx || n_ranks < 1 || rank < 0 || rank >= n_ranks)
    return vh_fail(ctx, VH_ERR_ARG, "vh_comm_init: bad rank / n_ranks");
  for (int p : ctx->peer_rank)
    if (p < 0 || p >= n_ranks || p == rank)
      return vh_fail(ctx, VH_ERR_ARG, "vh_comm_init: halo plan names a peer outside the communicator");
  return VH_OK;
}

extern "C" int vh_comm_init(vh_ctx *ctx, int rank, int n_ranks, const void *unique_id)
{
  VH_TRY(comm_args_ok(ctx, rank, n_ranks));
  ctx->rank    = rank;
  ctx->n_ranks = n_ranks;
  if (n_ranks == 1)
    return VH_OK;
  if (!unique_id)
    return vh_fail(ctx, VH_ERR_ARG, "vh_comm_init: unique_id is null");
  if (!load_nccl())
    return vh_fail(ctx, VH_ERR_NCCL, api().err);
  VH_CUDA(cudaSetDevice(ctx->device));
  nccl_uid id;
  std::memcpy(&id, unique_id, sizeof(id));
  nccl_comm comm = nullptr;
  VH_NCCL(api().CommInitRank(&comm, n_ranks, id, rank));
  ctx->nccl_comm   = comm;
  ctx->nccl_holder = std::shared_ptr<void>(comm, [](void *c) {
    if (c && api().handle)
      api().CommDestroy((nccl_comm)c);
  });
  return comm_setup(ctx, rank, n_ranks, comm);
}

// Same communicator for another context of the same rank (the next mesh of an adaptive cycle, the levels of a multigrid
// hierarchy): ncclCommInitRank and the lazy peer connections of a new communicator cost seconds on 8 ranks, the mailboxes
// and the ghost-push mapping of a context milliseconds.  Collective like vh_comm_init; the communicator lives until its
// last context is destroyed.  The contexts sharing one are driven by one host thread in the same order on every rank.
extern "C" int vh_comm_share(vh_ctx *ctx, vh_ctx *donor)
{
  if (!ctx || !donor || ctx == donor)
    return vh_fail(ctx, VH_ERR_ARG, "vh_comm_share: null or identical contexts");
  if (donor->device != ctx->device)
    return vh_fail(ctx, VH_ERR_ARG, "vh_comm_share: the two contexts live on different devices");
  VH_TRY(comm_args_ok(ctx, donor->rank, donor->n_ranks));
  ctx->rank    = donor->rank;
  ctx->n_ranks = donor->n_ranks;
  if (donor->n_ranks == 1)
    return VH_OK;
  if (!donor->nccl_comm || !donor->nccl_holder)
    return vh_fail(ctx, VH_ERR_STATE, "vh_comm_share: the donor context has no communicator (vh_comm_init)");
  VH_CUDA(cudaSetDevice(ctx->device));
  VH_CUDA(cudaStreamSynchronize(donor->stream)); // nothing of the donor is in flight on the communicator
  ctx->nccl_comm   = donor->nccl_comm;
  ctx->nccl_holder = donor->nccl_holder;
  return comm_setup(ctx, ctx->rank, ctx->n_ranks, (nccl_comm)ctx->nccl_comm);
}

static int comm_setup(vh_ctx *ctx, int rank, int n_ranks, nccl_comm comm)
{
  // ---- peer-memory mailboxes for the scalar all-reduces (VH_P2P=0 keeps ncclAllReduce) ----
  const char *e = getenv("VH_P2P");
  if (n_ranks <= VH_P2P_MAX_RANKS && !(e && e[0] == '0'))
    {
      int                ok = vh_p2p_alloc_local(ctx, n_ranks) == VH_OK;
      cudaIpcMemHandle_t mine;
      std::memset(&mine, 0, sizeof(mine));
      if (ok && cudaIpcGetMemHandle(&mine, ctx->p2p_mbox) != cudaSuccess)
        ok = 0;
      cudaGetLastError();
      // exchange the IPC handles (and the per-rank status) with one ncclAllGather of bytes
      const size_t rec = sizeof(cudaIpcMemHandle_t) + 8;
      char        *d_all = nullptr;
      VH_CUDA(cudaMalloc((void **)&d_all, rec * n_ranks));
      std::vector<char> h_all(rec * n_ranks, 0);
      std::memcpy(&h_all[rec * rank], &mine, sizeof(mine));
      h_all[rec * rank + sizeof(mine)] = (char)ok;
      VH_CUDA(cudaMemcpy(d_all + rec * rank, &h_all[rec * rank], rec, cudaMemcpyHostToDevice));
      VH_NCCL(api().AllGather(d_all + rec * rank, d_all, rec, /*ncclChar*/ 0, comm, ctx->stream));
      VH_CUDA(cudaStreamSynchronize(ctx->stream));
      VH_CUDA(cudaMemcpy(h_all.data(), d_all, rec * n_ranks, cudaMemcpyDeviceToHost));
      for (int r = 0; r < n_ranks; ++r)
        ok = ok && h_all[rec * r + sizeof(mine)];
      int opened = ok;
      if (ok)
        for (int r = 0; r < n_ranks && opened; ++r)
          {
            if (r == rank)
              continue;
            cudaIpcMemHandle_t h;
            std::memcpy(&h, &h_all[rec * r], sizeof(h));
            if (cudaIpcOpenMemHandle(&ctx->p2p_open[r], h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
              {
                cudaGetLastError();
                ctx->p2p_open[r] = nullptr;
                opened           = 0;
              }
          }
      // every rank must take the same path: agree on the minimum
      double *d_flag = reinterpret_cast<double *>(d_all); // reuse (rec * n_ranks >= 8 bytes, 256-byte aligned)
      double  h_flag = opened ? 1.0 : 0.0;
      VH_CUDA(cudaMemcpy(d_flag, &h_flag, sizeof(double), cudaMemcpyHostToDevice));
      VH_NCCL(api().AllReduce(d_flag, d_flag, 1, NCCL_FLOAT64, /*ncclMin*/ 3, comm, ctx->stream));
      VH_CUDA(cudaStreamSynchronize(ctx->stream));
      VH_CUDA(cudaMemcpy(&h_flag, d_flag, sizeof(double), cudaMemcpyDeviceToHost));
      cudaFree(d_all);
      if (h_flag == 1.0)
        {
          // the fused Gram-Schmidt kernel must be chosen identically on every rank: agree on the largest requirement
          double *d_mode = nullptr, h_mode = (double)vhk_mgs_mode_local(ctx);
          VH_CUDA(cudaMalloc((void **)&d_mode, sizeof(double)));
          VH_CUDA(cudaMemcpy(d_mode, &h_mode, sizeof(double), cudaMemcpyHostToDevice));
          VH_NCCL(api().AllReduce(d_mode, d_mode, 1, NCCL_FLOAT64, /*ncclMax*/ 2, comm, ctx->stream));
          VH_CUDA(cudaStreamSynchronize(ctx->stream));
          VH_CUDA(cudaMemcpy(&h_mode, d_mode, sizeof(double), cudaMemcpyDeviceToHost));
          cudaFree(d_mode);
          ctx->mgs_mode    = (int)h_mode;
          ctx->p2p         = true;
          ctx->p2p_dev.me  = rank;
          ctx->p2p_dev.n   = n_ranks;
          ctx->p2p_dev.err = ctx->p2p_err;
          for (int r = 0; r < n_ranks; ++r)
            ctx->p2p_dev.peer[r] = static_cast<VhP2PCell *>(r == rank ? ctx->p2p_mbox : ctx->p2p_open[r]);
          VH_TRY(setup_ghost_push(ctx, comm));
        }
    }
  return VH_OK;
}

// (Re)allocates this rank's mailbox for n_ranks senders and points p2p_dev at it as a one-rank communicator; vh_comm_init
// then replaces the peer pointers by the IPC-mapped mailboxes of the other ranks.
int vh_p2p_alloc_local(vh_ctx *ctx, int n_ranks)
{
  VH_CUDA(cudaSetDevice(ctx->device));
  if (ctx->p2p_mbox)
    {
      VH_CUDA(cudaStreamSynchronize(ctx->stream));
      cudaFree(ctx->p2p_mbox);
      ctx->p2p_mbox = nullptr;
    }
  const size_t bytes = sizeof(VhP2PCell) * (VH_P2P_SLOTS + 1) * n_ranks; // all-reduce slots, then one ghost-push flag per sender
  VH_CUDA(cudaMalloc(&ctx->p2p_mbox, bytes));
  VH_CUDA(cudaMemset(ctx->p2p_mbox, 0, bytes));
  if (!ctx->p2p_err)
    {
      VH_CUDA(cudaMalloc((void **)&ctx->p2p_err, sizeof(int)));
      VH_CUDA(cudaMemset(ctx->p2p_err, 0, sizeof(int)));
      VH_CUDA(cudaMalloc((void **)&ctx->mgs_tickets, sizeof(unsigned int) * (VH_MAX_RESTART + 2)));
      VH_CUDA(cudaMemset(ctx->mgs_tickets, 0, sizeof(unsigned int) * (VH_MAX_RESTART + 2)));
    }
  ctx->p2p_seq = 0;
  ctx->p2p_dev = VhP2P();
  ctx->p2p_dev.peer[0] = static_cast<VhP2PCell *>(ctx->p2p_mbox);
  ctx->p2p_dev.me      = 0;
  ctx->p2p_dev.n       = 1;
  ctx->p2p_dev.err     = ctx->p2p_err;
  return VH_OK;
}

void vh_comm_destroy(vh_ctx *ctx)
{
  if (ctx->p2p_mbox)
    { // nobody may still be spinning on / writing to a mailbox (or pushing into a vector) that is about to go away:
      // finish our own work, then meet the other ranks (vh_destroy is collective like every other call)
      cudaStreamSynchronize(ctx->stream);
      if (ctx->p2p && ctx->n_ranks > 1)
        {
          k_p2p_allreduce<<<1, 32, 0, ctx->stream>>>(ctx->p2p_dev, ctx->p2p_seq + 1, ctx->scal + VH_SCAL_MISC, 1);
          ctx->p2p_seq += 1;
          cudaStreamSynchronize(ctx->stream);
        }
      for (int r = 0; r < VH_P2P_MAX_RANKS; ++r)
        {
          if (ctx->p2p_open[r])
            cudaIpcCloseMemHandle(ctx->p2p_open[r]);
          if (ctx->zpush_open[r])
            cudaIpcCloseMemHandle(ctx->zpush_open[r]);
          ctx->p2p_open[r] = ctx->zpush_open[r] = nullptr;
        }
      ctx->zpush = false;
      cudaFree(ctx->p2p_mbox);
      cudaFree(ctx->p2p_err);
      cudaFree(ctx->mgs_tickets);
      ctx->p2p_mbox    = nullptr;
      ctx->p2p_err     = nullptr;
      ctx->mgs_tickets = nullptr;
      ctx->p2p         = false;
    }
  ctx->nccl_holder.reset(); // ncclCommDestroy when the last context sharing the communicator lets go
  ctx->nccl_comm = nullptr;
}

int vhk_allreduce_sum(vh_ctx *ctx, double *dev, int n)
{
  if (ctx->n_ranks == 1)
    return VH_OK;
  if (!ctx->nccl_comm)
    return vh_fail(ctx, VH_ERR_STATE, "multi-rank context without vh_comm_init");
  if (ctx->p2p)
    { // latency path: one 32-thread kernel that posts to / polls the peer mailboxes over NVLink
      k_p2p_allreduce<<<1, 32, 0, ctx->stream>>>(ctx->p2p_dev, ctx->p2p_seq + 1, dev, n);
      ctx->p2p_seq += n;
      VH_LAUNCH_CHECK();
      return VH_OK;
    }
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
