// Multi-GPU exchange of the micro->macro coupling scalars and of the macro solution.
//
// No reference call site: in the reference the macro solver reads the micro solutions through
// raw pointers (src/elliptic-elliptic/manager.cpp:25-26, src/tools/pde_data.h:76-90).  With the
// macro DoFs sharded over ranks that hand-off becomes one allgather of N (bio: 2N) doubles and one
// broadcast of the macro solution per fixed-point iteration.  Both are
// latency-bound (<= 133 KB), so they are issued on the context's stream directly behind the
// coupling kernel, which writes its results in place into this rank's block of the gather buffer.
//
// Partition: rank r owns the contiguous block [r*B, min((r+1)*B, n_total)) with B = ceil(n_total/R).
#include <dlfcn.h>
#include <nccl.h>  // types only: the library is resolved at run time, see nccl_api()

#include <cstdint>
#include <cstring>

#include "ctx.h"

namespace msgpu {

// NCCL is bound lazily with dlopen instead of at link time.  A host process that also uses
// torch.distributed already carries torch's bundled libnccl.so.2; loading a second, older system
// copy at library-load time would shadow it (and break `import torch` if libmsgpu came first).
// RTLD_NOLOAD picks up whatever copy the process already has; only a stand-alone C++ host falls
// back to the system library.
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  const char* (*GetErrorString)(ncclResult_t);
  bool ok = false;
};

static const NcclApi& nccl_api() {
  static NcclApi api = [] {
    NcclApi a{};
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return a;
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    a.AllGather = reinterpret_cast<decltype(a.AllGather)>(dlsym(h, "ncclAllGather"));
    a.Broadcast = reinterpret_cast<decltype(a.Broadcast)>(dlsym(h, "ncclBroadcast"));
    a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(dlsym(h, "ncclGroupStart"));
    a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(dlsym(h, "ncclGroupEnd"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather && a.Broadcast && a.GroupStart &&
           a.GroupEnd && a.GetErrorString;
    return a;
  }();
  return api;
}

static int nccl_fail(msgpu_ctx* ctx, ncclResult_t r, const char* what) {
  ctx->err = std::string(what) + ": " + nccl_api().GetErrorString(r);
  return MSGPU_ERR_COMM;
}

int comm_allgather_macro_buffer(msgpu_ctx* ctx, int fields);

int comm_block(const msgpu_ctx* ctx) { return (ctx->n_total + ctx->n_ranks - 1) / ctx->n_ranks; }

int comm_ensure_buffers(msgpu_ctx* ctx) {
  const size_t padded = static_cast<size_t>(comm_block(ctx)) * ctx->n_ranks;
  if (ctx->gather_capacity >= 2 * padded) return MSGPU_OK;
  if (ctx->d_c) cudaFree(ctx->d_c);
  if (ctx->d_macro) cudaFree(ctx->d_macro);
  ctx->d_c = ctx->d_macro = nullptr;
  if (cudaMalloc(reinterpret_cast<void**>(&ctx->d_c), 2 * padded * sizeof(double)) != cudaSuccess ||
      cudaMalloc(reinterpret_cast<void**>(&ctx->d_macro), 2 * padded * sizeof(double)) != cudaSuccess) {
    ctx->err = "cannot allocate the gather buffers";
    return MSGPU_ERR_CUDA;
  }
  cudaMemsetAsync(ctx->d_c, 0, 2 * padded * sizeof(double), ctx->stream);
  ctx->gather_capacity = 2 * padded;
  return MSGPU_OK;
}

// In-place allgather of this rank's block(s) of d_c.  Layout: c0 at [0, padded), c1 at [padded, 2*padded).
int comm_allgather_coupling(msgpu_ctx* ctx, double*, double*) {
  ncclComm_t comm = static_cast<ncclComm_t>(ctx->nccl_comm);
  if (!comm) {
    ctx->err = "no NCCL communicator and no peer-memory exchange attached";
    return MSGPU_ERR_COMM;
  }
  const int B = comm_block(ctx);
  const size_t padded = static_cast<size_t>(B) * ctx->n_ranks;
  const int fields = ctx->desc.problem == MSGPU_BIO ? 2 : 1;
  const NcclApi& nc = nccl_api();
  ncclResult_t r = nc.GroupStart();
  if (r != ncclSuccess) return nccl_fail(ctx, r, "ncclGroupStart");
  for (int f = 0; f < fields; ++f) {
    double* base = ctx->d_c + f * padded;
    r = nc.AllGather(base + static_cast<size_t>(ctx->rank) * B, base, B, ncclDouble, comm, ctx->stream);
    if (r != ncclSuccess) return nccl_fail(ctx, r, "ncclAllGather");
  }
  r = nc.GroupEnd();
  if (r != ncclSuccess) return nccl_fail(ctx, r, "ncclGroupEnd");
  ctx->launches += 1;
  return MSGPU_OK;
}

// ---- peer-memory exchange ---------------------------------------------------------------------
// Every rank owns one exchange buffer: data[2 parities][2 fields][padded] doubles followed by one
// 32-bit flag per rank.  k_coupling (kernels_cg.cu) stores each integral into the data area of ALL
// ranks through peer-mapped pointers (NVLink / NVSwitch stores) and its last CTA writes the call
// number into flag[this rank] of every peer.  The consumer side is a stream memory operation
// (cuStreamWaitValue32, >=) on its own flags — no host synchronisation, no spinning kernel — followed
// by kernels / copies that read the parity just completed in place.  Two parities are enough: a rank
// can only be one call ahead of a peer, because its own wait needs that peer's flag of the same call.
namespace {
size_t p2p_padded(const msgpu_ctx* ctx) { return static_cast<size_t>(comm_block(ctx)) * ctx->n_ranks; }
size_t p2p_bytes(const msgpu_ctx* ctx) { return 4 * p2p_padded(ctx) * sizeof(double) + 64 * sizeof(unsigned int); }
unsigned int* p2p_flags(const msgpu_ctx* ctx, void* base) {
  return reinterpret_cast<unsigned int*>(static_cast<char*>(base) + 4 * p2p_padded(ctx) * sizeof(double));
}
typedef int (*WaitValue32Fn)(cudaStream_t, unsigned long long, unsigned int, unsigned int);
}  // namespace

void comm_p2p_fill(msgpu_ctx* ctx, CouplingArgs& a) {
  a.n_peers = 0;
  if (!ctx->p2p_ready) return;
  ctx->p2p_iter += 1;
  const size_t padded = p2p_padded(ctx);
  const size_t parity = ctx->p2p_iter & 1u;
  a.n_peers = ctx->n_ranks;
  for (int p = 0; p < ctx->n_ranks; ++p) {
    double* data = static_cast<double*>(ctx->p2p_peer[p]) + parity * 2 * padded;
    a.peer_c0[p] = data + ctx->k_offset;
    a.peer_c1[p] = data + padded + ctx->k_offset;
    a.peer_flag[p] = p2p_flags(ctx, ctx->p2p_peer[p]) + ctx->rank;
  }
  a.flag_value = ctx->p2p_iter;
  a.done_counter = ctx->d_p2p_done;
}

int comm_p2p_collect(msgpu_ctx* ctx) {
  const size_t padded = p2p_padded(ctx);
  const size_t parity = ctx->p2p_iter & 1u;
  WaitValue32Fn wait = reinterpret_cast<WaitValue32Fn>(ctx->wait_value_fn);
  unsigned int* flags = p2p_flags(ctx, ctx->p2p_local);
  for (int p = 0; p < ctx->n_ranks; ++p) {
    if (p == ctx->rank) continue;  // this rank's own stores are ordered by the stream itself
    const int rc = wait(ctx->stream, static_cast<unsigned long long>(reinterpret_cast<uintptr_t>(flags + p)), ctx->p2p_iter,
                        0u /* CU_STREAM_WAIT_VALUE_GEQ */);
    if (rc != 0) {
      ctx->err = "cuStreamWaitValue32 failed (" + std::to_string(rc) + ")";
      return MSGPU_ERR_COMM;
    }
  }
  // consumers read the parity just completed in place (no copy)
  ctx->c_view = static_cast<const double*>(ctx->p2p_local) + parity * 2 * padded;
  return MSGPU_OK;
}

void comm_destroy(msgpu_ctx* ctx) {
  if (ctx->p2p_ready) {
    cudaStreamSynchronize(ctx->stream);
    for (int p = 0; p < ctx->n_ranks; ++p)
      if (p != ctx->rank && ctx->p2p_peer[p]) cudaIpcCloseMemHandle(ctx->p2p_peer[p]);
    ctx->p2p_ready = 0;
  }
  if (ctx->p2p_local) cudaFree(ctx->p2p_local);
  if (ctx->d_p2p_done) cudaFree(ctx->d_p2p_done);
  ctx->p2p_local = nullptr;
  ctx->d_p2p_done = nullptr;
  if (ctx->nccl_comm) {
    nccl_api().CommDestroy(static_cast<ncclComm_t>(ctx->nccl_comm));
    ctx->nccl_comm = nullptr;
  }
}

}  // namespace msgpu

using namespace msgpu;

extern "C" {

int msgpu_comm_unique_id(char id[128]) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId u;
  if (!nccl_api().ok || nccl_api().GetUniqueId(&u) != ncclSuccess) return MSGPU_ERR_COMM;
  std::memcpy(id, &u, 128);
  return MSGPU_OK;
}

int msgpu_partition(int n_total, int n_ranks, int rank, int* k0, int* k1) {
  if (n_total < 0 || n_ranks < 1 || rank < 0 || rank >= n_ranks || !k0 || !k1) return MSGPU_ERR_INVALID;
  const int B = (n_total + n_ranks - 1) / n_ranks;
  *k0 = rank * B < n_total ? rank * B : n_total;
  *k1 = (rank + 1) * B < n_total ? (rank + 1) * B : n_total;
  return MSGPU_OK;
}

int msgpu_comm_init(msgpu_ctx* ctx, const char id[128], int rank, int n_ranks, int n_total, int k_offset) {
  if (!ctx || n_ranks < 1 || n_ranks > 64 || rank < 0 || rank >= n_ranks) return MSGPU_ERR_INVALID;
  int k0, k1;
  msgpu_partition(n_total, n_ranks, rank, &k0, &k1);
  if (k_offset != k0 || (ctx->N != 0 && ctx->N != k1 - k0)) {
    ctx->err = "rank block does not match msgpu_partition()";
    return MSGPU_ERR_INVALID;
  }
  if (cudaSetDevice(ctx->device) != cudaSuccess) return MSGPU_ERR_CUDA;
  if (id) {  // id == NULL: no NCCL communicator, the peer-memory exchange (msgpu_comm_p2p_*) must follow
    ncclUniqueId u;
    std::memcpy(&u, id, 128);
    ncclComm_t comm;
    if (!nccl_api().ok) {
      ctx->err = "libnccl.so.2 cannot be loaded";
      return MSGPU_ERR_COMM;
    }
    ncclResult_t r = nccl_api().CommInitRank(&comm, n_ranks, u, rank);
    if (r != ncclSuccess) return nccl_fail(ctx, r, "ncclCommInitRank");
    ctx->nccl_comm = comm;
  }
  ctx->rank = rank;
  ctx->n_ranks = n_ranks;
  ctx->n_total = n_total;
  ctx->k_offset = k_offset;
  int rc = comm_ensure_buffers(ctx);
  if (rc) return rc;
  ctx->comm_ready = 1;
  if (ctx->nccl_comm && n_ranks > 1) {
    // NCCL builds its channels and connects the peers at the first collective (0.3 - 0.5 s): pay that here, in the
    // set-up, not inside the first fixed-point cycle (round 2: 285 ms of the first stop test on two GPUs)
    rc = comm_allgather_macro_buffer(ctx, 1);
    if (rc) return rc;
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
      ctx->err = "NCCL warm-up allgather failed";
      return MSGPU_ERR_CUDA;
    }
  }
  return MSGPU_OK;
}

int msgpu_comm_p2p_export(msgpu_ctx* ctx, unsigned char handle[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  if (!ctx || !ctx->comm_ready || !handle) return MSGPU_ERR_INVALID;
  if (ctx->n_ranks > 8) {
    ctx->err = "the peer-memory exchange supports up to 8 ranks";
    return MSGPU_ERR_INVALID;
  }
  if (cudaSetDevice(ctx->device) != cudaSuccess) return MSGPU_ERR_CUDA;
  if (!ctx->p2p_local) {
    if (cudaMalloc(&ctx->p2p_local, p2p_bytes(ctx)) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void**>(&ctx->d_p2p_done), sizeof(int)) != cudaSuccess) {
      ctx->err = "cannot allocate the exchange buffer";
      return MSGPU_ERR_CUDA;
    }
    cudaMemset(ctx->p2p_local, 0, p2p_bytes(ctx));
    cudaMemset(ctx->d_p2p_done, 0, sizeof(int));
  }
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, ctx->p2p_local) != cudaSuccess) {
    ctx->err = std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(cudaGetLastError());
    return MSGPU_ERR_CUDA;
  }
  std::memcpy(handle, &h, 64);
  return MSGPU_OK;
}

int msgpu_comm_p2p_attach(msgpu_ctx* ctx, const unsigned char* handles) {
  if (!ctx || !ctx->comm_ready || !ctx->p2p_local || !handles) return MSGPU_ERR_INVALID;
  if (cudaSetDevice(ctx->device) != cudaSuccess) return MSGPU_ERR_CUDA;
  cudaDriverEntryPointQueryResult st;
  void* fn = nullptr;
  if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &st) != cudaSuccess || !fn ||
      st != cudaDriverEntryPointSuccess) {
    cudaGetLastError();
    ctx->err = "cuStreamWaitValue32 is not available";
    return MSGPU_ERR_COMM;
  }
  ctx->wait_value_fn = fn;
  for (int p = 0; p < ctx->n_ranks; ++p) {
    if (p == ctx->rank) {
      ctx->p2p_peer[p] = ctx->p2p_local;
      continue;
    }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handles + static_cast<size_t>(p) * 64, 64);
    if (cudaIpcOpenMemHandle(&ctx->p2p_peer[p], h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      ctx->err = std::string("cudaIpcOpenMemHandle (rank ") + std::to_string(p) + "): " + cudaGetErrorString(cudaGetLastError());
      return MSGPU_ERR_COMM;
    }
  }
  ctx->p2p_ready = 1;
  return MSGPU_OK;
}

// Per-grid scalars of all ranks (error / residual norms for the stop test):
// local[fields][N of this rank] -> all[fields][n_total] on every rank.  Host buffers, staged through the
// padded gather scratch; one grouped ncclAllGather.  Without a communicator it is a copy.
}  // extern "C"

namespace msgpu {
// In-place allgather of this rank's block(s) of d_macro (layout [fields][padded]); device buffers only.
int comm_allgather_macro_buffer(msgpu_ctx* ctx, int fields) {
  ncclComm_t comm = static_cast<ncclComm_t>(ctx->nccl_comm);
  if (!comm) {
    ctx->err = "no NCCL communicator";
    return MSGPU_ERR_COMM;
  }
  const int B = comm_block(ctx);
  const size_t padded = static_cast<size_t>(B) * ctx->n_ranks;
  const NcclApi& nc = nccl_api();
  ncclResult_t r = nc.GroupStart();
  if (r != ncclSuccess) return nccl_fail(ctx, r, "ncclGroupStart");
  for (int f = 0; f < fields; ++f) {
    double* base = ctx->d_macro + f * padded;
    r = nc.AllGather(base + static_cast<size_t>(ctx->rank) * B, base, B, ncclDouble, comm, ctx->stream);
    if (r != ncclSuccess) return nccl_fail(ctx, r, "ncclAllGather");
  }
  r = nc.GroupEnd();
  if (r != ncclSuccess) return nccl_fail(ctx, r, "ncclGroupEnd");
  ctx->launches += 1;
  return MSGPU_OK;
}
}  // namespace msgpu

extern "C" {

int msgpu_comm_allgather(msgpu_ctx* ctx, const double* local, double* all, int fields) {
  if (!ctx || !local || !all || fields < 1 || fields > 2) return MSGPU_ERR_INVALID;
  if (!ctx->comm_ready || ctx->n_ranks == 1) {
    std::memcpy(all, local, sizeof(double) * static_cast<size_t>(fields) * ctx->N);
    return MSGPU_OK;
  }
  if (cudaSetDevice(ctx->device) != cudaSuccess) return MSGPU_ERR_CUDA;
  ncclComm_t comm = static_cast<ncclComm_t>(ctx->nccl_comm);
  if (!comm) {
    ctx->err = "msgpu_comm_allgather needs the NCCL communicator (msgpu_comm_init with an id)";
    return MSGPU_ERR_COMM;
  }
  const int B = comm_block(ctx);
  const size_t padded = static_cast<size_t>(B) * ctx->n_ranks;
  const NcclApi& nc = nccl_api();
  for (int f = 0; f < fields; ++f)
    cudaMemcpyAsync(ctx->d_macro + f * padded + static_cast<size_t>(ctx->rank) * B, local + static_cast<size_t>(f) * ctx->N,
                    sizeof(double) * ctx->N, cudaMemcpyHostToDevice, ctx->stream);
  ncclResult_t r = nc.GroupStart();
  if (r != ncclSuccess) return nccl_fail(ctx, r, "ncclGroupStart");
  for (int f = 0; f < fields; ++f) {
    double* base = ctx->d_macro + f * padded;
    r = nc.AllGather(base + static_cast<size_t>(ctx->rank) * B, base, B, ncclDouble, comm, ctx->stream);
    if (r != ncclSuccess) return nccl_fail(ctx, r, "ncclAllGather");
  }
  r = nc.GroupEnd();
  if (r != ncclSuccess) return nccl_fail(ctx, r, "ncclGroupEnd");
  ctx->launches += 1;
  for (int f = 0; f < fields; ++f)
    cudaMemcpyAsync(all + static_cast<size_t>(f) * ctx->n_total, ctx->d_macro + f * padded, sizeof(double) * ctx->n_total,
                    cudaMemcpyDeviceToHost, ctx->stream);
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
    ctx->err = "allgather failed";
    return MSGPU_ERR_CUDA;
  }
  return MSGPU_OK;
}

// Leaves the peer-memory exchange again (msgpu_coupling goes back to ncclAllGather).  Safe to call when nothing is
// attached: hosts use it to fall back consistently when msgpu_comm_p2p_attach failed on any rank.
int msgpu_comm_p2p_detach(msgpu_ctx* ctx) {
  if (!ctx) return MSGPU_ERR_INVALID;
  if (cudaSetDevice(ctx->device) != cudaSuccess) return MSGPU_ERR_CUDA;
  cudaStreamSynchronize(ctx->stream);
  for (int p = 0; p < ctx->n_ranks && p < 8; ++p) {
    if (p != ctx->rank && ctx->p2p_peer[p]) cudaIpcCloseMemHandle(ctx->p2p_peer[p]);
    ctx->p2p_peer[p] = nullptr;
  }
  cudaGetLastError();
  ctx->p2p_ready = 0;
  return MSGPU_OK;
}

// u_all / w_all: [n_total] host arrays; significant on root, overwritten on the other ranks.
int msgpu_broadcast_macro(msgpu_ctx* ctx, double* u_all, double* w_all, int root) {
  if (!ctx || !ctx->comm_ready || !ctx->setup_done || !u_all) return MSGPU_ERR_INVALID;
  if (cudaSetDevice(ctx->device) != cudaSuccess) return MSGPU_ERR_CUDA;
  ncclComm_t comm = static_cast<ncclComm_t>(ctx->nccl_comm);
  if (!comm) {
    ctx->err = "msgpu_broadcast_macro needs the NCCL communicator (msgpu_comm_init with an id)";
    return MSGPU_ERR_COMM;
  }
  const size_t padded = static_cast<size_t>(comm_block(ctx)) * ctx->n_ranks;
  const int fields = w_all ? 2 : 1;
  double* host[2] = {u_all, w_all};
  if (ctx->rank == root)
    for (int f = 0; f < fields; ++f)
      cudaMemcpyAsync(ctx->d_macro + f * padded, host[f], sizeof(double) * ctx->n_total, cudaMemcpyHostToDevice,
                      ctx->stream);
  ncclResult_t r = nccl_api().Broadcast(ctx->d_macro, ctx->d_macro, fields * padded, ncclDouble, root, comm, ctx->stream);
  if (r != ncclSuccess) return nccl_fail(ctx, r, "ncclBroadcast");
  ctx->launches += 1;
  cudaMemcpyAsync(ctx->d_u, ctx->d_macro + ctx->k_offset, sizeof(double) * ctx->N, cudaMemcpyDeviceToDevice, ctx->stream);
  if (w_all)
    cudaMemcpyAsync(ctx->d_w, ctx->d_macro + padded + ctx->k_offset, sizeof(double) * ctx->N, cudaMemcpyDeviceToDevice,
                    ctx->stream);
  if (ctx->rank != root)
    for (int f = 0; f < fields; ++f)
      cudaMemcpyAsync(host[f], ctx->d_macro + f * padded, sizeof(double) * ctx->n_total, cudaMemcpyDeviceToHost,
                      ctx->stream);
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
    ctx->err = "broadcast failed";
    return MSGPU_ERR_CUDA;
  }
  return MSGPU_OK;
}

}  // extern "C"
