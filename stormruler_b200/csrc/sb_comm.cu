// sb_comm.cu -- multi-GPU communicator: the symmetric slab + vector pool, CUDA-IPC peer mapping
// (P2P mode), NCCL loaded at run time (NCCL mode), and the halo exchange driver.
// Device-side protocol: sb_comm.cuh. Scheme: SURVEY.md 8e, DESIGN.md "Multi-GPU".
#include "sb_comm.cuh"

#include <dlfcn.h>
#include <nccl.h> // types and enums only: the library itself is dlopen'ed (libnccl.so.2)

#include <cstdlib>

#include "sb_op.cuh"

namespace sb {

// ---- NCCL, bound at run time so that the single-GPU path has no NCCL dependency at all ----------------
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    // If torch is already in the process its bundled libnccl.so.2 is resolved by soname; otherwise the
    // system library is used.
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (lib == nullptr) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (lib != nullptr) {
      api.lib = lib;
#define SB_SYM(field, name) *reinterpret_cast<void**>(&api.field) = dlsym(lib, name)
      SB_SYM(GetUniqueId, "ncclGetUniqueId");
      SB_SYM(CommInitRank, "ncclCommInitRank");
      SB_SYM(CommDestroy, "ncclCommDestroy");
      SB_SYM(AllReduce, "ncclAllReduce");
      SB_SYM(Send, "ncclSend");
      SB_SYM(Recv, "ncclRecv");
      SB_SYM(GroupStart, "ncclGroupStart");
      SB_SYM(GroupEnd, "ncclGroupEnd");
      SB_SYM(GetErrorString, "ncclGetErrorString");
#undef SB_SYM
      if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce || !api.Send || !api.Recv ||
          !api.GroupStart || !api.GroupEnd || !api.GetErrorString)
        api.lib = nullptr;
    }
  }
  return api.lib != nullptr ? &api : nullptr;
}

#define SB_NCCL(expr)                                                                                   \
  do {                                                                                                  \
    ncclResult_t _r = (expr);                                                                           \
    if (_r != ncclSuccess) {                                                                            \
      ::sb::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, nccl_api()->GetErrorString(_r)); \
      return SB_ERR_NCCL;                                                                               \
    }                                                                                                   \
  } while (0)

struct Blob { // SB_COMM_BLOB_BYTES
  cudaIpcMemHandle_t ipc; // 64 bytes
  ncclUniqueId nccl_id;   // 128 bytes (valid in rank 0's blob, NCCL mode)
  int32_t rank, world, mode, device;
  int64_t slab_bytes;
  int64_t vec_capacity;
  int32_t n_vectors;
  int32_t pid;
  char pad[256 - 64 - 128 - 16 - 16 - 8];
};
static_assert(sizeof(Blob) == SB_COMM_BLOB_BYTES, "blob layout");

// ---- vector storage ------------------------------------------------------------------------------------
// Cache limit: SB_VEC_CACHE_MB (0 disables the cache), default a quarter of the device memory.
static int64_t vec_cache_limit(sb_ctx* ctx) {
  if (ctx->vec_cache_limit < 0) {
    size_t free_b = 0, total_b = 0;
    if (const char* e = std::getenv("SB_VEC_CACHE_MB")) {
      ctx->vec_cache_limit = std::max<int64_t>(0, std::atoll(e)) << 20;
    } else if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
      ctx->vec_cache_limit = (int64_t) (total_b / 4);
    } else {
      ctx->vec_cache_limit = 0;
    }
  }
  return ctx->vec_cache_limit;
}

void vec_cache_release(sb_ctx* ctx) {
  if (ctx->vec_cache.empty()) return;
  cudaStreamSynchronize(ctx->stream);
  for (auto& kv : ctx->vec_cache) {
    ctx->vec_cap.erase(kv.second);
    cudaFree(kv.second);
  }
  ctx->vec_cache.clear();
  ctx->vec_cache_bytes = 0;
}

int vec_alloc(sb_ctx* ctx, size_t n, double** out) {
  *out = nullptr;
  if (ctx->comm.mode < 0) {
    const int64_t cap = pad_up((int64_t) n > 0 ? (int64_t) n : 1);
    double* d = nullptr;
    auto hit = ctx->vec_cache.find(cap);
    if (hit != ctx->vec_cache.end()) {
      d = hit->second;
      ctx->vec_cache.erase(hit);
      ctx->vec_cache_bytes -= (int64_t) sizeof(double) * cap;
    } else {
      cudaError_t e = cudaMalloc(&d, sizeof(double) * cap);
      if (e == cudaErrorMemoryAllocation && !ctx->vec_cache.empty()) {
        (void) cudaGetLastError();
        vec_cache_release(ctx); // give the cached blocks back and try once more
        e = cudaMalloc(&d, sizeof(double) * cap);
      }
      SB_CUDA(e);
      ctx->vec_cap[d] = cap;
    }
    SB_CUDA(cudaMemsetAsync(d, 0, sizeof(double) * cap, ctx->stream));
    *out = d;
    return SB_OK;
  }
  if ((int64_t) n > ctx->vec_capacity) {
    set_error("vector of %zu elements does not fit a pool block of %lld (sb_comm_prepare vec_capacity)", n,
              (long long) ctx->vec_capacity);
    return SB_ERR_INVALID;
  }
  if (ctx->pool_free.empty()) {
    set_error("vector pool exhausted (%d blocks; raise n_vectors in sb_comm_prepare)", ctx->pool_blocks);
    return SB_ERR_NOMEM;
  }
  double* d = ctx->pool_free.back();
  ctx->pool_free.pop_back();
  // Zero the owned part only: the halo tail belongs to the neighbours, who may already be storing the
  // values of their next apply into it (they can run ahead of this rank's host code); it is never read
  // before an exchange has filled it.
  SB_CUDA(cudaMemsetAsync(d, 0, sizeof(double) * (size_t) pad_up((int64_t) n > 0 ? (int64_t) n : 1), ctx->stream));
  *out = d;
  return SB_OK;
}

int vec_free(sb_ctx* ctx, double* d) {
  if (d == nullptr) return SB_OK;
  if (ctx->comm.mode < 0) {
    auto it = ctx->vec_cap.find(d);
    if (it == ctx->vec_cap.end()) {
      set_error("sb_vec_free: pointer was not allocated by sb_vec_alloc on this context (or was freed twice)");
      return SB_ERR_INVALID;
    }
    const int64_t bytes = (int64_t) sizeof(double) * it->second;
    for (auto range = ctx->vec_cache.equal_range(it->second); range.first != range.second; ++range.first)
      if (range.first->second == d) {
        set_error("sb_vec_free: vector freed twice");
        return SB_ERR_INVALID;
      }
    if (ctx->vec_cache_bytes + bytes <= vec_cache_limit(ctx)) {
      ctx->vec_cache.emplace(it->second, d); // kept for the next sb_vec_alloc of this size (stream-ordered reuse)
      ctx->vec_cache_bytes += bytes;
      return SB_OK;
    }
    ctx->vec_cap.erase(it);
    SB_CUDA(cudaStreamSynchronize(ctx->stream));
    SB_CUDA(cudaFree(d));
    return SB_OK;
  }
  SB_CUDA(cudaStreamSynchronize(ctx->stream));
  const unsigned char* p = reinterpret_cast<unsigned char*>(d);
  if (p < ctx->slab + kCtrlBytes || p >= ctx->slab + ctx->slab_bytes) {
    set_error("sb_vec_free: pointer does not belong to this context's vector pool");
    return SB_ERR_INVALID;
  }
  ctx->pool_free.push_back(d);
  return SB_OK;
}

int comm_teardown(sb_ctx* ctx) {
  if (ctx->comm.mode < 0) return SB_OK;
  cudaStreamSynchronize(ctx->stream);
  for (int r = 0; r < ctx->comm.world; ++r)
    if (r != ctx->comm.rank && ctx->comm.base[r] != nullptr) cudaIpcCloseMemHandle(ctx->comm.base[r]);
  if (ctx->nccl != nullptr && nccl_api() != nullptr) nccl_api()->CommDestroy((ncclComm_t) ctx->nccl);
  ctx->nccl = nullptr;
  cudaFree(ctx->slab);
  cudaFree(ctx->d_ar);
  cudaFree(ctx->d_sendbuf);
  ctx->slab = nullptr, ctx->d_ar = nullptr, ctx->d_sendbuf = nullptr;
  ctx->pool_free.clear();
  ctx->comm = CommDev{};
  return SB_OK;
}

__global__ void init_ctrl_kernel(CommCtrl* c) {
  const int t = threadIdx.x;
  if (t < 2 * kMaxRanks * 4) (&c->ar_slot[0][0][0])[t] = kArSentinel;
}

// ---- halo exchange: called by launch_apply before the apply kernel of a distributed operator -----------
int halo_exchange(sb_ctx* ctx, const sb_op* op, const double* x, const int* done, int64_t* x_off) {
  const HaloDev& h = op->halo;
  const int64_t total = h.send_ptr[h.n_nbr];
  const unsigned char* xb = reinterpret_cast<const unsigned char*>(x);
  if (xb < ctx->slab + kCtrlBytes || xb >= ctx->slab + ctx->slab_bytes) {
    set_error("distributed apply: x is not a vector of this context (its halo tail must live in the shared slab)");
    return SB_ERR_INVALID;
  }
  *x_off = (int64_t) (xb - ctx->slab);
  if (ctx->comm.mode == SB_COMM_P2P) return SB_OK; // the exchange is fused into the apply kernel (halo_pack_role)
  // NCCL: pack -> grouped send/recv straight into the halo tail of x
  NcclApi* api = nccl_api();
  if (total > ctx->sendbuf_cap) {
    SB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->d_sendbuf);
    ctx->sendbuf_cap = total + total / 4 + 1024;
    SB_CUDA(cudaMalloc(&ctx->d_sendbuf, sizeof(double) * ctx->sendbuf_cap));
  }
  const unsigned grid = (unsigned) std::max<int64_t>(1, std::min<int64_t>(2 * ctx->sm_count, (total + 4 * kThreads - 1) / (4 * kThreads)));
  SB_CUDA(launch_kernel(ctx, halo_pack_local_kernel, grid, kThreads, 0, h, x, ctx->d_sendbuf, done));
  ctx->launches++;
  ncclComm_t comm = (ncclComm_t) ctx->nccl;
  double* tail = const_cast<double*>(x) + op->halo_base;
  SB_NCCL(api->GroupStart());
  for (int k = 0; k < h.n_nbr; ++k) {
    SB_NCCL(api->Send(ctx->d_sendbuf + h.send_ptr[k], (size_t) (h.send_ptr[k + 1] - h.send_ptr[k]), ncclFloat64, h.nbr_rank[k],
                      comm, ctx->stream));
    SB_NCCL(api->Recv(tail + op->recv_ptr[k], (size_t) (op->recv_ptr[k + 1] - op->recv_ptr[k]), ncclFloat64, h.nbr_rank[k],
                      comm, ctx->stream));
  }
  SB_NCCL(api->GroupEnd());
  return SB_OK;
}

int nccl_allreduce_sum(sb_ctx* ctx, double* d_buf, int count) {
  NcclApi* api = nccl_api();
  SB_NCCL(api->AllReduce(d_buf, d_buf, (size_t) count, ncclFloat64, ncclSum, (ncclComm_t) ctx->nccl, ctx->stream));
  return SB_OK;
}

} // namespace sb

using namespace sb;

extern "C" {

int sb_comm_prepare(sb_ctx* ctx, int rank, int world, int mode, int64_t vec_capacity, int32_t n_vectors, void* h_blob) {
  SB_REQUIRE(ctx != nullptr && h_blob != nullptr, "null argument");
  SB_REQUIRE(world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world, "rank/world out of range (max 8 ranks)");
  SB_REQUIRE(mode == SB_COMM_NCCL || mode == SB_COMM_P2P, "unknown communicator mode");
  SB_REQUIRE(vec_capacity > 0 && n_vectors > 0, "vec_capacity and n_vectors must be positive");
  if (ctx->comm.mode >= 0) {
    set_error("communicator already prepared on this context");
    return SB_ERR_STATE;
  }
  SB_REQUIRE(ctx->work.empty(), "prepare the communicator before the first solve on this context");
  SB_CUDA(cudaSetDevice(ctx->device));
  vec_cache_release(ctx); // from here on vectors come from the symmetric slab
  const int64_t cap = pad_up(vec_capacity);
  const size_t bytes = kCtrlBytes + sizeof(double) * (size_t) cap * (size_t) n_vectors;
  SB_CUDA(cudaMalloc(&ctx->slab, bytes));
  SB_CUDA(cudaMemsetAsync(ctx->slab, 0, kCtrlBytes, ctx->stream));
  init_ctrl_kernel<<<1, 64, 0, ctx->stream>>>(reinterpret_cast<CommCtrl*>(ctx->slab));
  SB_CUDA(cudaGetLastError());
  SB_CUDA(cudaMalloc(&ctx->d_ar, sizeof(double) * 8));
  SB_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->slab_bytes = bytes, ctx->vec_capacity = cap, ctx->pool_blocks = n_vectors;
  ctx->pool_free.clear();
  for (int k = n_vectors - 1; k >= 0; --k)
    ctx->pool_free.push_back(reinterpret_cast<double*>(ctx->slab + kCtrlBytes) + (size_t) k * (size_t) cap);
  ctx->comm = CommDev{};
  ctx->comm.rank = rank, ctx->comm.world = world, ctx->comm.mode = mode;
  ctx->comm.timeout_ns = ctx->spin_timeout_ns;
  ctx->comm.base[rank] = ctx->slab;
  Blob b{};
  b.rank = rank, b.world = world, b.mode = mode, b.device = ctx->device;
  b.slab_bytes = (int64_t) bytes, b.vec_capacity = cap, b.n_vectors = n_vectors;
  if (mode == SB_COMM_P2P && world > 1) SB_CUDA(cudaIpcGetMemHandle(&b.ipc, ctx->slab));
  if (mode == SB_COMM_NCCL && world > 1) {
    if (nccl_api() == nullptr) {
      set_error("libnccl.so.2 could not be loaded (%s)", dlerror());
      return SB_ERR_NCCL;
    }
    if (rank == 0) SB_NCCL(nccl_api()->GetUniqueId(&b.nccl_id));
  }
  std::memcpy(h_blob, &b, sizeof(b));
  return SB_OK;
}

int sb_comm_connect(sb_ctx* ctx, const void* h_all_blobs) {
  SB_REQUIRE(ctx != nullptr && h_all_blobs != nullptr, "null argument");
  if (ctx->comm.mode < 0) {
    set_error("sb_comm_connect before sb_comm_prepare");
    return SB_ERR_STATE;
  }
  SB_CUDA(cudaSetDevice(ctx->device));
  const Blob* blobs = static_cast<const Blob*>(h_all_blobs);
  const int world = ctx->comm.world, rank = ctx->comm.rank;
  for (int r = 0; r < world; ++r) {
    SB_REQUIRE(blobs[r].rank == r && blobs[r].world == world && blobs[r].mode == ctx->comm.mode, "inconsistent rendezvous blobs");
    SB_REQUIRE(blobs[r].slab_bytes == (int64_t) ctx->slab_bytes && blobs[r].vec_capacity == ctx->vec_capacity,
               "ranks prepared slabs of different shape (vec_capacity and n_vectors must agree)");
  }
  if (world == 1) return SB_OK;
  if (ctx->comm.mode == SB_COMM_P2P) {
    for (int r = 0; r < world; ++r) {
      if (r == rank) continue;
      void* p = nullptr;
      SB_CUDA(cudaIpcOpenMemHandle(&p, blobs[r].ipc, cudaIpcMemLazyEnablePeerAccess));
      ctx->comm.base[r] = static_cast<unsigned char*>(p);
    }
  } else {
    ncclComm_t comm = nullptr;
    SB_NCCL(nccl_api()->CommInitRank(&comm, world, blobs[0].nccl_id, rank));
    ctx->nccl = comm;
  }
  return SB_OK;
}

int sb_comm_destroy(sb_ctx* ctx) {
  SB_REQUIRE(ctx != nullptr, "ctx is null");
  // solver workspaces live in the pool: give them back first
  for (double* w : ctx->work) vec_free(ctx, w);
  ctx->work.clear(), ctx->work_n = 0;
  for (double* w : ctx->basis) vec_free(ctx, w);
  ctx->basis.clear(), ctx->basis_n = 0;
  return comm_teardown(ctx);
}

int sb_comm_status(sb_ctx* ctx, uint64_t* h_error) {
  SB_REQUIRE(ctx != nullptr && h_error != nullptr, "null argument");
  *h_error = 0;
  if (ctx->comm.mode < 0) return SB_OK;
  unsigned long long e = 0;
  SB_CUDA(cudaMemcpy(&e, &reinterpret_cast<CommCtrl*>(ctx->slab)->error, sizeof(e), cudaMemcpyDeviceToHost));
  *h_error = e;
  return SB_OK;
}

} // extern "C"
