// sb_common.cuh -- context, error handling and layout constants shared by the library TUs.
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <unordered_map>
#include <type_traits>
#include <vector>

#include "../../include/stormb200.h"

namespace sb {

// ---- layout constants ("SB_TREE v1") ---------------------------------------------
// A CTA tile is 2048 consecutive elements / rows: 8 warps x 4 sub-iterations x 32 lanes x 2.
// Vector capacities and the ELL leading dimension are padded to a multiple of kTile so every
// lane can use unguarded 128-bit accesses; only reductions mask by the logical length.
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kSub = 4;
constexpr int kTile = kThreads * 2 * kSub; // 2048
constexpr int kMaxDots = 3;
constexpr int32_t kColPad = INT32_MIN;

__host__ __device__ inline int64_t pad_up(int64_t n) { return ((n + kTile - 1) / kTile) * kTile; }
__host__ __device__ inline int64_t num_tiles(int64_t n) { return (n + kTile - 1) / kTile; }

// ---- errors ----------------------------------------------------------------------------------
void set_error(const char* fmt, ...);

#define SB_CUDA(expr)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      ::sb::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return SB_ERR_CUDA;                                                                  \
    }                                                                                      \
  } while (0)

#define SB_REQUIRE(cond, msg)                                                   \
  do {                                                                          \
    if (!(cond)) {                                                              \
      ::sb::set_error("%s:%d: invalid argument: %s", __FILE__, __LINE__, msg);  \
      return SB_ERR_INVALID;                                                    \
    }                                                                           \
  } while (0)

#define SB_TRY(expr)        \
  do {                      \
    int _rc = (expr);       \
    if (_rc != SB_OK) return _rc; \
  } while (0)

// ---- multi-GPU communicator (sb_comm.cu) ---------------------------------------------------------
// One process per GPU. Every rank allocates ONE slab: a control block followed by a pool of
// equally sized vector blocks. In P2P mode the slab is exported with CUDA IPC and mapped by every
// peer, so a vector at offset o in my slab lives at offset o in each peer's slab ("symmetric" as long
// as all ranks allocate in the same order, which SPMD solver code does) and kernels can store
// straight into a neighbour's halo tail over NVLink.
constexpr int kMaxRanks = 8;
constexpr size_t kCtrlBytes = 64 * 1024;
constexpr unsigned long long kArSentinel = 0xFFF8DEADBEEF0001ull; // a NaN payload arithmetic never produces

// Control block at the start of every rank's slab (8-byte words; written by peers over NVLink).
struct CommCtrl {
  unsigned long long halo_flag[kMaxRanks]; // [src]: halo values of apply #seq from rank src have landed
  unsigned long long ack_flag[kMaxRanks];  // [src]: rank src has finished reading the halos of apply #seq
  unsigned long long apply_seq;            // distributed applies completed by this rank (bumped by the one-CTA kernel
                                           // that follows each apply, so it is stable while an apply kernel runs)
  unsigned long long ar_seq;               // all-reduces completed so far by this rank
  unsigned long long pack_ticket;          // last-CTA detection of the pack kernel
  unsigned long long error;                // set when a spin wait gave up (code | waited-for rank); once non-zero every
                                           // later wait of this rank returns at once, so queued kernels drain, and the
                                           // host reports SB_ERR_COMM (sb_comm_status, end of every fused solve)
  unsigned long long push_ticket;          // last-tile detection of a pushing producer kernel (halo_push_tile)
  unsigned long long pad[3];
  unsigned long long ar_slot[2][kMaxRanks][4]; // all-reduce mailboxes (value is the flag), by seq parity
};
static_assert(sizeof(CommCtrl) <= kCtrlBytes, "control block too large");

// What kernels need to reach the peers (passed by value).
struct CommDev {
  int32_t rank = 0, world = 1, mode = -1; // mode: -1 none, SB_COMM_NCCL, SB_COMM_P2P
  int32_t pad = 0;
  unsigned long long timeout_ns = 0;      // bound of every peer spin wait (sb_ctx::spin_timeout_ns)
  unsigned char* base[kMaxRanks] = {};    // mapped slab of every rank (base[rank] = my own)
  __host__ __device__ CommCtrl* ctrl(int r) const { return reinterpret_cast<CommCtrl*>(base[r]); }
};

// Device-resident reduction scratch: per-tile partial sums and the output slots.
struct MegaCtrl; // sb_mega.cuh

struct RedScratch {
  double* partials = nullptr; // [kMaxDots][cap_tiles]
  int64_t cap_tiles = 0;
  double* result = nullptr;   // [64] result slots of the stand-alone dots
  double* slots = nullptr;    // [kMaxDots][cap_tiles] partial sums of launches with an in-kernel reducer (sb_finals.cuh):
                              // a NaN sentinel everywhere between kernels, the value is the flag
};

} // namespace sb

// State of a fused CG / BiCGStab solve (sb_solver_bodies.cuh: the functors that update it).
struct SolverState {
  double gamma, alpha, beta, rho, omega;
  double initial_err, abs_err, rel_err;
  double abs_tol, rel_tol;
  long long iteration, max_iter;
  long long n_hist, n_trace, hist_cap, trace_cap;
  int done, converged;
};

// Device block of a fused solve. The stepwise schedule keeps TWO versions of the state: a kernel that consumes a
// reduction ("folds" it, sb_kernels.cuh: fold_reduce / fold_wait) reads version v, its CTA 0 writes version v^1 and
// raises ready[v^1]; the other CTAs read version v^1 once the flag is up, so nobody ever reads a field that is being
// written; which version is current is a kernel argument. Every version has cache lines of its own (an SM that has
// pulled in a line of version v must not have pulled in a stale piece of version v^1 with it). `final_` is the state at
// the moment the stopping rule fired (what the host reads back), `done` the sticky stop flag every later kernel checks
// first.
struct alignas(256) StateSlot {
  SolverState s;
};
struct SolveBlock {
  StateSlot slot[2];
  StateSlot final_slot;
  alignas(128) int done;
  alignas(128) int ready[2]; // ready[v] != 0: version v is complete
  __host__ __device__ SolverState& ver(int v) { return slot[v].s; }
  __host__ __device__ SolverState& final_() { return final_slot.s; }
};

struct sb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  sb::RedScratch red;
  double* h_pinned = nullptr; // pinned staging (results, flags)
  size_t pinned_doubles = 0;
  int64_t launches = 0;
  int sm_count = 0;
  // solver workspaces, grown on demand and reused across solves
  std::vector<double*> work;
  size_t work_n = 0;
  struct SolveBlock* d_solve = nullptr;  // device: state versions of the fused CG / BiCGStab solve, stop flag
  double* d_hist = nullptr;
  int64_t hist_cap = 0;
  double* d_trace = nullptr;
  int64_t trace_cap = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::vector<cudaEvent_t>* prof_mid = nullptr; // profiled solve: launch_final records an event in front of every final stage
  // fused GMRES (sb_gmres.cu): cached Krylov basis, device scalars (H records, betas), pinned mirror
  std::vector<double*> basis;
  size_t basis_n = 0;
  double* d_gmres_scal = nullptr;
  size_t gmres_scal_cap = 0;
  double** d_gmres_ptrs = nullptr;
  double* h_gmres = nullptr;
  // single-GPU vector storage (sb_comm.cu: vec_alloc / vec_free): freed blocks are kept for reuse, because the
  // reference's solvers allocate their workspaces inside every solve() (SolverCg.hpp:57-59, SolverGmres.hpp: m + 1
  // basis vectors) and cudaMalloc + cudaFree of 4 GB costs ~0.4 s per solve at 10 M cells. Reuse is ordered by the
  // context's stream (the zero-fill of the next owner is enqueued behind every kernel of the previous one).
  std::unordered_map<double*, int64_t> vec_cap;            // live + cached blocks -> capacity in doubles
  std::unordered_multimap<int64_t, double*> vec_cache;     // capacity -> cached block
  int64_t vec_cache_bytes = 0, vec_cache_limit = -1;       // limit < 0: not initialised yet
  // multi-GPU (sb_comm.cu); comm.mode < 0: single GPU
  sb::CommDev comm;
  unsigned char* slab = nullptr;
  size_t slab_bytes = 0;
  int64_t vec_capacity = 0; // doubles per pool block
  std::vector<double*> pool_free; // LIFO free list of pool blocks
  int32_t pool_blocks = 0;
  void* nccl = nullptr;           // ncclComm_t
  double* d_ar = nullptr;         // NCCL mode: all-reduce staging [4]
  double* d_sendbuf = nullptr;    // NCCL mode: packed halo values
  int64_t sendbuf_cap = 0;
  // experiments only (SB_DEBUG env): bit1 = skip the halo exchange, bit2 = skip the cross-rank all-reduce
  // (results are wrong on purpose; used to attribute multi-GPU time, never set in tests or bench lines)
  int stream_operator = 1; // apply kernels fetch the operator slices with an L2 evict-first policy (SB_TUNE_STREAM_OPERATOR;
                           // a fused solve sets it from its tuning bits and restores it)
  int debug = 0;
  // persistent whole-solve kernel (sb_mega.cu): grid barrier / mailbox block, timeline scratch
  struct sb::MegaCtrl* d_mega = nullptr;
  unsigned long long* d_timeline = nullptr;
  int64_t timeline_cap = 0;
  unsigned long long spin_timeout_ns = 0; // bound of every device-side spin wait (SB_SPIN_TIMEOUT_S, default 120 s)
  bool pdl_launch = false; // the next launch_kernel carries the programmatic-serialization attribute (fused solvers:
                           // sb_solver_opts::tuning, SB_TUNE_PDL_*); every kernel of the library starts with
                           // griddepcontrol.launch_dependents + griddepcontrol.wait, so any launch may carry it
  uint32_t tuning = 0;     // SB_TUNE_* bits of the solve in progress (read by the launch helpers)
};

namespace sb {
// Launch on the context's stream.
template<class... KArgs, class... Args>
inline cudaError_t launch_kernel(sb_ctx* ctx, void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid), cfg.blockDim = dim3(block), cfg.dynamicSmemBytes = smem, cfg.stream = ctx->stream;
  cudaLaunchAttribute attr{};
  if (ctx->pdl_launch) { // programmatic dependent launch: resident early, blocks in griddepcontrol.wait (sb_comm.cuh)
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr, cfg.numAttrs = 1;
  }
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Launches made while one of these is alive carry the programmatic-serialization attribute (if `on`).
struct PdlScope {
  sb_ctx* ctx;
  PdlScope(sb_ctx* c, bool on) : ctx(c) { ctx->pdl_launch = on; }
  ~PdlScope() { ctx->pdl_launch = false; }
  PdlScope(const PdlScope&) = delete;
  PdlScope& operator=(const PdlScope&) = delete;
};

int ensure_red_scratch(sb_ctx* ctx, int64_t n);
// vector storage: pool block in multi-GPU mode, cudaMalloc otherwise (zero-filled either way)
int vec_alloc(sb_ctx* ctx, size_t n, double** out);
int vec_free(sb_ctx* ctx, double* d);
void vec_cache_release(sb_ctx* ctx); // cudaFree every cached block (context teardown, out-of-memory retry)
int comm_teardown(sb_ctx* ctx);
} // namespace sb
