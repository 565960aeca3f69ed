// sb_mega.cu -- host side of the persistent whole-solve kernel (sb_mega.cuh): occupancy, cooperative launch,
// the per-context control block, the timeline read-back.
#include "sb_mega.cuh"

#include <algorithm>
#include <atomic>

namespace sb {

static __global__ void mega_ctrl_init_kernel(MegaCtrl* mc) {
  if (threadIdx.x < 8) (&mc->box[0][0])[threadIdx.x] = kArSentinel;
}

static int ensure_mega_ctrl(sb_ctx* ctx) {
  if (ctx->d_mega != nullptr) return SB_OK;
  SB_CUDA(cudaMalloc(&ctx->d_mega, sizeof(MegaCtrl)));
  SB_CUDA(cudaMemsetAsync(ctx->d_mega, 0, sizeof(MegaCtrl), ctx->stream));
  mega_ctrl_init_kernel<<<1, 32, 0, ctx->stream>>>(ctx->d_mega);
  SB_CUDA(cudaGetLastError());
  return SB_OK;
}

bool mega_supported(const sb_ctx* ctx, const sb_op* op) {
  if (op->d.form != SB_FORM_COEF || op->d.blk == nullptr) return false; // the TMA-staged blocked layout only
  if (ctx->comm.world > 1 && ctx->comm.mode != SB_COMM_P2P) return false; // the all-reduce runs inside the kernel
  if (ctx->comm.world > 1 && !op->distributed) return false;              // reductions of this context span the ranks
  if (ctx->debug & 6) return false;                                       // attribution experiments of the stepwise path
  switch (op->d.width) {
    case 0: case 1: case 2: case 3: case 4: case 5: case 6: case 7: case 8: case 10: case 12: case 14: case 16: return true;
    default: return false;
  }
}

template<int KIND, int W>
static int launch_one(sb_ctx* ctx, const MegaArgs& args) {
  auto kern = krylov_persistent_kernel<KIND, W>;
  constexpr int smem = MegaRing<W>::cta_bytes;
  static std::atomic<uint64_t> configured{0};
  static std::atomic<int> ctas_per_sm{0};
  const uint64_t bit = 1ull << (ctx->device & 63);
  if (!(configured.load(std::memory_order_acquire) & bit)) {
    SB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int nb = 0;
    SB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kThreads, smem));
    if (nb < 1) {
      set_error("persistent kernel does not fit an SM (width %d)", W);
      return SB_ERR_CUDA;
    }
    ctas_per_sm.store(nb, std::memory_order_release);
    configured.fetch_or(bit, std::memory_order_release);
  }
  // Grid: as many CTAs as fit the device at once (cooperative launch), shrunk so that the tiles divide evenly over
  // them: R rounds of `grid` tiles, the last one short by less than R tiles.
  const int64_t resident = (int64_t) ctas_per_sm.load(std::memory_order_acquire) * ctx->sm_count;
  const int64_t tiles = std::max<int64_t>(1, num_tiles(args.op.n));
  const int64_t rounds = (tiles + resident - 1) / resident;
  const unsigned grid = (unsigned) ((tiles + rounds - 1) / rounds);
  SB_CUDA(cudaMemsetAsync(&ctx->d_mega->arrive, 0, sizeof(unsigned long long), ctx->stream));
  SB_CUDA(cudaMemsetAsync(&ctx->d_mega->release, 0, sizeof(unsigned long long), ctx->stream));
  SB_CUDA(cudaMemsetAsync(ctx->d_mega->dyn, 0, sizeof(ctx->d_mega->dyn), ctx->stream));
  void* kargs[] = {const_cast<MegaArgs*>(&args)};
  SB_CUDA(cudaLaunchCooperativeKernel((const void*) kern, dim3(grid), dim3(kThreads), kargs, (size_t) smem, ctx->stream));
  ctx->launches++;
  return SB_OK;
}

int launch_mega(sb_ctx* ctx, const sb_op* op, const MegaLaunch& L) {
  SB_TRY(ensure_mega_ctrl(ctx));
  SB_TRY(ensure_red_scratch(ctx, op->d.n));
  MegaArgs a{};
  a.op = op->d;
  a.x = L.x, a.r = L.r, a.p = L.p, a.v = L.v, a.t = L.t, a.rt = L.rt;
  a.blk = L.blk, a.hist = L.hist, a.trace = L.trace;
  a.red = RedPtrs{ctx->red.partials, ctx->red.cap_tiles};
  a.mc = ctx->d_mega;
  a.timeout_ns = ctx->spin_timeout_ns;
  a.ad.comm.world = 1, a.ad.comm.rank = 0;
  if (op->distributed && ctx->comm.world > 1) {
    a.ad.comm = ctx->comm, a.ad.halo = op->halo;
    a.ad.n_pack = op->halo.n_nbr > 0 ? 1 : 0;
    // the apply inputs (p, r) must be vectors of the symmetric slab: their halo tails are written by the peers
    for (const double* v : {L.p, L.r}) {
      const unsigned char* b = reinterpret_cast<const unsigned char*>(v);
      if (b < ctx->slab + kCtrlBytes || b >= ctx->slab + ctx->slab_bytes) {
        set_error("distributed solve: workspace is not a vector of this context's slab");
        return SB_ERR_INVALID;
      }
    }
    a.off_p = reinterpret_cast<const unsigned char*>(L.p) - ctx->slab;
    a.off_r = reinterpret_cast<const unsigned char*>(L.r) - ctx->slab;
  }
  a.timeline = nullptr, a.timeline_iters = 0;
  if (L.timeline_iters > 0) {
    const int64_t words = (int64_t) L.timeline_iters * kMegaStamps;
    if (words > ctx->timeline_cap) {
      SB_CUDA(cudaStreamSynchronize(ctx->stream));
      cudaFree(ctx->d_timeline);
      ctx->d_timeline = nullptr, ctx->timeline_cap = 0;
      SB_CUDA(cudaMalloc(&ctx->d_timeline, sizeof(unsigned long long) * words));
      ctx->timeline_cap = words;
    }
    SB_CUDA(cudaMemsetAsync(ctx->d_timeline, 0, sizeof(unsigned long long) * words, ctx->stream));
    a.timeline = ctx->d_timeline, a.timeline_iters = L.timeline_iters;
  }
#define SB_MEGA_W(W)                                                              \
  case W:                                                                         \
    return L.kind == Kind::Cg ? launch_one<(int) Kind::Cg, W>(ctx, a) : launch_one<(int) Kind::BiCgStab, W>(ctx, a);
  switch (std::max(1, a.op.width)) {
    SB_MEGA_W(1) SB_MEGA_W(2) SB_MEGA_W(3) SB_MEGA_W(4) SB_MEGA_W(5) SB_MEGA_W(6) SB_MEGA_W(7) SB_MEGA_W(8)
    SB_MEGA_W(10) SB_MEGA_W(12) SB_MEGA_W(14) SB_MEGA_W(16)
    default:
      set_error("operator width %d not supported by the persistent kernel", a.op.width);
      return SB_ERR_INVALID;
  }
#undef SB_MEGA_W
}

// Host-visible failure of the last persistent launch (0: none). Call after the stream has been drained.
int mega_status(sb_ctx* ctx, unsigned long long* code) {
  *code = 0;
  if (ctx->d_mega == nullptr) return SB_OK;
  SB_CUDA(cudaMemcpy(code, &ctx->d_mega->abort, sizeof(*code), cudaMemcpyDeviceToHost));
  return SB_OK;
}

} // namespace sb
