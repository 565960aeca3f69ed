// sb_solvers.cu -- fused CG and BiCGStab (the two solvers BASELINE.json's targets are quoted on).
//
// Reference: CgSolver (source/Storm/Solvers/SolverCg.hpp:54-126) and BiCgStabSolver
// (SolverBiCgStab.hpp:59-165), driven as IterativeSolver::solve (Solver.hpp:116-147), no
// preconditioner. Every statement keeps the reference's per-element operation order, so with the
// same reduction tree the iterates are bit-identical to the CPU restatement (oracle, ORC_RED_TREE).
//
// What is different is the schedule (SURVEY.md a13/a14 "fused minimum"):
//   * scalars (alpha, beta, rho, omega, gamma, residual, iteration, stop flag) live in a device
//     struct, so an iteration needs no host sync;
//   * every dot/norm is fused into the kernel that produces its operand (apply or update), and its last stage
//     -- SB_TREE's final sum over the tile partials, the all-reduce over the ranks, the scalar update -- is
//     folded into the kernel that consumes the result (sb_kernels.cuh: fold_prologue): no one-CTA kernels
//     between the steps (they remain for the initialisation and for NCCL mode);
//   * BiCGStab's `x += alpha*p` is deferred into the final update kernel, which evaluates
//     x = (x + alpha*p) + omega*r with the same two roundings per term as the reference;
//   * once the device-side stop flag is set, the remaining queued kernels return immediately, so the
//     solution is exactly the reference's iterate at the stopping iteration.
// CG: 3 launches / iteration, 9 vector passes + 1 apply. BiCGStab: 5 launches, 15 passes + 2 applies.
#include "sb_solver_bodies.cuh"

#include <cmath>
#include <string>

namespace sb {

// ---- host side ------------------------------------------------------------------------------------
template<int ND, class Body, class Final>
int launch_ew(sb_ctx* ctx, int64_t n, const Body& body, const Final& fin, const int* done, bool pdl = false,
              bool pdl_final = false, unsigned long long* ar_wait_ns = nullptr) {
  RedPtrs red{ctx->red.partials, ctx->red.cap_tiles};
  {
    PdlScope pdl_scope(ctx, pdl);
    SB_CUDA(launch_kernel(ctx, ew_kernel<ND, Body>, (unsigned) num_tiles(n), kThreads, 0, n, body, red, done));
  }
  ctx->launches++;
  if constexpr (ND > 0) return launch_final<ND>(ctx, n, fin, done, nullptr, ar_wait_ns, pdl_final);
  return SB_OK;
}

// An element-wise step of the stepwise schedule (ew_solver_kernel). `push_y` != null: the step's output is the input of
// the next apply and the kernel forwards its boundary values to the neighbours itself (multi-GPU, in-kernel
// collectives). `ra` (ND > 0): the reductions are finished by the kernel's last CTA; without it a one-CTA final stage
// follows.
template<int ND, class Body, class Final>
int launch_ew_solver(sb_ctx* ctx, const sb_op* op, int64_t n, const Body& body, const double* push_y, const Final& fin,
                     const int* done, bool pdl, bool pdl_final, const ReducerArgs* reducer, unsigned long long* ar_wait_ns) {
  ReducerArgs ra;
  if (reducer != nullptr) ra = *reducer, ra.n_tiles = num_tiles(n);
  const bool in_kernel = ND > 0 && ra.kind != kFinalNone;
  if (!in_kernel) ra = ReducerArgs{};
  const RedPtrs red = in_kernel ? RedPtrs{ra.slots, ra.cap_tiles} : RedPtrs{ctx->red.partials, ctx->red.cap_tiles};
  PushArgs pa;
  pa.n_tiles = num_tiles(n);
  if (push_y != nullptr) {
    pa.comm = ctx->comm, pa.halo = op->halo, pa.y = push_y, pa.push = 1, pa.lazy = (ctx->tuning & SB_TUNE_PUSH_LAZY) ? 1 : 0;
    pa.y_off = (int64_t) (reinterpret_cast<const unsigned char*>(push_y) - ctx->slab);
  }
  {
    PdlScope pdl_scope(ctx, pdl);
    if (in_kernel)
      SB_CUDA(launch_kernel(ctx, ew_solver_kernel<ND, Body, true>, (unsigned) (num_tiles(n) + 1), kThreads, 0, n, body, red, done, pa, ra));
    else
      SB_CUDA(launch_kernel(ctx, ew_solver_kernel<ND, Body, false>, (unsigned) num_tiles(n), kThreads, 0, n, body, red, done, pa, ra));
  }
  ctx->launches++;
  if constexpr (ND > 0) {
    if (!in_kernel) return launch_final<ND>(ctx, n, fin, done, nullptr, ar_wait_ns, pdl_final);
  }
  return SB_OK;
}

static int ensure_work(sb_ctx* ctx, size_t n, size_t count) {
  if (ctx->work_n < n) {
    SB_CUDA(cudaStreamSynchronize(ctx->stream));
    for (double* w : ctx->work) SB_TRY(vec_free(ctx, w));
    ctx->work.clear();
    ctx->work_n = n;
  }
  while (ctx->work.size() < count) {
    double* d = nullptr;
    SB_TRY(vec_alloc(ctx, ctx->work_n, &d)); // pool block in multi-GPU mode (apply inputs need a halo tail)
    ctx->work.push_back(d);
  }
  return SB_OK;
}

static int ensure_records(sb_ctx* ctx, int64_t hist_cap, int64_t trace_cap) {
  if (ctx->d_solve == nullptr) SB_CUDA(cudaMalloc(&ctx->d_solve, sizeof(SolveBlock)));
  if (hist_cap > ctx->hist_cap) {
    SB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->d_hist);
    SB_CUDA(cudaMalloc(&ctx->d_hist, sizeof(double) * hist_cap));
    ctx->hist_cap = hist_cap;
  }
  if (trace_cap > ctx->trace_cap) {
    SB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->d_trace);
    SB_CUDA(cudaMalloc(&ctx->d_trace, sizeof(double) * trace_cap));
    ctx->trace_cap = trace_cap;
  }
  return SB_OK;
}

int mega_status(sb_ctx* ctx, unsigned long long* code); // sb_mega.cu

// x and b must be sb_vec_alloc vectors of at least n elements: every kernel uses unguarded 128-bit accesses up to the
// padded length and the apply reads 512-byte runs by bulk copy.
static int check_vector(sb_ctx* ctx, const double* v, int64_t n, const char* what) {
  if (ctx->comm.mode >= 0) {
    const unsigned char* b = reinterpret_cast<const unsigned char*>(v);
    if (b < ctx->slab + kCtrlBytes || b >= ctx->slab + ctx->slab_bytes || n > ctx->vec_capacity) {
      set_error("%s is not a vector of this context's pool (or is shorter than the operator's %lld rows)", what, (long long) n);
      return SB_ERR_INVALID;
    }
    return SB_OK;
  }
  for (const auto& kv : ctx->vec_cap)
    if (v >= kv.first && v < kv.first + kv.second) {
      if (v + pad_up(n) <= kv.first + kv.second) return SB_OK;
      set_error("%s holds fewer than the operator's %lld rows (padded to %lld)", what, (long long) n, (long long) pad_up(n));
      return SB_ERR_INVALID;
    }
  set_error("%s was not allocated by sb_vec_alloc on this context", what);
  return SB_ERR_INVALID;
}

// What sb_solver_opts::tuning == 0 selects (profiles/r02_ab_*):
//  * STREAM_OPERATOR always: one GPU, 1.23 M / 2.5 M / 10.1 M cells: BiCGStab +9 % / +10 % / +-0, CG +5 % / +15 % / +-0;
//    10.1 M cells on 2 / 4 / 8 GPUs: BiCGStab +4 % / +17 % / +12 %;
//  * NO_ACK always (it only concerns the distributed apply inside a fused solve): 8 GPUs +5 %, 2 GPUs +-0;
//  * PUSH_ON_PRODUCE | PUSH_LAZY when the rank's apply kernel is at most ~two waves of tiles: there the pack chain
//    of the apply (stores, fence, ticket, fence, flags) lands 7-10 us after the boundary tiles need it (8 GPUs, 617
//    tiles per rank: +10 % on top of the two above), while with several waves in front of the boundary tiles it is
//    hidden anyway and the pushing producers only cost (4 GPUs, 1 235 tiles: 7 138 against 7 171 it/s).
// Every other bit (eager push, in-kernel reducer, programmatic dependent launch) measured slower or equal everywhere.
constexpr int64_t kLazyPushMaxTiles = 1024;
static uint32_t default_tuning(const sb_ctx* ctx, const sb_op* op) {
  uint32_t t = SB_TUNE_STREAM_OPERATOR | SB_TUNE_NO_ACK;
  if (ctx->comm.world > 1 && ctx->comm.mode == SB_COMM_P2P && op->distributed && num_tiles(op->d.n) <= kLazyPushMaxTiles)
    t |= SB_TUNE_PUSH_ON_PRODUCE | SB_TUNE_PUSH_LAZY;
  return t;
}

struct Solve {
  sb_ctx* ctx;
  const sb_op* op;
  double* x;
  const double* b;
  int64_t n;
  Recorder rec;      // rec.st = version 0 of the state (the one-CTA final stages update it in place)
  SolveBlock* blk;
  const int* done;
  double *p, *r, *z, *rt, *t, *v;
  std::vector<cudaEvent_t>* prof = nullptr; // profile=1: one event before every kernel + one per iteration end
  // folded schedule
  bool folded = false;
  int ver = 0;              // current version of the state
  bool pending_end = false; // BiCGStab: the reduction of the last final update has not been consumed yet
  unsigned long long* tl = nullptr; // optional timeline: [launch ordinal] longest wait (halo flag / other ranks' sums)
  int64_t tl_cap = 0, tl_pos = 0;

  int mark() {
    if (prof != nullptr) {
      cudaEvent_t e;
      SB_CUDA(cudaEventCreate(&e));
      prof->push_back(e); // owned by the solve's guard from here on
      SB_CUDA(cudaEventRecord(e, ctx->stream));
    }
    return SB_OK;
  }
  // two timeline words per kernel slot and iteration: [0] the slot's own in-kernel wait (apply: longest halo-flag wait
  // of a boundary CTA; folding kernel: the reducer's wait), [1] the wait of the one-CTA final stage behind the slot for
  // the other ranks' sums. Every slot takes its pair, used or not, so the words stay aligned with the slots.
  struct WaitPair {
    unsigned long long *own = nullptr, *ar = nullptr;
  };
  WaitPair wait_slot() {
    WaitPair w;
    if (tl != nullptr && tl_pos + 2 <= tl_cap) w.own = tl + tl_pos, w.ar = tl + tl_pos + 1;
    tl_pos += 2;
    return w;
  }
  RedPtrs red_set(int k) const {
    return RedPtrs{ctx->red.partials + (int64_t) k * kMaxDots * ctx->red.cap_tiles, ctx->red.cap_tiles};
  }
  bool dist() const { return op->distributed && ctx->comm.world > 1 && ctx->comm.mode == SB_COMM_P2P; }

  int init(Kind kind) {
    // r <- b - A x fused with <r,r>  (Operator::Residual, Operator.hpp:95-99)
    EpiResidual epi{b};
    if (kind == Kind::Cg) {
      SB_TRY((launch_apply<1, true>(ctx, op, x, r, epi, PublishFinal<CgInitFinal>{CgInitFinal{rec}, blk}, nullptr)));
      // p <- r; with push-on-produce this is the producer of the first apply's input (skipped, like that apply, when the
      // initial residual already met the tolerance: a push that no apply consumes would leave a stale halo flag behind)
      if (push) SB_TRY((launch_ew_solver<0>(ctx, op, n, CopyBody{p, r}, p, NoFinal{}, done, false, false, nullptr, nullptr)));
      else SB_TRY((launch_ew<0>(ctx, n, CopyBody{p, r}, NoFinal{}, nullptr)));
    } else {
      SB_TRY((launch_apply<1, true>(ctx, op, x, r, epi, PublishFinal<BiInitFinal>{BiInitFinal{rec}, blk}, nullptr)));
      SB_TRY((launch_ew<0>(ctx, n, BiInitBody{rt, r}, NoFinal{}, nullptr))); // r~ <- r
    }
    return SB_OK;
  }

  // One kernel of the folded schedule: the element-wise step `body` (ND reductions of its own, into partial set 1)
  // behind the fold of the reduction in front of it (FND sums from partial set `from_set`; `active` false: none).
  template<int ND, int FND, class Body, class Final>
  int launch_fold(const Body& body, const Final& fin, bool active, int from_set, bool after_apply, int64_t rows) {
    Fold<FND, Final> f;
    f.n_tiles = active ? num_tiles(n) : -1;
    f.red = red_set(from_set);
    f.fin = fin;
    f.blk = blk, f.in = ver;
    if (dist()) {
      f.comm = ctx->comm;
      if (after_apply && !(ctx->debug & 2)) f.bump = ctx->comm.ctrl(ctx->comm.rank);
    }
    f.wait_ns = wait_slot().own; // one pair per kernel slot, folding or not: the timeline stays aligned with the slots
    const unsigned grid = (unsigned) (num_tiles(rows) + 1); // CTA 0 reduces, CTA k > 0 owns tile k - 1
    auto kern = ew_fold_kernel<ND, Body, FND, Final>;
    constexpr int smem = EwStage<Body>::cta;
    static std::atomic<uint64_t> configured{0}; // the attribute is per device
    const uint64_t bit = 1ull << (ctx->device & 63);
    if (!(configured.load(std::memory_order_acquire) & bit)) {
      SB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      configured.fetch_or(bit, std::memory_order_release);
    }
    SB_CUDA(launch_kernel(ctx, kern, grid, kThreads, smem, rows, body, red_set(1), done, f));
    ctx->launches++;
    if (active) ver ^= 1;
    return SB_OK;
  }

  ApplyOpts folded_apply() {
    ApplyOpts ao;
    ao.fold_later = true, ao.halo_wait_ns = wait_slot().own;
    return ao;
  }

  int iterate_folded(Kind kind) {
    if (kind == Kind::Cg) {
      SB_TRY(mark());
      SB_TRY((launch_apply<1, false>(ctx, op, p, z, EpiXY{}, NoFinal{}, done, folded_apply())));
      SB_TRY(mark());
      SB_TRY((launch_fold<1, 1>(CgUpdateBody{nullptr, x, r, p, z}, CgAlphaFinal{rec}, true, 0, true, n)));
      SB_TRY(mark());
      SB_TRY((launch_fold<0, 1>(CgDirectionBody{nullptr, p, r}, CgBetaFinal{rec}, true, 1, false, n)));
    } else {
      SB_TRY(mark());
      SB_TRY((launch_fold<0, 2>(BiDirectionBody{nullptr, p, r, v}, BiEndFinal{rec}, pending_end, 1, false, n)));
      SB_TRY(mark());
      SB_TRY((launch_apply<1, false>(ctx, op, p, v, EpiUY{rt}, NoFinal{}, done, folded_apply())));
      SB_TRY(mark());
      SB_TRY((launch_fold<0, 1>(BiHalfBody{nullptr, r, v}, BiAlphaFinal{rec}, true, 0, true, n)));
      SB_TRY(mark());
      SB_TRY((launch_apply<2, false>(ctx, op, r, t, EpiYYandYX{}, NoFinal{}, done, folded_apply())));
      SB_TRY(mark());
      SB_TRY((launch_fold<2, 2>(BiEndBody{nullptr, x, r, p, t, rt}, BiOmegaFinal{rec}, true, 0, true, n)));
      pending_end = true;
    }
    SB_TRY(mark());
    return SB_OK;
  }

  // After the last iteration: consume what is still pending (BiCGStab's last <r,r>, <r~,r>).
  int finish_folded(Kind kind) {
    if (kind == Kind::BiCgStab && pending_end)
      SB_TRY((launch_fold<0, 2>(BiDirectionBody{nullptr, p, r, v}, BiEndFinal{rec}, true, 1, false, 0)));
    return SB_OK;
  }

  // Stepwise schedule: one kernel per step. Tuning (SB_TUNE_*):
  // `in_kernel`: every reduction is finished by the last CTA of the kernel that produces it (sb_finals.cuh); without
  //   it a one-CTA final stage follows every reducing kernel;
  // `push`: the producers of the apply inputs (p; BiCGStab also r) forward the boundary values themselves;
  // `pdl_final` / `pdl_after` / `pdl_apply`: programmatic-serialization attribute on the final stages / on the kernels
  //   behind a reduction / on the applies.
  bool push = false, no_ack = false, pdl_final = false, pdl_after = false, pdl_apply = false, in_kernel = false;

  ReducerArgs reducer(FinalKind kind, int nd, unsigned long long* wait_ns) const {
    ReducerArgs ra;
    if (!in_kernel) return ra;
    ra.kind = kind, ra.nd = nd;
    ra.slots = ctx->red.slots, ra.cap_tiles = ctx->red.cap_tiles;
    ra.rec = rec, ra.blk = blk;
    if (dist() && !(ctx->debug & 4)) ra.comm = ctx->comm;
    ra.comm.timeout_ns = ctx->spin_timeout_ns;
    ra.wait_ns = wait_ns;
    return ra;
  }

  template<int ND, bool RESID, class Epi, class Final>
  int stepwise_apply(const double* in, double* out, const Epi& epi, const Final& fin, FinalKind kind) {
    ApplyOpts ao;
    ao.halo_mode = push ? ((ctx->tuning & SB_TUNE_PUSH_LAZY) ? 3 : 2) : (no_ack ? 1 : 0);
    ao.pdl = pdl_apply, ao.pdl_final = pdl_final;
    const WaitPair w = wait_slot();
    ao.halo_wait_ns = w.own, ao.ar_wait_ns = w.ar;
    (void) kind; // an apply's reduction always ends in the one-CTA final stage (sb_op.cuh: apply_kernel_tma)
    return launch_apply<ND, RESID>(ctx, op, in, out, epi, fin, done, ao);
  }

  template<int ND, class Body, class Final>
  int stepwise_ew(const Body& body, const double* push_y, const Final& fin, FinalKind kind) {
    const WaitPair w = wait_slot();
    const ReducerArgs ra = reducer(kind, ND, w.ar);
    return launch_ew_solver<ND>(ctx, op, n, body, push ? push_y : nullptr, fin, done, pdl_after, pdl_final,
                                (ND > 0 && in_kernel) ? &ra : nullptr, w.ar);
  }

  int iterate(Kind kind) {
    if (folded) return iterate_folded(kind);
    const SolverState* st = rec.st;
    if (kind == Kind::Cg) {
      SB_TRY(mark());
      SB_TRY((stepwise_apply<1, false>(p, z, EpiXY{}, PublishFinal<CgAlphaFinal>{CgAlphaFinal{rec}, blk}, kFinalCgAlpha)));
      SB_TRY(mark());
      SB_TRY((stepwise_ew<1>(CgUpdateBody{st, x, r, p, z}, nullptr, PublishFinal<CgBetaFinal>{CgBetaFinal{rec}, blk}, kFinalCgBeta)));
      SB_TRY(mark());
      SB_TRY((stepwise_ew<0>(CgDirectionBody{st, p, r}, p, NoFinal{}, kFinalNone)));
    } else {
      SB_TRY(mark());
      SB_TRY((stepwise_ew<0>(BiDirectionBody{st, p, r, v}, p, NoFinal{}, kFinalNone)));
      SB_TRY(mark());
      SB_TRY((stepwise_apply<1, false>(p, v, EpiUY{rt}, PublishFinal<BiAlphaFinal>{BiAlphaFinal{rec}, blk}, kFinalBiAlpha)));
      SB_TRY(mark());
      SB_TRY((stepwise_ew<0>(BiHalfBody{st, r, v}, r, NoFinal{}, kFinalNone)));
      SB_TRY(mark());
      SB_TRY((stepwise_apply<2, false>(r, t, EpiYYandYX{}, PublishFinal<BiOmegaFinal>{BiOmegaFinal{rec}, blk}, kFinalBiOmega)));
      SB_TRY(mark());
      SB_TRY((stepwise_ew<2>(BiEndBody{st, x, r, p, t, rt}, nullptr, PublishFinal<BiEndFinal>{BiEndFinal{rec}, blk}, kFinalBiEnd)));
    }
    SB_TRY(mark());
    return SB_OK;
  }
};

// Destroys the CUDA objects of a solve on every return path.
struct SolveGuard {
  std::vector<cudaEvent_t> events;
  cudaGraphExec_t graph_exec = nullptr;
  cudaGraph_t graph = nullptr;
  int make(cudaEvent_t* e, unsigned flags = cudaEventDefault) {
    SB_CUDA(cudaEventCreateWithFlags(e, flags));
    events.push_back(*e);
    return SB_OK;
  }
  std::vector<cudaEvent_t> mid_events; // profiled solve: one in front of every one-CTA final stage
  sb_ctx* ctx = nullptr;
  ~SolveGuard() {
    if (ctx != nullptr) ctx->prof_mid = nullptr;
    for (cudaEvent_t e : mid_events) cudaEventDestroy(e);
    for (cudaEvent_t e : events) cudaEventDestroy(e);
    if (graph_exec != nullptr) cudaGraphExecDestroy(graph_exec);
    if (graph != nullptr) cudaGraphDestroy(graph);
  }
};

static int run_solver(sb_ctx* ctx, const sb_op* op, Kind kind, double* x, const double* b,
                      const sb_solver_opts* opts, sb_solver_report* report, double* h_hist, int64_t hist_cap,
                      double* h_trace, int64_t trace_cap) {
  SB_REQUIRE(ctx != nullptr && op != nullptr && x != nullptr && b != nullptr && opts != nullptr && report != nullptr,
             "null argument");
  SB_REQUIRE(x != b, "x and b must not alias");
  SB_REQUIRE(opts->num_iterations >= 0, "num_iterations must be >= 0");
  SB_REQUIRE(hist_cap >= 0 && trace_cap >= 0, "negative capacity");
  SB_REQUIRE(opts->schedule >= SB_SCHEDULE_AUTO && opts->schedule <= SB_SCHEDULE_FOLDED, "unknown schedule");
  SB_REQUIRE(opts->timeline_iters >= 0 && (opts->timeline_iters == 0 || opts->h_timeline != nullptr), "timeline buffer");
  const int64_t n = op->d.n;
  SB_CUDA(cudaSetDevice(ctx->device));
  SB_TRY(check_vector(ctx, x, n, "x"));
  SB_TRY(check_vector(ctx, b, n, "b"));
  SB_TRY(ensure_work(ctx, (size_t) n, kind == Kind::Cg ? 3 : 5));
  SB_TRY(ensure_records(ctx, h_hist ? hist_cap : 0, h_trace ? trace_cap : 0));
  SB_TRY(ensure_red_scratch(ctx, n));
  const bool profile = opts->profile != 0 && !opts->use_graph;
  // Schedule: stepwise (one kernel per step + a one-CTA stage per reduction, graph replay) unless the caller asks for
  // one of the other two. Measured on the B200 (profiles/r02_*): a grid-wide barrier in global memory
  // costs 5-7 us (the arriving CTA must first drain its stores) against 2.7 us for a kernel boundary inside a replayed
  // graph, and a folded reduction makes every resident CTA of the consumer wait for a reducer that runs on a loaded
  // machine (4-9 us) instead of on an idle one (3 us + one boundary).
  bool persistent = opts->schedule == SB_SCHEDULE_PERSISTENT && !profile && mega_supported(ctx, op);
  if (opts->schedule == SB_SCHEDULE_PERSISTENT && !persistent) {
    set_error("the persistent schedule needs a coefficient-form operator (blocked layout), no per-kernel profile, and "
              "in-kernel (P2P) collectives");
    return SB_ERR_INVALID;
  }
  const bool can_fold = !(ctx->debug & 4) && (ctx->comm.world <= 1 || (ctx->comm.mode == SB_COMM_P2P && op->distributed));
  if (opts->schedule == SB_SCHEDULE_FOLDED && !can_fold) {
    set_error("the folded schedule needs in-kernel (P2P) collectives over a distributed operator, or one GPU");
    return SB_ERR_INVALID;
  }

  SolveBlock* pinned_blk = reinterpret_cast<SolveBlock*>(ctx->h_pinned);
  *pinned_blk = SolveBlock{};
  SolverState& h = pinned_blk->ver(0);
  h.abs_tol = opts->abs_tol, h.rel_tol = opts->rel_tol;
  h.max_iter = opts->num_iterations;
  h.hist_cap = h_hist ? hist_cap : 0, h.trace_cap = h_trace ? trace_cap : 0;
  SB_CUDA(cudaMemcpyAsync(ctx->d_solve, pinned_blk, sizeof(SolveBlock), cudaMemcpyHostToDevice, ctx->stream));

  Solve S;
  S.ctx = ctx, S.op = op, S.x = x, S.b = b, S.n = n;
  S.blk = ctx->d_solve;
  S.rec = Recorder{&ctx->d_solve->ver(0), h_hist ? ctx->d_hist : nullptr, h_trace ? ctx->d_trace : nullptr};
  S.done = &ctx->d_solve->done;
  S.p = ctx->work[0], S.r = ctx->work[1];
  S.z = S.rt = S.t = S.v = nullptr;
  if (kind == Kind::Cg) {
    S.z = ctx->work[2];
  } else {
    S.rt = ctx->work[2], S.t = ctx->work[3], S.v = ctx->work[4];
  }
  S.folded = opts->schedule == SB_SCHEDULE_FOLDED;
  // tuning of the stepwise schedule (include/stormb200.h: SB_TUNE_*); 0 = the defaults
  const uint32_t tuning = opts->tuning == 0 ? default_tuning(ctx, op) : (opts->tuning & ~SB_TUNE_OFF);
  struct TuningScope { // the launch helpers read the bits from the context while this solve is in progress
    sb_ctx* c;
    ~TuningScope() { c->tuning = 0, c->stream_operator = 1; }
  } tuning_scope{ctx};
  ctx->tuning = tuning, ctx->stream_operator = (tuning & SB_TUNE_STREAM_OPERATOR) ? 1 : 0;
  const bool p2p = ctx->comm.world > 1 && ctx->comm.mode == SB_COMM_P2P && op->distributed && !(ctx->debug & 2);
  S.push = p2p && (tuning & SB_TUNE_PUSH_ON_PRODUCE) && op->halo.n_nbr > 0 && op->halo.push_ptr != nullptr && !S.folded;
  S.no_ack = p2p && (tuning & SB_TUNE_NO_ACK);
  S.pdl_final = (tuning & SB_TUNE_PDL_FINAL) != 0 && !(ctx->comm.world > 1 && ctx->comm.mode == SB_COMM_NCCL);
  S.pdl_after = (tuning & SB_TUNE_PDL_AFTER_FINAL) != 0 && !(ctx->comm.world > 1 && ctx->comm.mode == SB_COMM_NCCL);
  S.in_kernel = (tuning & SB_TUNE_IN_KERNEL_REDUCER) != 0 && !(ctx->comm.world > 1 && (ctx->comm.mode == SB_COMM_NCCL || !op->distributed));
  S.pdl_apply = (tuning & SB_TUNE_PDL_APPLY) != 0 && !(ctx->comm.world > 1 && ctx->comm.mode == SB_COMM_NCCL);
  SolveGuard guard;
  const int64_t launches0 = ctx->launches;
  SB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  SB_TRY(S.init(kind));
  cudaEvent_t ev_mid;
  SB_TRY(guard.make(&ev_mid));
  SB_CUDA(cudaEventRecord(ev_mid, ctx->stream));
  std::vector<cudaEvent_t>& prof_events = guard.events; // profile events are appended behind ev_mid / evs
  const int per_iter = (kind == Kind::Cg) ? 3 : 5;       // profiled kernel slots
  size_t prof_first = 0;
  const int64_t tl_words = profile ? 2 * (int64_t) per_iter * opts->num_iterations : 0;
  if (tl_words > 0) { // in-kernel waits of the profiled run: two words per kernel slot and iteration
    if (tl_words > ctx->timeline_cap) {
      SB_CUDA(cudaStreamSynchronize(ctx->stream));
      cudaFree(ctx->d_timeline);
      ctx->d_timeline = nullptr, ctx->timeline_cap = 0;
      SB_CUDA(cudaMalloc(&ctx->d_timeline, sizeof(unsigned long long) * tl_words));
      ctx->timeline_cap = tl_words;
    }
    SB_CUDA(cudaMemsetAsync(ctx->d_timeline, 0, sizeof(unsigned long long) * tl_words, ctx->stream));
    S.tl = ctx->d_timeline, S.tl_cap = tl_words;
  }

  if (persistent) {
    // ONE cooperative launch runs the whole iteration loop (sb_mega.cuh); the stop rule is evaluated on the device
    if (opts->num_iterations > 0) {
      MegaLaunch L{};
      L.kind = kind, L.x = x, L.r = S.r, L.p = S.p;
      L.v = kind == Kind::Cg ? S.z : S.v, L.t = S.t, L.rt = S.rt;
      L.blk = ctx->d_solve, L.hist = S.rec.hist, L.trace = S.rec.trace;
      L.timeline_iters = opts->timeline_iters;
      SB_TRY(launch_mega(ctx, op, L));
    }
  } else {
    // Graph replay: the kernel arguments never change (scalars are read from the device state), so one captured
    // graph is replayed; this removes the per-kernel launch cost that matters once a rank holds ~1 M cells. The
    // folded schedule bakes the state version into the arguments: the graph spans an even number of folds -- two
    // BiCGStab iterations (3 folds each; the first iteration, which has no reduction in front of it, runs outside
    // the graph) or one CG iteration (2 folds). Iterations queued beyond the stopping one are no-ops.
    const int its_per_graph = (S.folded && kind == Kind::BiCgStab) ? 2 : 1;
    int64_t it = 0;
    if (opts->use_graph && opts->num_iterations > 0) {
      if (its_per_graph == 2) {
        SB_TRY(S.iterate(kind));
        it = 1;
      }
      if (it < opts->num_iterations) {
        SB_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        const int64_t before = ctx->launches;
        int rc = SB_OK;
        for (int k = 0; k < its_per_graph && rc == SB_OK; ++k) rc = S.iterate(kind);
        cudaError_t ce = cudaStreamEndCapture(ctx->stream, &guard.graph);
        ctx->launches = before;
        if (rc != SB_OK) return rc;
        SB_CUDA(ce);
        SB_CUDA(cudaGraphInstantiate(&guard.graph_exec, guard.graph, 0));
      }
    }
    // + the one-CTA final stages (with the in-kernel reducer only those behind the applies remain)
    const int launches_per_iter = S.folded ? per_iter : (S.in_kernel ? per_iter + ((kind == Kind::Cg) ? 1 : 2) : ((kind == Kind::Cg) ? 5 : 8));

    // Convergence polling: a flag copy is queued every `check` iterations and examined one batch later,
    // so the host never drains the stream while it still has work to enqueue.
    const int check = opts->check_every > 0 ? opts->check_every : 32;
    int* h_flags = reinterpret_cast<int*>(ctx->h_pinned + 256);
    h_flags[0] = h_flags[1] = 0;
    cudaEvent_t evs[2];
    SB_TRY(guard.make(&evs[0], cudaEventDisableTiming));
    SB_TRY(guard.make(&evs[1], cudaEventDisableTiming));
    prof_first = guard.events.size();
    int slot = 0;
    bool pending[2] = {false, false};
    bool stop = false;
    while (it < opts->num_iterations && !stop) {
      const int64_t batch_end = std::min<int64_t>(it + check, opts->num_iterations);
      while (it < batch_end) {
        if (guard.graph_exec != nullptr) {
          SB_CUDA(cudaGraphLaunch(guard.graph_exec, ctx->stream));
          ctx->launches += launches_per_iter * its_per_graph;
          it += its_per_graph;
        } else {
          if (profile) S.prof = &prof_events, guard.ctx = ctx, ctx->prof_mid = &guard.mid_events;
          SB_TRY(S.iterate(kind));
          ++it;
        }
      }
      SB_CUDA(cudaMemcpyAsync(&h_flags[slot], &ctx->d_solve->done, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      SB_CUDA(cudaEventRecord(evs[slot], ctx->stream));
      pending[slot] = true;
      const int prev = slot ^ 1;
      if (pending[prev]) {
        SB_CUDA(cudaEventSynchronize(evs[prev]));
        pending[prev] = false;
        if (h_flags[prev] != 0) stop = true;
      }
      slot ^= 1;
    }
    S.prof = nullptr, ctx->prof_mid = nullptr;
    if (S.folded) SB_TRY(S.finish_folded(kind));
  }
  SB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  SB_CUDA(cudaMemcpyAsync(pinned_blk, ctx->d_solve, sizeof(SolveBlock), cudaMemcpyDeviceToHost, ctx->stream));
  SB_CUDA(cudaStreamSynchronize(ctx->stream));
  // the state at the moment the stopping rule fired; a solve always ends with the flag set (tolerance met, or the
  // iteration count reached num_iterations)
  const SolverState out = pinned_blk->done ? pinned_blk->final_() : pinned_blk->ver(S.ver);
  float ms = 0.f, ms_iter = 0.f;
  SB_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  SB_CUDA(cudaEventElapsedTime(&ms_iter, ev_mid, ctx->ev1));
  report->iter_ms = ms_iter;
  report->n_kernel_slots = persistent ? 0 : per_iter;
  report->schedule = persistent ? SB_SCHEDULE_PERSISTENT : (S.folded ? SB_SCHEDULE_FOLDED : SB_SCHEDULE_STEPWISE);
  for (int k = 0; k < SB_MAX_KERNEL_SLOTS; ++k) report->kernel_ms[k] = 0.0;
  for (int k = 0; k < SB_MAX_KERNEL_SLOTS; ++k) report->wait_ms[k] = report->ar_wait_ms[k] = report->final_ms[k] = 0.0;
  if (profile && !persistent) {
    // events come in groups of per_iter + 1 per iteration
    const size_t group = (size_t) per_iter + 1;
    for (size_t g = prof_first; g + group <= prof_events.size(); g += group)
      for (int k = 0; k < per_iter; ++k) {
        float dt = 0.f;
        cudaEventElapsedTime(&dt, prof_events[g + k], prof_events[g + k + 1]);
        report->kernel_ms[k] += dt;
      }
    // one event per final stage: CG slots 0, 1; BiCGStab slots 1, 3, 4 (none when the reductions end in-kernel)
    const int cg_slots[2] = {0, 1}, bi_slots[3] = {1, 3, 4};
    const int nf = kind == Kind::Cg ? 2 : 3;
    const int* fslot = kind == Kind::Cg ? cg_slots : bi_slots;
    const size_t iters_profiled = (prof_events.size() - prof_first) / group;
    if (guard.mid_events.size() == iters_profiled * (size_t) nf)
      for (size_t i = 0; i < iters_profiled; ++i)
        for (int j = 0; j < nf; ++j) {
          float dt = 0.f;
          cudaEventElapsedTime(&dt, guard.mid_events[i * nf + j], prof_events[prof_first + i * group + fslot[j] + 1]);
          report->final_ms[fslot[j]] += dt;
        }
    if (tl_words > 0) { // longest in-kernel wait per launch, summed per slot
      std::vector<unsigned long long> tl((size_t) tl_words);
      SB_CUDA(cudaMemcpy(tl.data(), ctx->d_timeline, sizeof(unsigned long long) * tl_words, cudaMemcpyDeviceToHost));
      for (int64_t q = 0; q + 1 < S.tl_pos && q + 1 < tl_words; q += 2) {
        report->wait_ms[(q / 2) % per_iter] += 1e-6 * (double) tl[(size_t) q];
        report->ar_wait_ms[(q / 2) % per_iter] += 1e-6 * (double) tl[(size_t) q + 1];
      }
    }
  }
  // a device-side spin wait that gave up (lost peer, rank skew beyond SB_SPIN_TIMEOUT_S): values are meaningless
  unsigned long long fail = 0;
  if (persistent) SB_TRY(mega_status(ctx, &fail));
  if (fail == 0 && ctx->comm.mode == SB_COMM_P2P && ctx->comm.world > 1)
    SB_CUDA(cudaMemcpy(&fail, &reinterpret_cast<CommCtrl*>(ctx->slab)->error, sizeof(fail), cudaMemcpyDeviceToHost));
  if (fail != 0) {
    set_error("a device-side wait timed out (code 0x%llx: 0xA/0xB... halo ack/flag of rank, 0xC... all-reduce value of "
              "rank, 0xD... grid barrier); the context is unusable", fail);
    return SB_ERR_COMM;
  }
  report->converged = out.converged;
  report->iterations = out.iteration;
  report->initial_err = out.initial_err;
  report->abs_err = out.abs_err;
  report->rel_err = out.rel_err;
  report->n_hist = std::min<int64_t>(out.n_hist, out.hist_cap);
  report->n_trace = std::min<int64_t>(out.n_trace, out.trace_cap);
  report->solve_ms = ms;
  report->launches = ctx->launches - launches0;
  if (h_hist && report->n_hist > 0)
    SB_CUDA(cudaMemcpy(h_hist, ctx->d_hist, sizeof(double) * report->n_hist, cudaMemcpyDeviceToHost));
  if (h_trace && report->n_trace > 0)
    SB_CUDA(cudaMemcpy(h_trace, ctx->d_trace, sizeof(double) * report->n_trace, cudaMemcpyDeviceToHost));
  if (persistent && opts->timeline_iters > 0 && opts->num_iterations > 0)
    SB_CUDA(cudaMemcpy(opts->h_timeline, ctx->d_timeline, sizeof(uint64_t) * SB_TIMELINE_WORDS * (size_t) opts->timeline_iters,
                       cudaMemcpyDeviceToHost));
  return SB_OK;
}

} // namespace sb

using namespace sb;

extern "C" {

int sb_cg_solve(sb_ctx* ctx, const sb_op* op, double* x, const double* b, const sb_solver_opts* opts,
                sb_solver_report* report, double* h_hist, int64_t hist_cap, double* h_trace, int64_t trace_cap) {
  return run_solver(ctx, op, Kind::Cg, x, b, opts, report, h_hist, hist_cap, h_trace, trace_cap);
}

int sb_bicgstab_solve(sb_ctx* ctx, const sb_op* op, double* x, const double* b, const sb_solver_opts* opts,
                      sb_solver_report* report, double* h_hist, int64_t hist_cap, double* h_trace,
                      int64_t trace_cap) {
  return run_solver(ctx, op, Kind::BiCgStab, x, b, opts, report, h_hist, hist_cap, h_trace, trace_cap);
}

int sb_solve_host(sb_ctx* ctx, const sb_op* op, const char* solver, double* h_x, const double* h_b,
                  const sb_solver_opts* opts, sb_solver_report* report, double* h_hist, int64_t hist_cap) {
  SB_REQUIRE(ctx != nullptr && op != nullptr && solver != nullptr && h_x != nullptr && h_b != nullptr, "null argument");
  const std::string s{solver};
  SB_REQUIRE(s == "cg" || s == "bicgstab", "solver must be \"cg\" or \"bicgstab\"");
  const size_t n = (size_t) op->d.n;
  // x and b staging vectors live after the solver workspaces
  SB_TRY(ensure_work(ctx, n, 7));
  double* d_x = ctx->work[5];
  double* d_b = ctx->work[6];
  SB_CUDA(cudaMemcpyAsync(d_x, h_x, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  SB_CUDA(cudaMemcpyAsync(d_b, h_b, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  SB_TRY(run_solver(ctx, op, s == "cg" ? Kind::Cg : Kind::BiCgStab, d_x, d_b, opts, report, h_hist, hist_cap,
                    nullptr, 0));
  SB_CUDA(cudaMemcpyAsync(h_x, d_x, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  SB_CUDA(cudaStreamSynchronize(ctx->stream));
  return SB_OK;
}

} // extern "C"
