// sb_solvers.cu -- fused CG and BiCGStab (the two solvers BASELINE.json's targets are quoted on).
//
// Reference: CgSolver (source/Storm/Solvers/SolverCg.hpp:54-126) and BiCgStabSolver
// (SolverBiCgStab.hpp:59-165), driven as IterativeSolver::solve (Solver.hpp:116-147), no
// preconditioner. Every statement keeps the reference's per-element operation order, so with the
// same reduction tree the iterates are bit-identical to the CPU restatement (oracle, ORC_RED_TREE).
//
// What is different is the schedule (SURVEY.md a13/a14 "fused minimum"):
//   * scalars (alpha, beta, rho, omega, gamma, residual, iteration, stop flag) live in a device
//     struct; the last CTA of each reducing kernel updates them, so an iteration needs no host sync;
//   * every dot/norm is fused into the kernel that produces its operand (apply or update);
//   * BiCGStab's `x += alpha*p` is deferred into the final update kernel, which evaluates
//     x = (x + alpha*p) + omega*r with the same two roundings per term as the reference;
//   * once the device-side stop flag is set, the remaining queued kernels return immediately, so the
//     solution is exactly the reference's iterate at the stopping iteration.
// CG: 3 kernels / iteration, 9 vector passes + 1 apply. BiCGStab: 5 kernels, 15 passes + 2 applies.
#include "sb_op.cuh"

#include <cmath>
#include <string>

namespace sb {} // namespace sb

struct SolverState {
  double gamma, alpha, beta, rho, omega;
  double initial_err, abs_err, rel_err;
  double abs_tol, rel_tol;
  long long iteration, max_iter;
  long long n_hist, n_trace, hist_cap, trace_cap;
  int done, converged;
};

namespace sb {

__device__ __forceinline__ double safe_divide(double x, double y) {
  // Crow/MathUtils.hpp:49-52
  return (y == 0.0) ? 0.0 : __ddiv_rn(x, y);
}

struct Recorder {
  SolverState* st;
  double* hist;
  double* trace;
  __device__ void push_trace(double v) const {
    if (trace != nullptr && st->n_trace < st->trace_cap) trace[st->n_trace] = v;
    st->n_trace++;
  }
  __device__ void push_hist(double v) const {
    if (hist != nullptr && st->n_hist < st->hist_cap) hist[st->n_hist] = v;
    st->n_hist++;
  }
  // Solver.hpp:124-128: early exit when the initial residual is already below abs_tol.
  __device__ void init_error(double err) const {
    st->initial_err = err, st->abs_err = err, st->rel_err = 0.0;
    st->iteration = 0;
    push_hist(err);
    if (st->abs_tol > 0.0 && err < st->abs_tol) st->converged = 1, st->done = 1;
    if (st->max_iter <= 0) st->done = 1;
  }
  // Solver.hpp:132-140: one pass of the iteration loop after iterate() returned `err`.
  __device__ void iteration_error(double err) const {
    st->abs_err = err;
    st->rel_err = __ddiv_rn(err, st->initial_err); // no zero guard
    push_hist(err);
    bool conv = (st->abs_tol > 0.0) && (err < st->abs_tol);
    conv |= (st->rel_tol > 0.0) && (st->rel_err < st->rel_tol);
    st->iteration++;
    if (conv) st->converged = 1;
    if (conv || st->iteration >= st->max_iter) st->done = 1;
  }
};

// ---- CG -------------------------------------------------------------------------------------------
struct CgInitFinal { // after r = b - A x fused with <r,r>   (SolverCg.hpp:73,80,83)
  Recorder rec;
  __device__ void operator()(const double* s) const {
    rec.st->gamma = s[0];
    rec.push_trace(s[0]);
    rec.init_error(sqrt(s[0]));
  }
};
struct CgAlphaFinal { // after z = A p fused with <p,z>       (:95-96)
  Recorder rec;
  __device__ void operator()(const double* s) const {
    rec.push_trace(s[0]);
    rec.st->alpha = safe_divide(rec.st->gamma, s[0]);
  }
};
struct CgBetaFinal { // after the x/r update fused with <r,r>  (:109,114,121,124)
  Recorder rec;
  __device__ void operator()(const double* s) const {
    const double gamma_bar = rec.st->gamma;
    rec.st->gamma = s[0];
    rec.push_trace(s[0]);
    rec.st->beta = safe_divide(s[0], gamma_bar);
    rec.iteration_error(sqrt(s[0]));
  }
};

struct CopyBody { // p <- r
  double* dst;
  const double* src;
  struct Regs {
    double2 v;
  };
  __device__ __forceinline__ void load(int64_t e0, Regs& r) const { r.v = ld2(src, e0); }
  __device__ __forceinline__ void run(int64_t e0, int64_t, Regs& r, double (&)[1]) const { st2(dst, e0, r.v); }
};

struct CgUpdateBody { // x += alpha*p ; r -= alpha*z ; acc += r.r     (:97-98,114)
  const SolverState* st;
  double *x, *r;
  const double *p, *z;
  struct Regs {
    double2 x, r, p, z;
  };
  __device__ __forceinline__ void load(int64_t e0, Regs& g) const {
    g.x = ld2(x, e0), g.p = ld2(p, e0), g.r = ld2(r, e0), g.z = ld2(z, e0);
  }
  __device__ __forceinline__ void run(int64_t e0, int64_t n, Regs& g, double (&acc)[1]) const {
    const double alpha = st->alpha;
    double2 xn, rn;
    xn.x = __dadd_rn(g.x.x, __dmul_rn(alpha, g.p.x));
    xn.y = __dadd_rn(g.x.y, __dmul_rn(alpha, g.p.y));
    rn.x = __dsub_rn(g.r.x, __dmul_rn(alpha, g.z.x));
    rn.y = __dsub_rn(g.r.y, __dmul_rn(alpha, g.z.y));
    st2(x, e0, xn);
    st2(r, e0, rn);
    acc_pair(acc[0], e0, n, __dmul_rn(rn.x, rn.x), __dmul_rn(rn.y, rn.y));
  }
};

struct CgDirectionBody { // p <- r + beta*p     (:122)
  const SolverState* st;
  double* p;
  const double* r;
  struct Regs {
    double2 p, r;
  };
  __device__ __forceinline__ void load(int64_t e0, Regs& g) const { g.p = ld2(p, e0), g.r = ld2(r, e0); }
  __device__ __forceinline__ void run(int64_t e0, int64_t, Regs& g, double (&)[1]) const {
    const double beta = st->beta;
    double2 pn;
    pn.x = __dadd_rn(g.r.x, __dmul_rn(beta, g.p.x));
    pn.y = __dadd_rn(g.r.y, __dmul_rn(beta, g.p.y));
    st2(p, e0, pn);
  }
};

// ---- BiCGStab -------------------------------------------------------------------------------------
struct BiInitFinal { // r = b - A x, r~ = r, rho = <r~,r>      (SolverBiCgStab.hpp:83,88-91)
  Recorder rec;
  __device__ void operator()(const double* s) const {
    rec.st->rho = s[0];
    rec.push_trace(s[0]);
    rec.init_error(sqrt(s[0]));
  }
};
struct BiAlphaFinal { // after v = A p fused with <r~,v>        (:137,139)
  Recorder rec;
  __device__ void operator()(const double* s) const {
    rec.push_trace(s[0]);
    rec.st->alpha = safe_divide(rec.st->rho, s[0]);
  }
};
struct BiOmegaFinal { // after t = A r fused with <t,t>, <t,r>  (:158-160)
  Recorder rec;
  __device__ void operator()(const double* s) const {
    // g++ evaluates safe_divide's arguments right to left: <t,t> is traced before <t,r>.
    rec.push_trace(s[0]);
    rec.push_trace(s[1]);
    rec.st->omega = safe_divide(s[1], s[0]);
  }
};
struct BiEndFinal { // after the final update fused with <r,r> and <r~,r>   (:164 and next :115-117)
  Recorder rec;
  __device__ void operator()(const double* s) const {
    const double nrm = sqrt(s[0]);
    rec.push_trace(nrm);
    rec.iteration_error(nrm);
    if (!rec.st->done) {
      // head of the next iteration: rho_bar <- rho, rho <- <r~,r>, beta <- (alpha*rho)/(omega*rho_bar)
      const double rho_bar = rec.st->rho;
      rec.st->rho = s[1];
      rec.push_trace(s[1]);
      rec.st->beta = safe_divide(__dmul_rn(rec.st->alpha, s[1]), __dmul_rn(rec.st->omega, rho_bar));
    }
  }
};

struct BiInitBody { // r~ <- r after the fused residual (r already stored by the apply kernel)
  double* rt;
  const double* r;
  struct Regs {
    double2 v;
  };
  __device__ __forceinline__ void load(int64_t e0, Regs& g) const { g.v = ld2(r, e0); }
  __device__ __forceinline__ void run(int64_t e0, int64_t, Regs& g, double (&)[1]) const { st2(rt, e0, g.v); }
};

struct BiDirectionBody { // iteration 0: p <- r ; else p <- r + beta*(p - omega*v)   (:111-119)
  const SolverState* st;
  double* p;
  const double *r, *v;
  struct Regs {
    double2 p, r, v;
  };
  __device__ __forceinline__ void load(int64_t e0, Regs& g) const {
    g.r = ld2(r, e0);
    if (st->iteration != 0) g.p = ld2(p, e0), g.v = ld2(v, e0);
  }
  __device__ __forceinline__ void run(int64_t e0, int64_t, Regs& g, double (&)[1]) const {
    double2 pn = g.r;
    if (st->iteration != 0) {
      const double beta = st->beta, omega = st->omega;
      pn.x = __dadd_rn(g.r.x, __dmul_rn(beta, __dsub_rn(g.p.x, __dmul_rn(omega, g.v.x))));
      pn.y = __dadd_rn(g.r.y, __dmul_rn(beta, __dsub_rn(g.p.y, __dmul_rn(omega, g.v.y))));
    }
    st2(p, e0, pn);
  }
};

struct BiHalfBody { // r -= alpha*v     (:141); x += alpha*p is deferred to BiEndBody
  const SolverState* st;
  double* r;
  const double* v;
  struct Regs {
    double2 r, v;
  };
  __device__ __forceinline__ void load(int64_t e0, Regs& g) const { g.r = ld2(r, e0), g.v = ld2(v, e0); }
  __device__ __forceinline__ void run(int64_t e0, int64_t, Regs& g, double (&)[1]) const {
    const double alpha = st->alpha;
    double2 rn;
    rn.x = __dsub_rn(g.r.x, __dmul_rn(alpha, g.v.x));
    rn.y = __dsub_rn(g.r.y, __dmul_rn(alpha, g.v.y));
    st2(r, e0, rn);
  }
};

struct BiEndBody { // x = (x + alpha*p) + omega*r ; r -= omega*t ; acc0 += r.r ; acc1 += r~.r   (:140,161-164)
  const SolverState* st;
  double *x, *r;
  const double *p, *t, *rt;
  struct Regs {
    double2 x, r, p, t, rt;
  };
  __device__ __forceinline__ void load(int64_t e0, Regs& g) const {
    g.x = ld2(x, e0), g.p = ld2(p, e0), g.r = ld2(r, e0), g.t = ld2(t, e0), g.rt = ld2(rt, e0);
  }
  __device__ __forceinline__ void run(int64_t e0, int64_t n, Regs& g, double (&acc)[2]) const {
    const double alpha = st->alpha, omega = st->omega;
    double2 xn, rn;
    xn.x = __dadd_rn(__dadd_rn(g.x.x, __dmul_rn(alpha, g.p.x)), __dmul_rn(omega, g.r.x));
    xn.y = __dadd_rn(__dadd_rn(g.x.y, __dmul_rn(alpha, g.p.y)), __dmul_rn(omega, g.r.y));
    rn.x = __dsub_rn(g.r.x, __dmul_rn(omega, g.t.x));
    rn.y = __dsub_rn(g.r.y, __dmul_rn(omega, g.t.y));
    st2(x, e0, xn);
    st2(r, e0, rn);
    acc_pair(acc[0], e0, n, __dmul_rn(rn.x, rn.x), __dmul_rn(rn.y, rn.y));
    acc_pair(acc[1], e0, n, __dmul_rn(g.rt.x, rn.x), __dmul_rn(g.rt.y, rn.y));
  }
};

// ---- host side ------------------------------------------------------------------------------------
template<int ND, class Body, class Final>
int launch_ew(sb_ctx* ctx, int64_t n, const Body& body, const Final& fin, const int* done) {
  RedPtrs red{ctx->red.partials, ctx->red.cap_tiles};
  SB_CUDA(launch_kernel(ctx, ew_kernel<ND, Body>, (unsigned) num_tiles(n), kThreads, 0, n, body, red, done));
  ctx->launches++;
  if constexpr (ND > 0) return launch_final<ND>(ctx, n, fin, done);
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
  if (ctx->d_state == nullptr) SB_CUDA(cudaMalloc(&ctx->d_state, sizeof(SolverState)));
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

enum class Kind { Cg, BiCgStab };

struct Solve {
  sb_ctx* ctx;
  const sb_op* op;
  double* x;
  const double* b;
  int64_t n;
  Recorder rec;
  const int* done;
  double *p, *r, *z, *rt, *t, *v;
  std::vector<cudaEvent_t>* prof = nullptr; // profile=1: one event before every kernel + one per iteration end

  int mark() {
    if (prof != nullptr) {
      cudaEvent_t e;
      SB_CUDA(cudaEventCreate(&e));
      SB_CUDA(cudaEventRecord(e, ctx->stream));
      prof->push_back(e);
    }
    return SB_OK;
  }

  int init(Kind kind) {
    // r <- b - A x fused with <r,r>  (Operator::Residual, Operator.hpp:95-99)
    EpiResidual epi{b};
    if (kind == Kind::Cg) {
      SB_TRY((launch_apply<1, true>(ctx, op, x, r, epi, CgInitFinal{rec}, nullptr)));
      SB_TRY((launch_ew<0>(ctx, n, CopyBody{p, r}, NoFinal{}, nullptr))); // p <- r
    } else {
      SB_TRY((launch_apply<1, true>(ctx, op, x, r, epi, BiInitFinal{rec}, nullptr)));
      SB_TRY((launch_ew<0>(ctx, n, BiInitBody{rt, r}, NoFinal{}, nullptr))); // r~ <- r
    }
    return SB_OK;
  }

  int iterate(Kind kind) {
    const SolverState* st = rec.st;
    if (kind == Kind::Cg) {
      SB_TRY(mark());
      SB_TRY((launch_apply<1, false>(ctx, op, p, z, EpiXY{}, CgAlphaFinal{rec}, done)));
      SB_TRY(mark());
      SB_TRY((launch_ew<1>(ctx, n, CgUpdateBody{st, x, r, p, z}, CgBetaFinal{rec}, done)));
      SB_TRY(mark());
      SB_TRY((launch_ew<0>(ctx, n, CgDirectionBody{st, p, r}, NoFinal{}, done)));
    } else {
      SB_TRY(mark());
      SB_TRY((launch_ew<0>(ctx, n, BiDirectionBody{st, p, r, v}, NoFinal{}, done)));
      SB_TRY(mark());
      SB_TRY((launch_apply<1, false>(ctx, op, p, v, EpiUY{rt}, BiAlphaFinal{rec}, done)));
      SB_TRY(mark());
      SB_TRY((launch_ew<0>(ctx, n, BiHalfBody{st, r, v}, NoFinal{}, done)));
      SB_TRY(mark());
      SB_TRY((launch_apply<2, false>(ctx, op, r, t, EpiYYandYX{}, BiOmegaFinal{rec}, done)));
      SB_TRY(mark());
      SB_TRY((launch_ew<2>(ctx, n, BiEndBody{st, x, r, p, t, rt}, BiEndFinal{rec}, done)));
    }
    SB_TRY(mark());
    return SB_OK;
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
  const int64_t n = op->d.n;
  SB_CUDA(cudaSetDevice(ctx->device));
  SB_TRY(ensure_work(ctx, (size_t) n, kind == Kind::Cg ? 3 : 5));
  SB_TRY(ensure_records(ctx, h_hist ? hist_cap : 0, h_trace ? trace_cap : 0));
  SB_TRY(ensure_red_scratch(ctx, n));

  SolverState h{};
  h.abs_tol = opts->abs_tol, h.rel_tol = opts->rel_tol;
  h.max_iter = opts->num_iterations;
  h.hist_cap = h_hist ? hist_cap : 0, h.trace_cap = h_trace ? trace_cap : 0;
  SolverState* pinned_state = reinterpret_cast<SolverState*>(ctx->h_pinned);
  *pinned_state = h;
  SB_CUDA(cudaMemcpyAsync(ctx->d_state, pinned_state, sizeof(SolverState), cudaMemcpyHostToDevice, ctx->stream));

  Solve S;
  S.ctx = ctx, S.op = op, S.x = x, S.b = b, S.n = n;
  S.rec = Recorder{ctx->d_state, h_hist ? ctx->d_hist : nullptr, h_trace ? ctx->d_trace : nullptr};
  S.done = &ctx->d_state->done;
  S.p = ctx->work[0], S.r = ctx->work[1];
  if (kind == Kind::Cg) {
    S.z = ctx->work[2];
  } else {
    S.rt = ctx->work[2], S.t = ctx->work[3], S.v = ctx->work[4];
  }
  const int64_t launches0 = ctx->launches;
  SB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  SB_TRY(S.init(kind));
  cudaEvent_t ev_mid;
  SB_CUDA(cudaEventCreate(&ev_mid));
  SB_CUDA(cudaEventRecord(ev_mid, ctx->stream));
  std::vector<cudaEvent_t> prof_events;
  const bool profile = opts->profile != 0 && !opts->use_graph;

  // One captured graph per iteration: the kernel arguments never change (scalars are read from the
  // device state), so the same graph is replayed; this removes the per-kernel launch cost that
  // matters once a rank holds ~1 M cells.
  cudaGraphExec_t graph_exec = nullptr;
  if (opts->use_graph && opts->num_iterations > 0) {
    cudaGraph_t graph = nullptr;
    SB_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    const int64_t before = ctx->launches;
    const int rc = S.iterate(kind);
    cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
    ctx->launches = before;
    if (rc != SB_OK) return rc;
    SB_CUDA(ce);
    SB_CUDA(cudaGraphInstantiate(&graph_exec, graph, 0));
    SB_CUDA(cudaGraphDestroy(graph));
  }
  const int per_iter = (kind == Kind::Cg) ? 3 : 5;       // profiled kernel slots (each includes its final stage)
  const int launches_per_iter = (kind == Kind::Cg) ? 5 : 8; // + one-CTA final-reduce launches

  // Convergence polling: a flag copy is queued every `check` iterations and examined one batch later,
  // so the host never drains the stream while it still has work to enqueue.
  const int check = opts->check_every > 0 ? opts->check_every : 32;
  int* h_flags = reinterpret_cast<int*>(ctx->h_pinned + 256);
  h_flags[0] = h_flags[1] = 0;
  cudaEvent_t evs[2];
  SB_CUDA(cudaEventCreateWithFlags(&evs[0], cudaEventDisableTiming));
  SB_CUDA(cudaEventCreateWithFlags(&evs[1], cudaEventDisableTiming));
  int64_t it = 0;
  int slot = 0;
  bool pending[2] = {false, false};
  bool stop = false;
  while (it < opts->num_iterations && !stop) {
    const int64_t batch_end = std::min<int64_t>(it + check, opts->num_iterations);
    for (; it < batch_end; ++it) {
      if (graph_exec != nullptr) {
        SB_CUDA(cudaGraphLaunch(graph_exec, ctx->stream));
        ctx->launches += launches_per_iter;
      } else {
        if (profile) S.prof = &prof_events;
        SB_TRY(S.iterate(kind));
      }
    }
    SB_CUDA(cudaMemcpyAsync(&h_flags[slot], &ctx->d_state->done, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
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
  SB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  SB_CUDA(cudaMemcpyAsync(pinned_state, ctx->d_state, sizeof(SolverState), cudaMemcpyDeviceToHost, ctx->stream));
  SB_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaEventDestroy(evs[0]);
  cudaEventDestroy(evs[1]);
  if (graph_exec != nullptr) cudaGraphExecDestroy(graph_exec);
  const SolverState out = *pinned_state;
  float ms = 0.f, ms_iter = 0.f;
  SB_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  SB_CUDA(cudaEventElapsedTime(&ms_iter, ev_mid, ctx->ev1));
  cudaEventDestroy(ev_mid);
  report->iter_ms = ms_iter;
  report->n_kernel_slots = per_iter;
  for (int k = 0; k < SB_MAX_KERNEL_SLOTS; ++k) report->kernel_ms[k] = 0.0;
  if (profile) {
    // events come in groups of per_iter + 1 per iteration
    const size_t group = (size_t) per_iter + 1;
    for (size_t g = 0; g + group <= prof_events.size(); g += group)
      for (int k = 0; k < per_iter; ++k) {
        float dt = 0.f;
        cudaEventElapsedTime(&dt, prof_events[g + k], prof_events[g + k + 1]);
        report->kernel_ms[k] += dt;
      }
    for (cudaEvent_t e : prof_events) cudaEventDestroy(e);
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
