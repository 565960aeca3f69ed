// sb_gmres.cu -- fused restarted GMRES(m) (BASELINE.json configs[2]).
//
// Reference: BaseGmresSolver (source/Storm/Solvers/SolverGmres.hpp:42-262) driven by
// InnerOuterIterativeSolver (Solver.hpp:154-259) and IterativeSolver::solve (Solver.hpp:116-147), no
// preconditioner (FGMRES without a preconditioner is the same algorithm). Same
// statements, same per-element operation order, same reduction tree: bit-identical to the reference
// headers run with the oracle's tree reductions.
//
// What is different is the schedule. The reference does, per inner step k, one apply and k+1 pairs of
// (dot, axpy) with a host round trip after every dot: (5k+8) vector passes and k+2 synchronisations.
// Here:
//   * the Arnoldi process never waits for the host. Each modified-Gram-Schmidt step is ONE kernel:
//     q_{k+1} -= H(i,k) q_i fused with the NEXT projection <q_{k+1}, q_{i+1}> (4 passes per basis vector
//     instead of 5; the first projection rides on the operator apply, the last kernel produces the norm);
//     H(i,k) is read from device memory, where the one-CTA final stage of the previous kernel put it;
//   * the Givens rotations, the residual estimate |beta_{k+1}| and the stopping rule run on the HOST, a
//     few steps behind the device, from the H columns that arrive by asynchronous copies. They are the
//     reference's own scalar statements (SolverGmres.hpp:176-191, 207-212) with the same libm hypot, so
//     every scalar is bit-identical by construction. The device needs nothing back from the host inside a
//     restart cycle; Arnoldi steps queued beyond the stopping iteration only write basis vectors that the
//     solution update does not read;
//   * the solution update x += sum beta_i q_i is one kernel that keeps the reference's left-to-right
//     accumulation per element.
// Per inner step k: B_apply + (4k+6) V, the contract figure of SURVEY.md 8d.
#include "sb_op.cuh"

#include <cmath>
#include <tuple>

namespace sb {

// dst[0] = sum (or its square root): where the Arnoldi kernels pick their coefficients up.
struct StoreAtFinal {
  double* dst;
  int take_sqrt;
  __device__ void operator()(const double* s) const { dst[0] = take_sqrt ? sqrt(s[0]) : s[0]; }
};

// y -= h*a ; acc += y_new . b      (b == nullptr: b is y itself -> acc += y_new^2)
// SolverGmres.hpp:157-161: H(i,k) = <q_{k+1}, q_i>; q_{k+1} -= H(i,k) q_i, and the next <q_{k+1}, q_{i+1}>.
struct GmresOrthoBody {
  const double* h;
  double* y;
  const double* a;
  const double* b;
  struct Regs {
    double2 y, a, b;
  };
  __device__ __forceinline__ void load(int64_t e0, Regs& r) const {
    r.y = ld2(y, e0), r.a = ld2(a, e0);
    if (b != nullptr) r.b = ld2(b, e0);
  }
  __device__ __forceinline__ void run(int64_t e0, int64_t n, Regs& r, double (&acc)[1]) const {
    const double hv = *h;
    double2 yn;
    yn.x = __dsub_rn(r.y.x, __dmul_rn(hv, r.a.x));
    yn.y = __dsub_rn(r.y.y, __dmul_rn(hv, r.a.y));
    st2(y, e0, yn);
    const double2 o = (b != nullptr) ? r.b : yn;
    acc_pair(acc[0], e0, n, __dmul_rn(yn.x, o.x), __dmul_rn(yn.y, o.y));
  }
};

// y /= *s      (SolverGmres.hpp:116,162: q /= beta, q /= H(k+1,k))
struct GmresScaleBody {
  const double* s;
  double* y;
  struct Regs {
    double2 y;
  };
  __device__ __forceinline__ void load(int64_t e0, Regs& r) const { r.y = ld2(y, e0); }
  __device__ __forceinline__ void run(int64_t e0, int64_t, Regs& r, double (&)[1]) const {
    const double sv = *s;
    st2(y, e0, make_double2(__ddiv_rn(r.y.x, sv), __ddiv_rn(r.y.y, sv)));
  }
};

// x += beta_0 q_0; x += beta_1 q_1; ... in that order per element (SolverGmres.hpp:233-236)
constexpr int kMaxBasis = 129;
struct GmresUpdateBody {
  double* x;
  const double* const* q; // device array of basis vectors
  const double* beta;     // device
  int count;
  struct Regs {
    double2 x;
  };
  __device__ __forceinline__ void load(int64_t e0, Regs& r) const { r.x = ld2(x, e0); }
  __device__ __forceinline__ void run(int64_t e0, int64_t, Regs& r, double (&)[1]) const {
    double2 xv = r.x;
    int i = 0;
    for (; i + 4 <= count; i += 4) { // loads of four basis vectors in flight, additions in order
      double2 v[4];
      double bt[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = ld2(q[i + u], e0), bt[u] = beta[i + u];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        xv.x = __dadd_rn(xv.x, __dmul_rn(bt[u], v[u].x));
        xv.y = __dadd_rn(xv.y, __dmul_rn(bt[u], v[u].y));
      }
    }
    for (; i < count; ++i) {
      const double2 v = ld2(q[i], e0);
      const double bt = beta[i];
      xv.x = __dadd_rn(xv.x, __dmul_rn(bt, v.x));
      xv.y = __dadd_rn(xv.y, __dmul_rn(bt, v.y));
    }
    st2(x, e0, xv);
  }
};

template<int ND, class Body, class Final>
static int gm_launch_ew(sb_ctx* ctx, int64_t n, const Body& body, const Final& fin) {
  RedPtrs red{ctx->red.partials, ctx->red.cap_tiles};
  SB_CUDA(launch_kernel(ctx, ew_kernel<ND, Body>, (unsigned) num_tiles(n), kThreads, 0, n, body, red, (const int*) nullptr));
  ctx->launches++;
  if constexpr (ND > 0) return launch_final<ND>(ctx, n, fin, nullptr);
  return SB_OK;
}

// Host mirror of the reference's scalar state (SolverGmres.hpp:45-46): beta, cs, sn, H.
struct GmresScalars {
  int m;
  std::vector<double> beta, cs, sn, H; // H is (m+1) x m, row-major like Solvers/MatrixDense.hpp
  double& h(int i, int j) { return H[(size_t) i * m + j]; }
  explicit GmresScalars(int m_) : m(m_), beta((size_t) m_ + 1, 0.0), cs((size_t) m_, 0.0), sn((size_t) m_, 0.0), H((size_t) (m_ + 1) * m_, 0.0) {}

  // Crow/MathUtils.hpp:164-179
  static std::tuple<double, double, double> sym_ortho(double a, double b) {
    double cs_, sn_;
    const double rr = std::hypot(a, b);
    if (rr > 0.0) {
      cs_ = a / rr, sn_ = b / rr;
    } else {
      cs_ = 1.0, sn_ = 0.0;
    }
    return {cs_, sn_, rr};
  }

  // SolverGmres.hpp:176-191, after column k of H holds the raw projections and the norm. Returns |beta_{k+1}|.
  double rotate(int k) {
    for (int i = 0; i < k; ++i) {
      const double chi = cs[i] * h(i, k) + sn[i] * h(i + 1, k);
      h(i + 1, k) = -sn[i] * h(i, k) + cs[i] * h(i + 1, k);
      h(i, k) = chi;
    }
    std::tie(cs[k], sn[k], std::ignore) = sym_ortho(h(k, k), h(k + 1, k));
    h(k, k) = cs[k] * h(k, k) + sn[k] * h(k + 1, k);
    h(k + 1, k) = 0.0;
    beta[k + 1] = -sn[k] * beta[k], beta[k] *= cs[k];
    return std::abs(beta[k + 1]);
  }

  // SolverGmres.hpp:207-212
  void back_substitute(int k) {
    for (int i = k; i >= 0; --i) {
      for (int j = i + 1; j <= k; ++j) beta[i] -= h(i, j) * beta[j];
      beta[i] /= h(i, i);
    }
  }
};

} // namespace sb

using namespace sb;

extern "C" {

int sb_gmres_solve(sb_ctx* ctx, const sb_op* op, double* x, const double* b, const sb_gmres_opts* opts,
                   sb_solver_report* report, double* h_hist, int64_t hist_cap, double* h_trace, int64_t trace_cap) {
  SB_REQUIRE(ctx != nullptr && op != nullptr && x != nullptr && b != nullptr && opts != nullptr && report != nullptr,
             "null argument");
  SB_REQUIRE(x != b, "x and b must not alias");
  SB_REQUIRE(opts->num_iterations >= 0 && hist_cap >= 0 && trace_cap >= 0, "negative count");
  const int m = opts->num_inner_iterations > 0 ? opts->num_inner_iterations : 50; // Solver.hpp:159
  SB_REQUIRE(m + 1 <= kMaxBasis, "restart length too large (max 128)");
  const int look = opts->lookahead > 0 ? opts->lookahead : 3;
  const int64_t n = op->d.n;
  SB_CUDA(cudaSetDevice(ctx->device));
  SB_TRY(ensure_red_scratch(ctx, n));

  // ---- workspace: m+1 basis vectors (pool blocks in multi-GPU mode: apply inputs need a halo tail), cached
  if (ctx->basis_n < (size_t) n) {
    SB_CUDA(cudaStreamSynchronize(ctx->stream));
    for (double* v : ctx->basis) SB_TRY(vec_free(ctx, v));
    ctx->basis.clear();
    ctx->basis_n = (size_t) n;
  }
  while ((int) ctx->basis.size() < m + 1) {
    double* v = nullptr;
    SB_TRY(vec_alloc(ctx, ctx->basis_n, &v));
    ctx->basis.push_back(v);
  }
  double** q = ctx->basis.data();
  // device scalars: per step a record [H(0,k) .. H(k,k), H(k+1,k)] of m+2 slots; beta0; the betas of the update;
  // the pointer table of the basis
  const size_t rec = (size_t) m + 2;
  const size_t scal_doubles = rec * m + 1 + (size_t) m + 1;
  if (ctx->d_gmres_scal == nullptr || ctx->gmres_scal_cap < scal_doubles) {
    SB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->d_gmres_scal), cudaFree(ctx->d_gmres_ptrs);
    cudaFreeHost(ctx->h_gmres);
    SB_CUDA(cudaMalloc(&ctx->d_gmres_scal, sizeof(double) * scal_doubles));
    SB_CUDA(cudaMalloc(&ctx->d_gmres_ptrs, sizeof(double*) * kMaxBasis));
    SB_CUDA(cudaMallocHost(&ctx->h_gmres, sizeof(double) * (scal_doubles + kMaxBasis)));
    ctx->gmres_scal_cap = scal_doubles;
  }
  double* d_H = ctx->d_gmres_scal;
  double* d_beta0 = d_H + rec * m;
  double* d_betas = d_beta0 + 1;
  double* h_H = ctx->h_gmres;          // pinned mirror of the records
  double* h_beta0 = h_H + rec * m;
  double* h_betas = h_beta0 + 1;
  SB_CUDA(cudaMemcpyAsync(ctx->d_gmres_ptrs, q, sizeof(double*) * (size_t) (m + 1), cudaMemcpyHostToDevice, ctx->stream));
  SB_CUDA(cudaStreamSynchronize(ctx->stream)); // q is host memory owned by a vector that may grow later

  GmresScalars S(m);
  std::vector<double> hist, trace;
  const int64_t launches0 = ctx->launches;
  SB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));

  // r = b - A x fused with <r,r>; beta0 = sqrt; q0 /= beta0   (outer_init / inner_init, SolverGmres.hpp:72-117)
  auto queue_restart = [&]() -> int {
    SB_TRY((launch_apply<1, true>(ctx, op, x, q[0], EpiResidual{b}, StoreAtFinal{d_beta0, 1}, nullptr)));
    SB_TRY((gm_launch_ew<0>(ctx, n, GmresScaleBody{d_beta0, q[0]}, NoFinal{})));
    SB_CUDA(cudaMemcpyAsync(h_beta0, d_beta0, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    return SB_OK;
  };
  // one Arnoldi step (SolverGmres.hpp:149-162), H column record -> pinned host memory
  auto queue_step = [&](int k) -> int {
    double* hk = d_H + rec * k;
    SB_TRY((launch_apply<1, false>(ctx, op, q[k], q[k + 1], EpiUY{q[0]}, StoreAtFinal{hk, 0}, nullptr)));
    for (int i = 0; i < k; ++i)
      SB_TRY((gm_launch_ew<1>(ctx, n, GmresOrthoBody{hk + i, q[k + 1], q[i], q[i + 1]}, StoreAtFinal{hk + i + 1, 0})));
    SB_TRY((gm_launch_ew<1>(ctx, n, GmresOrthoBody{hk + k, q[k + 1], q[k], nullptr}, StoreAtFinal{hk + k + 1, 1})));
    SB_TRY((gm_launch_ew<0>(ctx, n, GmresScaleBody{hk + k + 1, q[k + 1]}, NoFinal{})));
    SB_CUDA(cudaMemcpyAsync(h_H + rec * k, hk, sizeof(double) * (size_t) (k + 2), cudaMemcpyDeviceToHost, ctx->stream));
    return SB_OK;
  };
  // inner_finalize (SolverGmres.hpp:194-237) with the stopping index k
  auto queue_update = [&](int k) -> int {
    S.back_substitute(k);
    SB_CUDA(cudaStreamSynchronize(ctx->stream)); // h_betas may still be read by the previous update's copy
    for (int i = 0; i <= k; ++i) h_betas[i] = S.beta[(size_t) i];
    SB_CUDA(cudaMemcpyAsync(d_betas, h_betas, sizeof(double) * (size_t) (k + 1), cudaMemcpyHostToDevice, ctx->stream));
    SB_TRY((gm_launch_ew<0>(ctx, n, GmresUpdateBody{x, (const double* const*) ctx->d_gmres_ptrs, d_betas, k + 1}, NoFinal{})));
    return SB_OK;
  };

  // events of this solve, destroyed on every return path
  struct Events {
    std::vector<cudaEvent_t> all;
    ~Events() {
      for (cudaEvent_t e : all) cudaEventDestroy(e);
    }
    int make(cudaEvent_t* e, unsigned flags) {
      SB_CUDA(cudaEventCreateWithFlags(e, flags));
      all.push_back(*e);
      return SB_OK;
    }
  } events;
  std::vector<cudaEvent_t> ev((size_t) look + 1);
  for (auto& e : ev) SB_TRY(events.make(&e, cudaEventDisableTiming));

  // ---- IterativeSolver::solve, Solver.hpp:116-147
  SB_TRY(queue_restart());
  SB_CUDA(cudaStreamSynchronize(ctx->stream));
  const double initial = *h_beta0;
  trace.push_back(initial);
  hist.push_back(initial);
  double abs_err = initial, rel_err = 0.0;
  bool converged = false;
  int64_t iteration = 0;
  cudaEvent_t ev_mid;
  SB_TRY(events.make(&ev_mid, cudaEventDefault));
  SB_CUDA(cudaEventRecord(ev_mid, ctx->stream));
  if (opts->abs_tol > 0.0 && initial < opts->abs_tol) {
    // Solver.hpp:124-128 calls finalize() here, which back-substitutes through an all-zero H and turns x into
    // NaN. Deliberate deviation: x is already within tolerance and is left untouched.
    converged = true;
  } else {
    int64_t queued = 0, processed = 0;
    bool stop = false;
    int last_inner = 0;
    bool cycle_open = false; // a restart cycle whose update has not been queued yet
    while (!stop) {
      const bool can_queue = queued < opts->num_iterations && queued - processed < look &&
                             !((queued % m) == 0 && queued != processed);
      if (can_queue) {
        const int k = (int) (queued % m);
        if (k == 0) { // inner_init: recompute q0 from the current x (the update of the last cycle is already queued)
          SB_TRY(queue_restart());
          cycle_open = true;
        }
        SB_TRY(queue_step(k));
        SB_CUDA(cudaEventRecord(ev[(size_t) (queued % (look + 1))], ctx->stream));
        ++queued;
        continue;
      }
      if (processed == queued) break; // num_iterations reached (or == 0)
      SB_CUDA(cudaEventSynchronize(ev[(size_t) (processed % (look + 1))]));
      const int k = (int) (processed % m);
      if (k == 0) {
        S.beta[0] = *h_beta0;
        trace.push_back(*h_beta0);
      }
      const double* col = h_H + rec * k;
      for (int i = 0; i <= k + 1; ++i) S.h(i, k) = col[i], trace.push_back(col[i]);
      abs_err = S.rotate(k);
      rel_err = abs_err / initial;
      hist.push_back(abs_err);
      ++processed;
      iteration = processed;
      last_inner = k;
      converged = (opts->abs_tol > 0.0 && abs_err < opts->abs_tol) || (opts->rel_tol > 0.0 && rel_err < opts->rel_tol);
      if (converged || processed >= opts->num_iterations) stop = true;
      if (k == m - 1 || stop) { // inner_finalize: at the end of a cycle, or from finalize() when stopped mid-cycle
        SB_TRY(queue_update(k));
        cycle_open = false;
      }
    }
    (void) last_inner, (void) cycle_open;
  }
  SB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  SB_CUDA(cudaStreamSynchronize(ctx->stream));
  float ms = 0.f, ms_iter = 0.f;
  SB_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  SB_CUDA(cudaEventElapsedTime(&ms_iter, ev_mid, ctx->ev1));
  report->converged = converged ? 1 : 0;
  report->iterations = iteration;
  report->initial_err = initial, report->abs_err = abs_err, report->rel_err = rel_err;
  report->n_hist = std::min<int64_t>((int64_t) hist.size(), h_hist ? hist_cap : 0);
  report->n_trace = std::min<int64_t>((int64_t) trace.size(), h_trace ? trace_cap : 0);
  report->solve_ms = ms, report->iter_ms = ms_iter;
  report->launches = ctx->launches - launches0;
  report->n_kernel_slots = 0;
  report->schedule = SB_SCHEDULE_STEPWISE;
  for (int k = 0; k < SB_MAX_KERNEL_SLOTS; ++k) report->kernel_ms[k] = 0.0, report->wait_ms[k] = 0.0;
  if (h_hist) std::memcpy(h_hist, hist.data(), sizeof(double) * (size_t) report->n_hist);
  if (h_trace) std::memcpy(h_trace, trace.data(), sizeof(double) * (size_t) report->n_trace);
  return SB_OK;
}

} // extern "C"
