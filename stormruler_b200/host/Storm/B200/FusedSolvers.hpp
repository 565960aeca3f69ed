// Storm/B200/FusedSolvers.hpp -- the fast path behind the reference's Solver interface.
//
// `Storm::B200::CgSolver` and `Storm::B200::BiCgStabSolver` are `Solver<DeviceVector>` objects
// (Storm/Solvers/Solver.hpp:43-57) with the public knobs and progress fields of IterativeSolver
// (:66-76), so application code that holds a `Solver<Vector>&` or sets `num_iterations`,
// `absolute_error_tolerance`, ... keeps working. Instead of init()/iterate() running one kernel
// and one host synchronisation per statement, solve() hands the whole solve to sb_cg_solve /
// sb_bicgstab_solve: same statements, same per-element operation order, same stopping rule as
// SolverCg.hpp:54-126 / SolverBiCgStab.hpp:59-165 (no preconditioner), with scalars resident on the
// device and every reduction fused into the kernel that produces its operand.
//
// The operator must be a Storm::FvmOperator (the fused kernels need its uploaded rows); any other
// Operator<DeviceVector> is rejected with an exception -- use the generic solver templates
// (CgSolver<DeviceVector> etc.) for those. A preconditioner is rejected for the same reason.
#pragma once

#include <Storm/B200/DeviceVector.hpp>

#include <Storm/Solvers/Solver.hpp>

#include <string>
#include <vector>

namespace Storm::B200 {

class FusedSolver : public Solver<DeviceVector> {
public:

  // IterativeSolver's public surface (Solver.hpp:66-76)
  size_t iteration{0};
  size_t num_iterations{2000};
  real_t absolute_error{0.0};
  real_t relative_error{0.0};
  real_t absolute_error_tolerance{1.0e-6};
  real_t relative_error_tolerance{1.0e-6};
  PreconditionerSide pre_side{PreconditionerSide::Right};
  std::unique_ptr<Preconditioner<DeviceVector>> pre_op{nullptr};
  std::string name;

  // additions
  int schedule{SB_SCHEDULE_AUTO};   ///< SB_SCHEDULE_*: persistent whole-solve kernel (default where available) or stepwise
  bool use_graph{true};             ///< stepwise schedule: replay one captured CUDA graph per iteration
  int check_every{0};               ///< host polls the device stop flag every this many iterations (0 = 32)
  bool record_history{false};       ///< keep residual_history / reduction_trace
  std::vector<double> residual_history; ///< [0] initial, [k] after iteration k
  std::vector<double> reduction_trace;  ///< every dot/norm value, reference call order
  sb_solver_report report{};

  bool solve(DeviceVector& x, const DeviceVector& b, const Operator<DeviceVector>& any_op) final {
    const auto* op = dynamic_cast<const FvmOperator*>(&any_op);
    if (op == nullptr) {
      throw std::runtime_error("stormb200: fused solvers need a Storm::FvmOperator; "
                               "use the generic solver templates for other operators");
    }
    if (pre_op != nullptr) {
      throw std::runtime_error("stormb200: fused solvers do not take a preconditioner; "
                               "use the generic solver templates");
    }
    if (x.size() != b.size()) { // the C ABI checks both against the operator's rows and their padded capacity
      throw std::runtime_error("stormb200: fused solve: x and b must have the same number of rows");
    }
    flush(); // queued statements (statement grouping) may produce x or b
    sb_solver_opts opts{};
    opts.schedule = schedule;
    opts.num_iterations = (int64_t) num_iterations;
    opts.abs_tol = absolute_error_tolerance, opts.rel_tol = relative_error_tolerance;
    opts.check_every = check_every, opts.use_graph = use_graph ? 1 : 0, opts.profile = 0;
    const int64_t hist_cap = record_history ? (int64_t) num_iterations + 2 : 0;
    const int64_t trace_cap = record_history ? trace_per_iteration() * (int64_t) num_iterations + 8 : 0;
    residual_history.assign((size_t) hist_cap, 0.0);
    reduction_trace.assign((size_t) trace_cap, 0.0);
    check(run(op->context(), op->handle(), x.data(), b.data(), &opts, &report,
              record_history ? residual_history.data() : nullptr, hist_cap,
              record_history ? reduction_trace.data() : nullptr, trace_cap),
          "fused solve");
    residual_history.resize((size_t) report.n_hist);
    reduction_trace.resize((size_t) report.n_trace);
    iteration = (size_t) report.iterations;
    absolute_error = report.abs_err, relative_error = report.rel_err;
    return report.converged != 0;
  }

protected:

  virtual int run(sb_ctx*, const sb_op*, double*, const double*, const sb_solver_opts*, sb_solver_report*,
                  double*, int64_t, double*, int64_t) = 0;
  virtual int64_t trace_per_iteration() const = 0;
};

/// Drop-in for CgSolver<DeviceVector> (SolverCg.hpp:47-128).
class CgSolver final : public FusedSolver {
  int run(sb_ctx* c, const sb_op* o, double* x, const double* b, const sb_solver_opts* so, sb_solver_report* r,
          double* h, int64_t hc, double* t, int64_t tc) override {
    return sb_cg_solve(c, o, x, b, so, r, h, hc, t, tc);
  }
  int64_t trace_per_iteration() const override { return 2; }
};

/// Drop-in for BiCgStabSolver<DeviceVector> (SolverBiCgStab.hpp:53-167).
class BiCgStabSolver final : public FusedSolver {
  int run(sb_ctx* c, const sb_op* o, double* x, const double* b, const sb_solver_opts* so, sb_solver_report* r,
          double* h, int64_t hc, double* t, int64_t tc) override {
    return sb_bicgstab_solve(c, o, x, b, so, r, h, hc, t, tc);
  }
  int64_t trace_per_iteration() const override { return 5; }
};

/// Drop-in for GmresSolver<DeviceVector> / FgmresSolver<DeviceVector> without a preconditioner
/// (SolverGmres.hpp:42-310): device-resident Arnoldi process, host-side Givens bookkeeping (sb_gmres_solve).
/// `num_inner_iterations` is InnerOuterIterativeSolver's restart length (Solver.hpp:159).
class GmresSolver final : public FusedSolver {
public:

  size_t inner_iteration{0};
  size_t num_inner_iterations{50};
  int lookahead{0};

private:

  int run(sb_ctx* c, const sb_op* o, double* x, const double* b, const sb_solver_opts* so, sb_solver_report* r,
          double* h, int64_t hc, double* t, int64_t tc) override {
    sb_gmres_opts go{};
    go.num_iterations = so->num_iterations, go.abs_tol = so->abs_tol, go.rel_tol = so->rel_tol;
    go.num_inner_iterations = (int32_t) num_inner_iterations, go.lookahead = lookahead;
    const int rc = sb_gmres_solve(c, o, x, b, &go, r, h, hc, t, tc);
    if (rc == 0 && r->iterations > 0) inner_iteration = (size_t) ((r->iterations - 1) % (int64_t) num_inner_iterations);
    return rc;
  }
  int64_t trace_per_iteration() const override { return (int64_t) num_inner_iterations + 3; }
};

} // namespace Storm::B200
