// dropin.cpp -- the reference's solver templates instantiated on Storm::DeviceVector.
//
// This translation unit is what a StormRuler application does after switching its Vector type:
// it includes the reference's OWN solver headers from where they lie (-I<reference>/source, nothing
// copied) after Storm/B200/DeviceVector.hpp and instantiates every solver on the device vector:
//   SolverCg.hpp, SolverCgs.hpp, SolverBiCgStab.hpp (BiCGStab, BiCGStab(l)), SolverGmres.hpp
//   (GMRES, FGMRES), SolverTfqmr.hpp (TFQMR, TFQMR1), SolverIdrs.hpp, SolverRichardson.hpp,
//   SolverNewton.hpp (JFNK: the nonlinear outer loop whose inner solve is BiCGStab),
// all driven by IterativeSolver::solve (Solver.hpp:116-147), plus the two fused fast-path solvers
// of Storm/B200/FusedSolvers.hpp. It is exported behind a small C entry point so the parity tests
// (Python, ctypes) can run it on vectors they own; g++ -std=c++23 builds it, linking libstormb200.so.
// The built libstorm_dropin.so travels to the GPU box (the reference tree does not exist there).
#include <Storm/B200/FusedSolvers.hpp>
#include <Storm/B200/GroupedSolvers.hpp>
#include <Storm/B200/ChebyshevPreconditioner.hpp>

#include <Storm/Solvers/SolverBiCgStab.hpp>
#include <Storm/Solvers/SolverCg.hpp>
#include <Storm/Solvers/SolverCgs.hpp>
#include <Storm/Solvers/SolverGmres.hpp>
#include <Storm/Solvers/SolverIdrs.hpp>
#include <Storm/Solvers/SolverRichardson.hpp>
#include <Storm/Solvers/SolverTfqmr.hpp>
// JFNK (SolverNewton.hpp:101-173). The header relies on BiCgStabSolver and
// std::numeric_limits being visible already ("fix the missing include"): included after them, untouched.
#include <limits>
#include <Storm/Solvers/SolverNewton.hpp>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

namespace {

using Storm::DeviceVector;

struct dropin_opts {
  int64_t num_iterations;
  double abs_tol;
  double rel_tol;
  int64_t num_inner_iterations; // <= 0: keep the solver's default (Solver.hpp:159)
  double relaxation_factor;     // Richardson only; <= 0 keeps the default
  int32_t use_graph;            // fused solvers only
  int32_t precond;              // 0: none, 1: Storm::JacobiPreconditioner, 2: the reference's IdentityPreconditioner,
                                // 3: Storm::ChebyshevPreconditioner (generic solvers)
  int32_t pre_side;             // 0: Left, 1: Right (the reference's default), 2: Symmetric
  int32_t cheb_degree;          // precond 3: terms of the polynomial (<= 0: the class default, 4)
  int32_t cheb_power_iterations; // precond 3: power iterations of build() (<= 0: default 10)
  double cheb_eig_ratio;        // precond 3: lambda_max / lambda_min (<= 0: default 30)
};

struct dropin_report {
  int32_t converged;
  int64_t iterations;
  double abs_err;
  double rel_err;
  int64_t n_hist;
  int64_t n_trace;
  int64_t n_apply;
};

// Constants of the playground's Cahn-Hilliard step (Playground.cpp:113) and the solver limits; a negative tolerance or
// num_iterations <= 0 keeps the IterativeSolver default (Solver.hpp:67-72), which is what solve<CgSolver> runs with.
struct dropin_ch_params {
  double tau, Gamma, sigma;
  int64_t num_iterations;
  double abs_tol, rel_tol;
  int32_t uniformed; // 0: solve<CgSolver> as the playground calls it (Playground.cpp:151); 1: through the reference's
                     // solve_non_uniform (Solver.hpp:271-292) -- the operator is affine, and only then does CG converge
};

thread_local std::string g_error;

struct Trace {
  double* data = nullptr;
  int64_t cap = 0, count = 0;
  static void push(void* user, double v) {
    auto* t = static_cast<Trace*>(user);
    if (t->data != nullptr && t->count < t->cap) t->data[t->count] = v;
    ++t->count;
  }
};

// Forwards to the FVM operator and samples the solver's public progress fields on every call
// (residual-history capture, method (2) of SURVEY.md 8c -- same scheme as the CPU oracle harness).
template<class SolverT>
struct SamplingOperator final : Storm::Operator<DeviceVector> {
  const Storm::FvmOperator* inner;
  const SolverT* solver;
  double* hist;
  int64_t hist_cap;
  mutable int64_t n_apply = 0;
  void mul(DeviceVector& y, const DeviceVector& x) const override {
    const int64_t it = (int64_t) solver->iteration;
    if (hist != nullptr && it < hist_cap) hist[it] = solver->absolute_error;
    inner->mul(y, x);
    ++n_apply;
  }
};

template<class SolverT>
int run_generic(sb_ctx* ctx, const sb_op* op, double* d_x, const double* d_b, size_t n, const dropin_opts* o,
                dropin_report* rep, double* hist, int64_t hist_cap, double* trace, int64_t trace_cap) {
  SolverT solver{};
  solver.num_iterations = (size_t) o->num_iterations;
  solver.absolute_error_tolerance = o->abs_tol;
  solver.relative_error_tolerance = o->rel_tol;
  if constexpr (requires { solver.num_inner_iterations; }) {
    if (o->num_inner_iterations > 0) solver.num_inner_iterations = (size_t) o->num_inner_iterations;
  }
  if constexpr (requires { solver.relaxation_factor; }) {
    if (o->relaxation_factor > 0.0) solver.relaxation_factor = o->relaxation_factor;
  }
  if (o->precond == 2) { // the one preconditioner the reference ships (Preconditioner.hpp:84-97), on the device vector
    solver.pre_op = std::make_unique<Storm::IdentityPreconditioner<DeviceVector>>();
    solver.pre_side = o->pre_side == 0 ? Storm::PreconditionerSide::Left
                                       : (o->pre_side == 2 ? Storm::PreconditionerSide::Symmetric : Storm::PreconditionerSide::Right);
  }
  if (o->precond == 3) {
    auto cheb = std::make_unique<Storm::ChebyshevPreconditioner<DeviceVector>>();
    if (o->cheb_degree > 0) cheb->degree = (size_t) o->cheb_degree;
    if (o->cheb_power_iterations > 0) cheb->num_power_iterations = (size_t) o->cheb_power_iterations;
    if (o->cheb_eig_ratio > 0.0) cheb->eig_ratio = o->cheb_eig_ratio;
    solver.pre_op = std::move(cheb);
    solver.pre_side = o->pre_side == 0 ? Storm::PreconditionerSide::Left
                                       : (o->pre_side == 2 ? Storm::PreconditionerSide::Symmetric : Storm::PreconditionerSide::Right);
  }
  if (o->precond == 1) {
    solver.pre_op = std::make_unique<Storm::JacobiPreconditioner>(ctx, op);
    solver.pre_side = o->pre_side == 0 ? Storm::PreconditionerSide::Left
                                       : (o->pre_side == 2 ? Storm::PreconditionerSide::Symmetric : Storm::PreconditionerSide::Right);
  }
  DeviceVector x = DeviceVector::view(ctx, d_x, n);
  const DeviceVector b = DeviceVector::view(ctx, const_cast<double*>(d_b), n);
  Trace tr{trace, trace_cap, 0};
  Storm::B200::g_observer = &Trace::push, Storm::B200::g_observer_user = &tr;
  const Storm::FvmOperator fvm{ctx, op};
  SamplingOperator<SolverT> sop;
  sop.inner = &fvm, sop.solver = &solver, sop.hist = hist, sop.hist_cap = hist_cap;
  bool converged = false;
  try {
    converged = solver.solve(x, b, sop);
  } catch (...) {
    Storm::B200::g_observer = nullptr;
    throw;
  }
  Storm::B200::g_observer = nullptr;
  const int64_t it = (int64_t) solver.iteration;
  if (hist != nullptr && it < hist_cap) hist[it] = solver.absolute_error;
  rep->converged = converged ? 1 : 0;
  rep->iterations = it;
  rep->abs_err = solver.absolute_error, rep->rel_err = solver.relative_error;
  rep->n_hist = it + 1, rep->n_trace = tr.count, rep->n_apply = sop.n_apply;
  return 0;
}

template<class SolverT>
int run_fused(sb_ctx* ctx, const sb_op* op, double* d_x, const double* d_b, size_t n, const dropin_opts* o,
              dropin_report* rep, double* hist, int64_t hist_cap, double* trace, int64_t trace_cap) {
  SolverT solver{};
  solver.num_iterations = (size_t) o->num_iterations;
  solver.absolute_error_tolerance = o->abs_tol;
  solver.relative_error_tolerance = o->rel_tol;
  solver.use_graph = o->use_graph != 0;
  solver.record_history = true;
  if constexpr (requires { solver.num_inner_iterations; }) {
    if (o->num_inner_iterations > 0) solver.num_inner_iterations = (size_t) o->num_inner_iterations;
  }
  DeviceVector x = DeviceVector::view(ctx, d_x, n);
  const DeviceVector b = DeviceVector::view(ctx, const_cast<double*>(d_b), n);
  const Storm::FvmOperator fvm{ctx, op};
  Storm::Solver<DeviceVector>& as_base = solver; // called through the reference's abstract interface
  const bool converged = as_base.solve(x, b, fvm);
  rep->converged = converged ? 1 : 0;
  rep->iterations = (int64_t) solver.iteration;
  rep->abs_err = solver.absolute_error, rep->rel_err = solver.relative_error;
  rep->n_hist = (int64_t) solver.residual_history.size();
  rep->n_trace = (int64_t) solver.reduction_trace.size();
  rep->n_apply = -1;
  if (hist != nullptr)
    std::memcpy(hist, solver.residual_history.data(), sizeof(double) * (size_t) std::min<int64_t>(rep->n_hist, hist_cap));
  if (trace != nullptr)
    std::memcpy(trace, solver.reduction_trace.data(), sizeof(double) * (size_t) std::min<int64_t>(rep->n_trace, trace_cap));
  return 0;
}

} // namespace

extern "C" {

#define DROPIN_API __attribute__((visibility("default")))

DROPIN_API const char* dropin_last_error(void) { return g_error.c_str(); }

// Solver names: cg cgs bicgstab bicgstabl gmres fgmres tfqmr tfqmr1 idrs richardson jfnk (the reference
// templates on DeviceVector), grouped_idrs grouped_bicgstabl (same algorithms, statements in groups), fused_cg
// fused_bicgstab fused_gmres (Storm::B200 fast path).
DROPIN_API int dropin_solve(const char* name, sb_ctx* ctx, const sb_op* op, double* d_x, const double* d_b,
                            size_t n, const dropin_opts* o, dropin_report* rep, double* hist, int64_t hist_cap,
                            double* trace, int64_t trace_cap) {
  const std::string s{name};
  try {
#define DROPIN_CASE(key, T)                                                                              \
  if (s == key)                                                                                          \
    return run_generic<Storm::T<DeviceVector>>(ctx, op, d_x, d_b, n, o, rep, hist, hist_cap, trace, trace_cap)
    DROPIN_CASE("cg", CgSolver);
    DROPIN_CASE("cgs", CgsSolver);
    DROPIN_CASE("bicgstab", BiCgStabSolver);
    DROPIN_CASE("bicgstabl", BiCgStabLSolver);
    DROPIN_CASE("gmres", GmresSolver);
    DROPIN_CASE("fgmres", FgmresSolver);
    DROPIN_CASE("tfqmr", TfqmrSolver);
    DROPIN_CASE("tfqmr1", Tfqmr1Solver);
    DROPIN_CASE("idrs", IdrsSolver);
    DROPIN_CASE("richardson", RichardsonSolver);
    DROPIN_CASE("jfnk", JfnkSolver);
#undef DROPIN_CASE
    // the same algorithms with their statements issued in groups (Storm/B200/GroupedSolvers.hpp)
    if (s == "grouped_idrs")
      return run_generic<Storm::B200::IdrsSolver>(ctx, op, d_x, d_b, n, o, rep, hist, hist_cap, trace, trace_cap);
    if (s == "grouped_bicgstabl")
      return run_generic<Storm::B200::BiCgStabLSolver>(ctx, op, d_x, d_b, n, o, rep, hist, hist_cap, trace, trace_cap);
    if (s == "fused_cg")
      return run_fused<Storm::B200::CgSolver>(ctx, op, d_x, d_b, n, o, rep, hist, hist_cap, trace, trace_cap);
    if (s == "fused_bicgstab")
      return run_fused<Storm::B200::BiCgStabSolver>(ctx, op, d_x, d_b, n, o, rep, hist, hist_cap, trace, trace_cap);
    if (s == "fused_gmres")
      return run_fused<Storm::B200::GmresSolver>(ctx, op, d_x, d_b, n, o, rep, hist, hist_cap, trace, trace_cap);
  } catch (const std::exception& e) {
    g_error = e.what();
    return -2;
  }
  g_error = "unknown solver name";
  return -1;
}

// One Cahn-Hilliard time step of the playground (Playground.cpp:133-175), the reference's own caller of the path,
// after the switch to the device types: CellField -> DeviceVector, stormDivGrad(mesh, u, dt, c) ->
// Storm::B200::div_grad(faces, u, dt, c), map -> Storm::B200::map and `real_t c` -> `auto c` in dF_dc (the function is
// traced, not called per element). Everything else is the playground's statement sequence; the solver is the reference's CgSolver template.
// `faces` is a faithful-form operator over the mesh (its own dt/prefill are not used). d_c: c (in), d_c_hat: the new
// c (out), d_w_hat: workspace holding the chemical potential of the last operator evaluation (out).
DROPIN_API int dropin_cahn_hilliard_step(sb_ctx* ctx, const sb_op* faces, const double* d_c, double* d_c_hat,
                                         double* d_w_hat, size_t n, const dropin_ch_params* p, dropin_report* rep,
                                         double* hist, int64_t hist_cap, double* trace, int64_t trace_cap) {
  try {
    const double tau = p->tau, Gamma = p->Gamma, sigma = p->sigma;
    const DeviceVector c = DeviceVector::view(ctx, const_cast<double*>(d_c), n);
    DeviceVector c_hat = DeviceVector::view(ctx, d_c_hat, n);
    DeviceVector w_hat = DeviceVector::view(ctx, d_w_hat, n);
    const Storm::FvmOperator mesh{ctx, faces};
    Trace tr{trace, trace_cap, 0};
    Storm::B200::g_observer = &Trace::push, Storm::B200::g_observer_user = &tr;

    constexpr auto dF_dc = [](auto c) { return 2.0 * c * (c - 1.0) * (2.0 * c - 1.0); };
    DeviceVector f{ctx, n};
    f <<= Storm::B200::map(dF_dc, c);

    c_hat <<= c;
    Storm::CgSolver<DeviceVector> solver{}; // what solve<CgSolver>(...) constructs (Solver.hpp:261-265)
    if (p->num_iterations > 0) solver.num_iterations = (size_t) p->num_iterations;
    if (p->abs_tol >= 0.0) solver.absolute_error_tolerance = p->abs_tol;
    if (p->rel_tol >= 0.0) solver.relative_error_tolerance = p->rel_tol;
    int64_t n_apply = 0;
    const auto op = Storm::make_operator<DeviceVector>([&](DeviceVector& c_out, const DeviceVector& c_in) {
      const int64_t it = (int64_t) solver.iteration;
      if (hist != nullptr && it < hist_cap) hist[it] = solver.absolute_error;
      w_hat <<= f + sigma * (c_in - c);
      Storm::B200::div_grad(mesh, w_hat, -Gamma, c_in);
      c_out <<= c_in;
      Storm::B200::div_grad(mesh, c_out, -tau, w_hat);
      ++n_apply;
    });
    bool converged = false;
    try {
      converged = p->uniformed != 0 ? Storm::solve_non_uniform(solver, c_hat, c, *op) : solver.solve(c_hat, c, *op);
    } catch (...) {
      Storm::B200::g_observer = nullptr;
      throw;
    }
    Storm::B200::g_observer = nullptr;
    const int64_t it = (int64_t) solver.iteration;
    if (hist != nullptr && it < hist_cap) hist[it] = solver.absolute_error;
    rep->converged = converged ? 1 : 0;
    rep->iterations = it;
    rep->abs_err = solver.absolute_error, rep->rel_err = solver.relative_error;
    rep->n_hist = it + 1, rep->n_trace = tr.count, rep->n_apply = n_apply;
    return 0;
  } catch (const std::exception& e) {
    g_error = e.what();
    return -2;
  }
}

// solve_non_uniform (Solver.hpp:271-292): an operator equation A(x) = b with an AFFINE operator, A(0) != 0 -- here
// A(x) = L x + shift with L the uploaded FVM operator. The reference's function template, instantiated on the device
// vector: z <- A(0), f <- b - z, then the solver runs on the "uniformed" operator y <- A(x) - z.
DROPIN_API int dropin_solve_non_uniform(const char* name, sb_ctx* ctx, const sb_op* op, double* d_x, const double* d_b,
                                        const double* d_shift, size_t n, const dropin_opts* o, dropin_report* rep,
                                        double* trace, int64_t trace_cap) {
  const std::string s{name};
  try {
    DeviceVector x = DeviceVector::view(ctx, d_x, n);
    const DeviceVector b = DeviceVector::view(ctx, const_cast<double*>(d_b), n);
    const DeviceVector shift = DeviceVector::view(ctx, const_cast<double*>(d_shift), n);
    const Storm::FvmOperator fvm{ctx, op};
    int64_t n_apply = 0;
    const auto affine = Storm::make_operator<DeviceVector>([&](DeviceVector& y, const DeviceVector& in) {
      fvm.mul(y, in);
      y += shift;
      ++n_apply;
    });
    Trace tr{trace, trace_cap, 0};
    auto run = [&](auto solver) {
      solver.num_iterations = (size_t) o->num_iterations;
      solver.absolute_error_tolerance = o->abs_tol, solver.relative_error_tolerance = o->rel_tol;
      Storm::B200::g_observer = &Trace::push, Storm::B200::g_observer_user = &tr;
      bool converged = false;
      try {
        converged = Storm::solve_non_uniform(solver, x, b, *affine);
      } catch (...) {
        Storm::B200::g_observer = nullptr;
        throw;
      }
      Storm::B200::g_observer = nullptr;
      rep->converged = converged ? 1 : 0;
      rep->iterations = (int64_t) solver.iteration;
      rep->abs_err = solver.absolute_error, rep->rel_err = solver.relative_error;
      rep->n_hist = 0, rep->n_trace = tr.count, rep->n_apply = n_apply;
      return 0;
    };
    if (s == "cg") return run(Storm::CgSolver<DeviceVector>{});
    if (s == "bicgstab") return run(Storm::BiCgStabSolver<DeviceVector>{});
    if (s == "gmres") return run(Storm::GmresSolver<DeviceVector>{});
    if (s == "idrs") return run(Storm::IdrsSolver<DeviceVector>{});
  } catch (const std::exception& e) {
    g_error = e.what();
    return -2;
  }
  g_error = "unknown solver name";
  return -1;
}

// A random program over a pool of device vectors -- chain-shaped and other statements, reductions, operator applies,
// fills, copies, pointer swaps, re-allocations, host reads -- run through the DeviceVector operators under whatever
// statement-grouping mode is set. The tests run the same seed with grouping off, on, and on with dependency-aware
// scheduling: every recorded value and every final vector must agree bit for bit (a scheduling mistake -- a statement
// moved past a launch it shares a vector with -- shows up as a difference).
DROPIN_API int dropin_random_program(sb_ctx* ctx, const sb_op* op, size_t n, uint64_t seed, int steps, const double* h_init,
                                     int n_vecs, double* h_final, double* h_record, int64_t record_cap, int64_t* n_record,
                                     int with_accumulate, int with_jacobi) {
  try {
    std::mt19937_64 rng{seed};
    auto pick = [&](int m) { return (int) (rng() % (uint64_t) m); };
    auto coef = [&]() { return ((double) (rng() % 2001) - 1000.0) / 1250.0; }; // [-0.8, 0.8]
    std::vector<DeviceVector> v((size_t) n_vecs);
    for (int k = 0; k < n_vecs; ++k) {
      v[(size_t) k] = DeviceVector{ctx, n};
      v[(size_t) k].upload(h_init + (size_t) k * n);
    }
    const Storm::FvmOperator fvm{ctx, op};
    int64_t nr = 0;
    auto record = [&](double x) {
      if (nr < record_cap) h_record[nr] = x;
      ++nr;
    };
    for (int s = 0; s < steps; ++s) {
      const int a = pick(n_vecs);
      int b = pick(n_vecs), d = pick(n_vecs), e = pick(n_vecs);
      const double c1 = coef(), c2 = coef();
      const int kind = pick(20);
      if (std::getenv("DROPIN_TRACE") != nullptr) std::fprintf(stderr, "step %d kind %d a %d b %d d %d e %d\n", s, kind, a, b, d, e);
      switch (kind) {
        case 0: v[a] <<= v[b] + c1 * v[d]; break;
        case 1: v[a] += c1 * v[b]; break;
        case 2: v[a] -= c1 * v[b]; break;
        case 3: v[a] <<= c1 * v[b] + c2 * v[d]; break;
        case 4:
          v[a] <<= v[b] - c1 * v[d];
          v[a] -= c2 * v[e];
          break;
        case 5: record(Storm::dot_product(v[a], v[b])); break;
        case 6: record(Storm::norm_2(v[a])); break;
        case 7:
        case 8:
          if (b == a) b = (a + 1) % n_vecs;
          fvm.mul(v[a], v[b]);
          break;
        case 9: v[a] <<= v[b] + c1 * (v[d] - c2 * v[e]); break; // nested: not a chain
        case 10: Storm::fill_with(v[a], c1); break;
        case 11: v[a] /= (2.0 + c1); break;
        case 12: std::swap(v[a], v[b]); break;
        case 13:
          if (b != a) v[a].assign(v[b], (rng() & 1) != 0);
          break;
        case 14: {
          const std::vector<double> h = v[a].to_host();
          double sum = 0.0;
          for (double x : h) sum += x;
          record(sum);
          break;
        }
        case 15: v[a] <<= v[b]; break;
        case 16: // stormDivGrad as the playground calls it (faithful-form operators only; skipped otherwise)
          if (b == a) b = (a + 1) % n_vecs;
          if (with_accumulate) Storm::B200::div_grad(fvm, v[a], 1.0e-4 * c1, v[b]); // scaled: the operator's norm is ~1e4
          break;
        case 17: // omega = <t,r>/<t,t> of BiCGStab: <y,y> rides on the apply and brings <y,x> along
          if (b == a) b = (a + 1) % n_vecs;
          fvm.mul(v[a], v[b]);
          record(Storm::norm_2(v[a]));
          record(Storm::dot_product(v[a], v[b]));
          break;
        case 18: // <r~, v> after v = A p, with an update of r~ possibly still queued
          if (b == a) b = (a + 1) % n_vecs;
          if (d == a) d = b;
          v[d] += c2 * v[e == a ? b : e];
          fvm.mul(v[a], v[b]);
          record(Storm::dot_product(v[d], v[a]));
          break;
        default: // the preconditioner slot (coefficient-form operators with a diagonal only)
          if (with_jacobi) Storm::JacobiPreconditioner{fvm}.mul(v[a], v[b]);
          break;
      }
    }
    for (int k = 0; k < n_vecs; ++k) v[(size_t) k].download(h_final + (size_t) k * n);
    *n_record = nr;
    return 0;
  } catch (const std::exception& ex) {
    g_error = ex.what();
    return -2;
  }
}

// Opt-in statement grouping of the generic path (Storm::B200::set_statement_grouping, DeviceVector.hpp): chain-shaped
// statements are queued and launched as sb_eval_group together with the reduction that follows them.
// 0: off (default), 1: on, 2: on + dependency-aware scheduling (consumers launch only the statements they depend on).
DROPIN_API void dropin_set_statement_grouping(int on) { Storm::B200::set_statement_grouping(on != 0, on == 2); }

// Reset the engine behind fill_randomly(DeviceVector&) to the reference's initial state.
DROPIN_API void dropin_reset_rng(void) { Storm::B200::random_engine() = std::mt19937_64{}; }

// Misuse checks exercised by the tests: mixing sizes must throw, not corrupt memory.
DROPIN_API int dropin_selftest_errors(sb_ctx* ctx) {
  int caught = 0;
  try {
    DeviceVector a{ctx, 10}, b{ctx, 11};
    a += b;
  } catch (const std::runtime_error&) {
    ++caught;
  }
  try {
    DeviceVector a{ctx, 10}, b{ctx, 11};
    (void) Storm::dot_product(a, b);
  } catch (const std::runtime_error&) {
    ++caught;
  }
  try {
    DeviceVector a{ctx, 8}, b{ctx, 8}, c{ctx, 8}, d{ctx, 8}, e{ctx, 8};
    a <<= ((b + c) + (d + e)) + a; // five distinct vectors: over the evaluator's operand limit
  } catch (const std::runtime_error&) {
    ++caught;
  }
  return caught;
}

} // extern "C"
