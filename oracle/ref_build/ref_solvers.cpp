// TEST INFRASTRUCTURE (oracle/_ref build) -- not product code.
//
// Compiles the reference's OWN solver headers, unmodified, from where they lie under
// /root/reference/source (nothing is copied into this repo) and exposes them behind a tiny C
// entry point so the Python tests and bench.py's `--impl reference` leg can run them:
//   Storm/Solvers/SolverCg.hpp, SolverCgs.hpp, SolverBiCgStab.hpp (BiCGStab + BiCGStab(l)),
//   SolverGmres.hpp (GMRES + FGMRES), SolverTfqmr.hpp (TFQMR + TFQMR1), SolverIdrs.hpp,
//   SolverRichardson.hpp, driven through Storm/Solvers/Solver.hpp:116-147.
//
// The legacy solver headers do not compile as shipped; the non-invasive recipe of
// SURVEY.md F5 / Appendix B is used: (i) an empty Storm/Bittern/MatrixDense.hpp shadows the real one
// on the include path (this TU contains no mesh code), (ii) three vanished names are supplied
// before the solver headers are included, (iii) the vector type defines its own scalar *= and /=.
//
// The vector type is a plain host array, so every arithmetic statement executed is the
// reference's: Bittern lazy expressions evaluated per element in index order, sequential
// reductions from 0.0. Two additions, both outside the arithmetic:
//   * dot_product / norm_2 on HostVec are intercepted by non-template overloads that forward to
//     the reference's generic templates (sequential mode) and append the result to a trace;
//     in ORC_RED_TREE mode they use the oracle's restatement of the GPU reduction tree instead,
//     which isolates reduction-order effects.
//   * fill_randomly on HostVec uses a resettable engine with the reference's construction
//     (std::mt19937_64{} + uniform_real_distribution(0,1), MatrixAlgorithms.hpp:140-153);
//     ref_fill_randomly_generic() exposes the reference template itself so tests can check the
//     two produce the same stream.
//
// Build: oracle/Makefile. Flags are part of the contract:
//   g++ -std=c++23 -O2 -ffp-contract=off   (no -march, no -ffast-math)

#include <Storm/Bittern/Matrix.hpp>

#include <array>
#include <cstring>
#include <memory>
#include <random>
#include <string>
#include <tuple>
#include <vector>

#include "../sb_oracle.h"

namespace Storm {

// (ii) compat names the legacy headers expect (Solvers/MatrixDense.hpp:150,174; Solver.hpp:281).
template<class R, class C>
using MatrixShape = std::tuple<R, C>;
template<class M>
constexpr auto& fill_with(M& m, double s) {
  return fill(m, s);
}
constexpr auto make_diagonal_matrix(auto shape, auto s) {
  return eye<double>(shape, double(s));
}

struct RefTrace {
  int mode = ORC_RED_SEQ;
  double* trace = nullptr;
  size_t cap = 0, count = 0;
  void push(double v) {
    if (trace != nullptr && count < cap) trace[count] = v;
    ++count;
  }
};
inline RefTrace g_trace;
inline std::mt19937_64 g_random_engine{};

// (iii) the host vector: rank-2 shape {n, 1} like Feathers/Field.hpp:77-79; assign() zero-fills
// like Field.hpp:82-84 (IDR(s) relies on it).
struct HostVec final : TargetMatrixInterface<HostVec> {
  std::vector<double> d;
  auto shape() const noexcept {
    return std::array<size_t, 2>{d.size(), 1};
  }
  void assign(const HostVec& o, bool copy = true) {
    d.assign(o.d.size(), 0.0);
    if (copy) d = o.d;
  }
  double& operator()(size_t i, size_t = 0) noexcept {
    return d[i];
  }
  const double& operator()(size_t i, size_t = 0) const noexcept {
    return d[i];
  }
  using TargetMatrixInterface<HostVec>::operator=;
  // works around the broken scalar operators of MatrixTarget.hpp:96-105 (F4)
  HostVec& operator*=(double s) {
    for (auto& v : d) v *= s;
    return *this;
  }
  HostVec& operator/=(double s) {
    for (auto& v : d) v /= s;
    return *this;
  }
  using TargetMatrixInterface<HostVec>::operator*=;
  using TargetMatrixInterface<HostVec>::operator/=;
};

// Interception of the reductions. Non-template overloads for every cv combination, because the
// generic Bittern entry points take forwarding references (SURVEY.md 8b "overload hazard").
inline double ref_dot_impl(const HostVec& a, const HostVec& b) {
  double v;
  if (g_trace.mode != ORC_RED_SEQ) {
    v = orc_dot((int64_t) a.d.size(), a.d.data(), b.d.data(), g_trace.mode);
  } else {
    v = dot_product<const HostVec&, const HostVec&>(a, b); // the reference template
  }
  g_trace.push(v);
  return v;
}
inline double ref_norm_impl(const HostVec& a) {
  double v;
  if (g_trace.mode != ORC_RED_SEQ) {
    v = orc_norm2((int64_t) a.d.size(), a.d.data(), g_trace.mode);
  } else {
    v = norm_2<const HostVec&>(a); // the reference template
  }
  g_trace.push(v);
  return v;
}
inline double dot_product(HostVec& a, HostVec& b) { return ref_dot_impl(a, b); }
inline double dot_product(HostVec& a, const HostVec& b) { return ref_dot_impl(a, b); }
inline double dot_product(const HostVec& a, HostVec& b) { return ref_dot_impl(a, b); }
inline double dot_product(const HostVec& a, const HostVec& b) { return ref_dot_impl(a, b); }
inline double norm_2(HostVec& a) { return ref_norm_impl(a); }
inline double norm_2(const HostVec& a) { return ref_norm_impl(a); }

inline HostVec& fill_randomly(HostVec& out) {
  std::uniform_real_distribution<double> distribution{0.0, 1.0};
  for (auto& v : out.d) v = distribution(g_random_engine);
  return out;
}

} // namespace Storm

#include <Storm/Solvers/SolverBiCgStab.hpp>
#include <Storm/Solvers/SolverCg.hpp>
#include <Storm/Solvers/SolverCgs.hpp>
#include <Storm/Solvers/SolverGmres.hpp>
#include <Storm/Solvers/SolverIdrs.hpp>
#include <Storm/Solvers/SolverRichardson.hpp>
#include <Storm/Solvers/SolverTfqmr.hpp>
// JFNK (SolverNewton.hpp:101-173) uses BiCgStabSolver and std::numeric_limits without including them
// ("fix the missing include"): included after SolverBiCgStab.hpp here, untouched.
#include <limits>
#include <Storm/Solvers/SolverNewton.hpp>

// The product's polynomial preconditioner is a template in the reference's own vector vocabulary: compiled here on the
// host vector, it is the checker of the very same template on the device vector (bit for bit, same reduction tree).
#include <Storm/B200/ChebyshevPreconditioner.hpp>

namespace {

using Storm::HostVec;

struct ref_opts {
  int64_t num_iterations;
  double abs_tol;
  double rel_tol;
  int64_t num_inner_iterations; // <= 0: keep the solver's default
  int32_t reduction_mode;
  double relaxation_factor; // Richardson only; <= 0 keeps the default 1e-4
  // optional preconditioner in the reference's pre_op slot (Solver.hpp:74-75): y = P x through a callback
  void (*pre_fn)(void* user, double* y, const double* x, size_t n);
  void* pre_user;
  int32_t pre_side;         // 0: Left, 1: Right (reference default), 2: Symmetric
  // pre_kind 3: Storm::ChebyshevPreconditioner<HostVec> (stormruler_b200/host/Storm/B200/ChebyshevPreconditioner.hpp) in
  // the pre_op slot instead of the callback; the parameters as in dropin_opts (<= 0: class defaults)
  int32_t pre_kind;
  int32_t cheb_degree;
  int32_t cheb_power_iterations;
  double cheb_eig_ratio;
};

struct ref_report {
  int32_t converged;
  int64_t iterations;
  double abs_err;
  double rel_err;
  int64_t n_hist;
  int64_t n_trace;
  int64_t n_apply;
};

typedef void (*ref_apply_fn)(void* user, double* y, const double* x, size_t n);

// Operator that forwards to a C callback and samples the solver's public progress fields
// (Solver.hpp:66-69) on every call: residual-history capture, method (2) of SURVEY.md 8c.
template<class SolverT>
struct CallbackOperator final : Storm::Operator<HostVec> {
  ref_apply_fn fn;
  void* user;
  const SolverT* solver;
  double* hist;
  int64_t hist_cap;
  mutable int64_t n_apply = 0;
  mutable int64_t max_seen = -1;
  void mul(HostVec& y, const HostVec& x) const override {
    const int64_t it = (int64_t) solver->iteration;
    if (hist != nullptr && it < hist_cap) hist[it] = solver->absolute_error;
    if (it > max_seen) max_seen = it;
    fn(user, y.d.data(), x.d.data(), x.d.size());
    ++n_apply;
  }
};

struct CallbackPreconditioner final : Storm::Preconditioner<HostVec> {
  ref_apply_fn fn;
  void* user;
  void mul(HostVec& y, const HostVec& x) const override { fn(user, y.d.data(), x.d.data(), x.d.size()); }
  void conj_mul(HostVec& x, const HostVec& y) const override { fn(user, x.d.data(), y.d.data(), y.d.size()); }
};

template<class SolverT>
int run(size_t n, ref_apply_fn fn, void* user, const double* b, double* x, const ref_opts* o,
        ref_report* rep, double* hist, int64_t hist_cap, double* trace, int64_t trace_cap) {
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
  if (o->pre_kind == 3) {
    auto cheb = std::make_unique<Storm::ChebyshevPreconditioner<HostVec>>();
    if (o->cheb_degree > 0) cheb->degree = (size_t) o->cheb_degree;
    if (o->cheb_power_iterations > 0) cheb->num_power_iterations = (size_t) o->cheb_power_iterations;
    if (o->cheb_eig_ratio > 0.0) cheb->eig_ratio = o->cheb_eig_ratio;
    solver.pre_op = std::move(cheb);
    solver.pre_side = o->pre_side == 0 ? Storm::PreconditionerSide::Left
                                       : (o->pre_side == 2 ? Storm::PreconditionerSide::Symmetric : Storm::PreconditionerSide::Right);
  } else if (o->pre_fn != nullptr) {
    auto pre = std::make_unique<CallbackPreconditioner>();
    pre->fn = o->pre_fn, pre->user = o->pre_user;
    solver.pre_op = std::move(pre);
    solver.pre_side = o->pre_side == 0 ? Storm::PreconditionerSide::Left
                                       : (o->pre_side == 2 ? Storm::PreconditionerSide::Symmetric : Storm::PreconditionerSide::Right);
  }
  HostVec xv, bv;
  xv.d.assign(x, x + n);
  bv.d.assign(b, b + n);
  Storm::g_trace = Storm::RefTrace{o->reduction_mode, trace, (size_t) trace_cap, 0};
  CallbackOperator<SolverT> op;
  op.fn = fn, op.user = user, op.solver = &solver, op.hist = hist, op.hist_cap = hist_cap;
  const bool converged = solver.solve(xv, bv, op);
  // hist[k] was written at the first apply of iteration k with the error returned by
  // iteration k-1; the last entry comes from the public field after solve() returns.
  const int64_t it = (int64_t) solver.iteration;
  if (hist != nullptr && it < hist_cap) hist[it] = solver.absolute_error;
  std::memcpy(x, xv.d.data(), n * sizeof(double));
  rep->converged = converged ? 1 : 0;
  rep->iterations = it;
  rep->abs_err = solver.absolute_error;
  rep->rel_err = solver.relative_error;
  rep->n_hist = it + 1;
  rep->n_trace = (int64_t) Storm::g_trace.count;
  rep->n_apply = op.n_apply;
  Storm::g_trace = Storm::RefTrace{};
  return 0;
}

} // namespace

extern "C" {

// Solver names: cg cgs bicgstab bicgstabl gmres fgmres tfqmr tfqmr1 idrs richardson jfnk.
int ref_solve(const char* name, size_t n, ref_apply_fn fn, void* user, const double* b, double* x,
              const ref_opts* o, ref_report* rep, double* hist, int64_t hist_cap, double* trace,
              int64_t trace_cap) {
  const std::string s{name};
#define REF_CASE(key, T) \
  if (s == key) return run<Storm::T<HostVec>>(n, fn, user, b, x, o, rep, hist, hist_cap, trace, trace_cap)
  REF_CASE("cg", CgSolver);
  REF_CASE("cgs", CgsSolver);
  REF_CASE("bicgstab", BiCgStabSolver);
  REF_CASE("bicgstabl", BiCgStabLSolver);
  REF_CASE("gmres", GmresSolver);
  REF_CASE("fgmres", FgmresSolver);
  REF_CASE("tfqmr", TfqmrSolver);
  REF_CASE("tfqmr1", Tfqmr1Solver);
  REF_CASE("idrs", IdrsSolver);
  REF_CASE("richardson", RichardsonSolver);
  REF_CASE("jfnk", JfnkSolver);
#undef REF_CASE
  return -1;
}

// solve_non_uniform (Solver.hpp:271-292), the reference's own function template on the host vector, for the affine
// operator A(x) = L x + shift (L through the callback). Mirrors dropin_solve_non_uniform of the product's drop-in TU.
int ref_solve_non_uniform(const char* name, size_t n, ref_apply_fn fn, void* user, const double* b, const double* shift,
                          double* x, const ref_opts* o, ref_report* rep, double* trace, int64_t trace_cap) {
  const std::string s{name};
  HostVec xv, bv, sv;
  xv.d.assign(x, x + n), bv.d.assign(b, b + n);
  if (shift != nullptr) sv.d.assign(shift, shift + n); // NULL: the callback is the affine operator itself
  int64_t n_apply = 0;
  const auto affine = Storm::make_operator<HostVec>([&](HostVec& y, const HostVec& in) {
    fn(user, y.d.data(), in.d.data(), n);
    if (shift != nullptr) y += sv;
    ++n_apply;
  });
  Storm::g_trace = Storm::RefTrace{o->reduction_mode, trace, (size_t) trace_cap, 0};
  auto run = [&](auto solver) {
    solver.num_iterations = (size_t) o->num_iterations;
    solver.absolute_error_tolerance = o->abs_tol, solver.relative_error_tolerance = o->rel_tol;
    const bool converged = Storm::solve_non_uniform(solver, xv, bv, *affine);
    std::memcpy(x, xv.d.data(), n * sizeof(double));
    rep->converged = converged ? 1 : 0;
    rep->iterations = (int64_t) solver.iteration;
    rep->abs_err = solver.absolute_error, rep->rel_err = solver.relative_error;
    rep->n_hist = 0, rep->n_trace = (int64_t) Storm::g_trace.count, rep->n_apply = n_apply;
    Storm::g_trace = Storm::RefTrace{};
    return 0;
  };
  if (s == "cg") return run(Storm::CgSolver<HostVec>{});
  if (s == "bicgstab") return run(Storm::BiCgStabSolver<HostVec>{});
  if (s == "gmres") return run(Storm::GmresSolver<HostVec>{});
  if (s == "idrs") return run(Storm::IdrsSolver<HostVec>{});
  Storm::g_trace = Storm::RefTrace{};
  return -1;
}

// Reset the engine behind fill_randomly(HostVec&) to the reference's initial state
// (default-seeded std::mt19937_64; the reference's own engine is a function-local static that
// cannot be reset).
void ref_reset_rng(void) {
  Storm::g_random_engine = std::mt19937_64{};
}

// The reference's own generic fill_randomly template (MatrixAlgorithms.hpp:140-153), for
// checking the stream above. Its static engine advances across calls.
void ref_fill_randomly_generic(size_t n, double* out) {
  struct Plain final : Storm::TargetMatrixInterface<Plain> {
    std::vector<double> d;
    auto shape() const noexcept { return std::array<size_t, 2>{d.size(), 1}; }
    double& operator()(size_t i, size_t = 0) noexcept { return d[i]; }
    const double& operator()(size_t i, size_t = 0) const noexcept { return d[i]; }
  } p;
  p.d.assign(n, 0.0);
  Storm::fill_randomly(p);
  std::memcpy(out, p.d.data(), n * sizeof(double));
}

// Reference BLAS-1 known answers (tests/unit/BitternReductions.cpp) evaluated by the
// reference templates on HostVec, for the golden file.
double ref_dot(size_t n, const double* a, const double* b) {
  HostVec av, bv;
  av.d.assign(a, a + n), bv.d.assign(b, b + n);
  return Storm::dot_product<const HostVec&, const HostVec&>(av, bv);
}
double ref_norm2(size_t n, const double* a) {
  HostVec av;
  av.d.assign(a, a + n);
  return Storm::norm_2<const HostVec&>(av);
}

const char* ref_build_info(void) {
  return "StormRuler solver headers compiled verbatim; g++ " __VERSION__
#ifdef __OPTIMIZE__
         " optimized"
#endif
#ifdef __FAST_MATH__
         " FAST_MATH"
#endif
#ifdef __FMA__
         " FMA"
#endif
      ;
}

} // extern "C"
