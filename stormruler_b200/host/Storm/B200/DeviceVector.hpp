// Storm/B200/DeviceVector.hpp -- the C++23 drop-in: StormRuler's solver headers, unchanged, on a B200.
//
// StormRuler's extension point for the Krylov path is not an FFI but two C++ contracts
// (SURVEY.md 8b):
//   * the `legacy_vector_like` concept            Storm/Solvers/Operator.hpp:39-45
//   * `Operator<Vector>::mul` (virtual)           Storm/Solvers/Operator.hpp:66-74
// plus the Bittern free functions / operators that the solver bodies find by overload resolution
// on the vector type (Storm/Bittern/MatrixAlgorithms.hpp, MatrixMath.hpp, MatrixTarget.hpp).
// This header supplies a vector type whose storage lives in HBM (`Storm::DeviceVector`), the lazy
// expression type its operators build (`Storm::DevExpr`, flattened to the postfix program of
// sb_eval so the kernel evaluates in the reference's association order), the reductions, and an
// `Operator<DeviceVector>` over an uploaded FVM operator (`Storm::FvmOperator`). With it,
//     Storm::BiCgStabLSolver<Storm::DeviceVector> solver;   solver.solve(x, b, op);
// compiles from the reference's own SolverBiCgStab.hpp and runs every vector statement as a
// hand-written sm_100a kernel through the C ABI of libstormb200.so (include/stormb200.h). There
// is no host fallback: every operation goes to the device or throws.
//
// Include order in a translation unit:
//     -I<repo>/stormruler_b200/host/compat  (first: shadows Storm/Bittern/MatrixDense.hpp)
//     #include <Storm/B200/DeviceVector.hpp>      // before any Storm/Solvers/Solver*.hpp
//     #include <Storm/Solvers/SolverCg.hpp> ...
//
// Overload-resolution contract (SURVEY.md 8b "hazard", probed): the generic Bittern entry points
// take forwarding references, which beat a `const DeviceVector&` parameter for non-const lvalues.
// Therefore every vector-taking function below exists for all const / non-const lvalue
// combinations as NON-template overloads; `DevExpr` is deliberately not a Storm::matrix (no
// shape()), and DeviceVector is deliberately not a Storm::output_matrix (its element accessor
// returns a by-value proxy that the scalar functors reject), so an accidental fall-through to the
// per-element generic path is a compile error instead of a silently slow run.
#pragma once

#include <Storm/Bittern/Matrix.hpp>

#include <algorithm>
#include <array>
#include <cmath>
#include <concepts>
#include <cstdint>
#include <cstring>
#include <initializer_list>
#include <random>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "../../../../include/stormb200.h"

namespace Storm {

// ---- names the legacy solver headers expect but the tree no longer defines ------
template<class R, class C>
using MatrixShape = std::tuple<R, C>; // Solvers/MatrixDense.hpp:150
template<class M>
constexpr auto& fill_with(M& m, double s) { // Solver.hpp:281, SolverBiCgStab.hpp:224, SolverTfqmr.hpp:77
  return fill(m, s);
}
constexpr auto make_diagonal_matrix(auto shape, auto s) { // Solvers/MatrixDense.hpp:174
  return eye<double>(shape, double(s));
}

namespace B200 {

/// Throws std::runtime_error carrying the library's message when a C-ABI call fails.
inline void check(int rc, const char* what) {
  if (rc != SB_OK) {
    throw std::runtime_error(std::string{"stormb200: "} + what + " failed (" + std::to_string(rc) +
                             "): " + sb_last_error());
  }
}

/// Optional observer of every reduction result (dot_product / norm_2), in call order: the hook the
/// parity tests use to compare the full scalar trace with the CPU reference.
using ReductionObserver = void (*)(void* user, double value);
inline ReductionObserver g_observer = nullptr;
inline void* g_observer_user = nullptr;
inline void observe(double v) {
  if (g_observer != nullptr) g_observer(g_observer_user, v);
}

/// Engine behind fill_randomly(DeviceVector&): the reference's construction, a default-seeded
/// std::mt19937_64 with a static lifetime (MatrixAlgorithms.hpp:140-153), made resettable.
inline std::mt19937_64& random_engine() {
  static std::mt19937_64 engine{};
  return engine;
}

/// ---- statement grouping (on by default, with dependency-aware scheduling) ----------------------------------------
/// Measured on the B200 at 10.1 M cells (profiles/r02_solver_sweep_*.json), reference templates unchanged: BiCGStab
/// 1593 -> 1877 it/s, IDR(4) 1224 -> 1591, CG 3286 -> 3537, BiCGStab(2) 1350 -> 1450, GMRES(50) 429 -> 446; no solver
/// slower. `set_statement_grouping(false)` restores one launch per statement.
/// With grouping on the element-wise statements the solver templates issue one by one are not
/// launched one by one: every statement that is a linear-combination chain `y = ((base +- c0*x0) +- c1*x1) ...`
/// (all vector updates of the reference solvers except the nested `r + beta*(p - omega*v)` forms) is queued, and
/// the queue is handed to the device as ONE sb_eval_group launch when something needs its result: a reduction (which
/// joins the group: statements + dot in one launch and one synchronisation), an operator apply, a statement of another
/// shape, a host access, a vector going away. Element-wise statements are independent per element and the group
/// runs them in order, so nothing changes bit-wise (tests/test_dropin_emulated.py runs every reference solver both
/// ways). Code that hands `DeviceVector::data()` to the C ABI itself must call `B200::flush()` first.
struct StatementQueue {
  bool enabled = true;
  /// Dependency-aware scheduling: a consumer (apply, reduction, other statement, host access) launches only the queued
  /// statements it conflicts with -- and those they in turn depend on -- instead of everything; the rest stay queued
  /// and join a later group (BiCGStab's `x += alpha*p` waits for `x += omega*r`, CG's for the direction update).
  /// Statements are only ever moved past statements and launches they share no vector with in a conflicting role
  /// (read-after-write, write-after-read, write-after-write), so every vector sees its operations in program order.
  bool reorder = true;
  static constexpr size_t kMaxQueued = 24;
  sb_ctx* ctx = nullptr;
  size_t n = 0;
  std::vector<sb_chain> stmts;
  /// An operator apply that has been asked for but not launched yet: if the next thing the solver does is a dot product
  /// of its output with its input or with a third vector (`lin_op.mul(z, p); dot_product(p, z)`, SolverCg.hpp:96-97),
  /// the two go to the device as one sb_apply_dot. Invariant: no queued statement writes the deferred apply's input or
  /// touches its output (such statements were launched when the apply was deferred; a statement queued after it
  /// launches it first). Without `reorder` nothing at all is queued while an apply is deferred.
  struct {
    bool active = false;
    sb_ctx* ctx = nullptr;
    const sb_op* op = nullptr;
    const double* x = nullptr;
    double* y = nullptr;
    size_t n = 0;
  } apply;

  /// <y, x> computed by the same kernel as <y, y> (sb_apply_dot_yy_yx), kept for the dot product that follows in
  /// `safe_divide(dot_product(t, r), dot_product(t, t))` (g++ evaluates the second argument first). Dropped by anything
  /// that could change x or y: every such path goes through flush(), try_enqueue() or defer_apply().
  struct {
    bool valid = false;
    const double* x = nullptr;
    const double* y = nullptr;
    size_t n = 0;
    double yx = 0.0;
  } spare;

  void launch_apply() {
    if (!apply.active) return;
    apply.active = false;
    check(sb_apply(apply.ctx, apply.op, apply.x, apply.y), "sb_apply");
  }
  void defer_apply(sb_ctx* c, const sb_op* op, const double* x, double* y, size_t len) {
    const double* w = y;
    flush_for(&x, 1, &w, 1);
    launch_apply(); // an earlier apply that is still pending comes first
    apply.active = true, apply.ctx = c, apply.op = op, apply.x = x, apply.y = y, apply.n = len;
  }

  static bool reads(const sb_chain& s, const double* p) {
    if (s.base == p) return true;
    for (int t = 0; t < s.n_terms; ++t)
      if (s.x[t] == p) return true;
    return false;
  }
  static bool depends(const sb_chain& a, const sb_chain& b) { return a.y == b.y || reads(a, b.y) || reads(b, a.y); }

  /// Launch the queued statements that conflict with a consumer reading R and writing W (all of them unless `reorder`),
  /// closed under dependencies on earlier statements, in program order, with an optional dot riding on the last
  /// launch. Returns how many statements were launched; the others stay queued in order.
  size_t launch_conflicting(const double* const* R, int nR, const double* const* W, int nW, const double* dot_a,
                            const double* dot_b, double* dot_out) {
    const size_t total = stmts.size();
    std::vector<char> take(total, reorder ? 0 : 1);
    if (reorder) {
      for (size_t i = 0; i < total; ++i) {
        const sb_chain& s = stmts[i];
        for (int k = 0; k < nR && !take[i]; ++k) take[i] = s.y == R[k];                 // the consumer reads what s writes
        for (int k = 0; k < nW && !take[i]; ++k) take[i] = s.y == W[k] || reads(s, W[k]); // ... or overwrites s's target / source
      }
      for (bool changed = true; changed;) {
        changed = false;
        for (size_t i = 0; i < total; ++i) {
          if (!take[i]) continue;
          for (size_t j = 0; j < i; ++j)
            if (!take[j] && depends(stmts[j], stmts[i])) take[j] = 1, changed = true;
        }
      }
    }
    std::vector<sb_chain> sel, keep;
    for (size_t i = 0; i < total; ++i) (take[i] ? sel : keep).push_back(stmts[i]);
    stmts.swap(keep); // a failing launch must not leave the launched statements queued for a second attempt
    const size_t ns = sel.size();
    if (ns == 0) return 0;
    const size_t tail = (ns - 1) / SB_GROUP_MAX_STMT * SB_GROUP_MAX_STMT; // first statement of the last launch
    for (size_t s0 = 0; s0 < ns; s0 += SB_GROUP_MAX_STMT) {
      const size_t cnt = std::min<size_t>(ns - s0, SB_GROUP_MAX_STMT);
      const bool with_dot = dot_a != nullptr && s0 == tail;
      check(sb_eval_group(ctx, n, (int) cnt, sel.data() + s0, with_dot ? 1 : 0, with_dot ? &dot_a : nullptr,
                          with_dot ? &dot_b : nullptr, with_dot ? dot_out : nullptr),
            "sb_eval_group");
    }
    return ns;
  }
  /// What must happen before a consumer that reads R and writes W runs.
  void flush_for(const double* const* R, int nR, const double* const* W, int nW) {
    spare.valid = false;
    if (apply.active) {
      bool conflict = !reorder;
      for (int k = 0; k < nR; ++k) conflict |= R[k] == apply.y;
      for (int k = 0; k < nW; ++k) conflict |= W[k] == apply.y || W[k] == apply.x;
      if (conflict) launch_apply();
    }
    if (!stmts.empty()) launch_conflicting(R, nR, W, nW, nullptr, nullptr, nullptr);
  }

  void flush() {
    spare.valid = false;
    launch_apply();
    const bool was = reorder;
    reorder = false; // everything
    try {
      if (!stmts.empty()) launch_conflicting(nullptr, 0, nullptr, 0, nullptr, nullptr, nullptr);
    } catch (...) {
      reorder = was;
      throw;
    }
    reorder = was;
  }
  /// <a, b> over the vectors as they are after the queued statements: the statements and the dot in one launch.
  double reduce(sb_ctx* c, const double* a, const double* b, size_t len) {
    double v = 0.0;
    if (spare.valid) {
      spare.valid = false;
      if (len == spare.n && ((a == spare.y && b == spare.x) || (a == spare.x && b == spare.y))) return spare.yx;
    }
    if (apply.active) { // the dot rides on the deferred apply when exactly one of its operands is the apply's output
      const bool ya = a == apply.y, yb = b == apply.y;
      if (c == apply.ctx && len == apply.n && ya && yb) { // <y, y>: take <y, x> along for the dot that usually follows
        double both[2] = {0.0, 0.0};
        apply.active = false;
        check(sb_apply_dot_yy_yx(apply.ctx, apply.op, apply.x, apply.y, both), "sb_apply_dot_yy_yx");
        spare.valid = true, spare.x = apply.x, spare.y = apply.y, spare.n = len, spare.yx = both[1];
        return both[0];
      }
      if (c == apply.ctx && len == apply.n && ya != yb) {
        const double* u = ya ? b : a;
        // statements still queued never touch the apply's x or y, but they may write u (dependency-aware mode)
        if (!stmts.empty()) launch_conflicting(&u, 1, nullptr, 0, nullptr, nullptr, nullptr);
        apply.active = false;
        check(sb_apply_dot(apply.ctx, apply.op, apply.x, apply.y, u == apply.x ? nullptr : u, &v), "sb_apply_dot");
        return v;
      }
      launch_apply();
    }
    if (!enabled || stmts.empty() || c != ctx || len != n) {
      flush();
      check(sb_dot(c, a, b, len, &v), "sb_dot");
      return v;
    }
    const double* R[2] = {a, b};
    if (launch_conflicting(R, 2, nullptr, 0, a, b, &v) == 0) check(sb_dot(c, a, b, len, &v), "sb_dot");
    return v;
  }
  /// Queue `y (aop)= expr` if it is a chain; false: the caller flushes and launches it directly.
  bool try_enqueue(sb_ctx* c, double* y, size_t len, int aop, const sb_expr& e) {
    if (!enabled || len == 0) return false;
    struct Node {
      int kind; // 0 vector, 1 scalar, 2 product c*x, 3 chain
      const double* v;
      double s;
      sb_chain ch;
    };
    Node st[SB_EXPR_MAX_OPS];
    int sp = 0;
    auto as_term = [](const Node& nd, const double*& x, double& coef) {
      if (nd.kind == 0) return x = nd.v, coef = 1.0, true; // a +- x == a +- 1.0*x exactly
      if (nd.kind == 2) return x = nd.v, coef = nd.s, true;
      return false;
    };
    auto start_chain = [&](const Node& nd, sb_chain& ch) {
      ch = sb_chain{};
      if (nd.kind == 0) return ch.base = nd.v, true;
      if (nd.kind == 2) return ch.x[0] = nd.v, ch.c[0] = nd.s, ch.sub[0] = 0, ch.n_terms = 1, true;
      return false;
    };
    for (int k = 0; k < e.n_ops; ++k) {
      const int op = e.ops[k];
      if (op <= SB_OP_VEC3) {
        st[sp++] = Node{0, e.vec[op], 0.0, {}};
      } else if (op >= SB_OP_SCAL0 && op <= SB_OP_SCAL3) {
        st[sp++] = Node{1, nullptr, e.scal[op - SB_OP_SCAL0], {}};
      } else if (op == SB_OP_MUL) {
        const Node r = st[--sp], l = st[--sp];
        if (l.kind == 1 && r.kind == 0) st[sp++] = Node{2, r.v, l.s, {}};
        else if (l.kind == 0 && r.kind == 1) st[sp++] = Node{2, l.v, r.s, {}}; // x*c == c*x
        else return false;
      } else if (op == SB_OP_ADD || op == SB_OP_SUB) {
        const Node r = st[--sp];
        Node l = st[--sp];
        const double* x = nullptr;
        double coef = 0.0;
        if (!as_term(r, x, coef)) return false; // a +- (chain): another association
        if (l.kind != 3) {
          sb_chain ch;
          if (!start_chain(l, ch)) return false;
          l = Node{3, nullptr, 0.0, ch};
        }
        if (l.ch.n_terms == SB_GROUP_MAX_TERMS) return false;
        l.ch.x[l.ch.n_terms] = x, l.ch.c[l.ch.n_terms] = coef, l.ch.sub[l.ch.n_terms] = op == SB_OP_SUB ? 1 : 0;
        l.ch.n_terms++;
        st[sp++] = l;
      } else {
        return false; // division, negation
      }
    }
    if (sp != 1) return false;
    sb_chain ch{};
    const Node& top = st[0];
    if (aop == SB_ASSIGN) {
      if (top.kind == 3) ch = top.ch;
      else if (top.kind == 0 || top.kind == 2) ch.x[0] = top.v, ch.c[0] = top.kind == 2 ? top.s : 1.0, ch.n_terms = 1;
      else return false;
    } else if (aop == SB_ADD_ASSIGN || aop == SB_SUB_ASSIGN) {
      const double* x = nullptr;
      double coef = 0.0;
      if (!as_term(top, x, coef)) return false; // y +- (a + b) is not (y +- a) +- b
      ch.base = y, ch.x[0] = x, ch.c[0] = coef, ch.sub[0] = aop == SB_SUB_ASSIGN ? 1 : 0, ch.n_terms = 1;
    } else {
      return false;
    }
    ch.y = y;
    spare.valid = false;
    launch_apply(); // the statement may read what a deferred apply writes
    if (!stmts.empty() && (c != ctx || len != n || stmts.size() >= kMaxQueued)) flush();
    ctx = c, n = len;
    // `y += a*p` right behind another update of the same y continues that chain ((y + ..) + a*p: the same roundings),
    // so y is read and written once by the group instead of twice
    if (!stmts.empty() && ch.base == y && stmts.back().y == y && stmts.back().n_terms + ch.n_terms <= SB_GROUP_MAX_TERMS) {
      bool reads_y = false;
      for (int t = 0; t < ch.n_terms; ++t) reads_y |= ch.x[t] == y;
      if (!reads_y) {
        sb_chain& last = stmts.back();
        for (int t = 0; t < ch.n_terms; ++t) {
          last.x[last.n_terms] = ch.x[t], last.c[last.n_terms] = ch.c[t], last.sub[last.n_terms] = ch.sub[t];
          last.n_terms++;
        }
        return true;
      }
    }
    stmts.push_back(ch);
    return true;
  }
};
/// One queue per host thread (one context = one device = one host thread at a time).
inline StatementQueue& statement_queue() {
  static thread_local StatementQueue q;
  return q;
}
/// Launch whatever is queued (no-op when nothing is, or when grouping is off).
inline void flush() {
  StatementQueue& q = statement_queue();
  q.spare.valid = false;
  if (!q.stmts.empty() || q.apply.active) q.flush();
}
/// Launch what a consumer reading `reads` and writing `writes` (device pointers) depends on; everything when the
/// dependency-aware mode is off.
inline void flush_for(std::initializer_list<const double*> reads, std::initializer_list<const double*> writes) {
  StatementQueue& q = statement_queue();
  q.spare.valid = false;
  if (!q.stmts.empty() || q.apply.active) q.flush_for(reads.begin(), (int) reads.size(), writes.begin(), (int) writes.size());
}
/// on: queue chain-shaped statements; reorder: additionally let consumers launch only what they depend on.
inline void set_statement_grouping(bool on, bool reorder = false) {
  flush();
  statement_queue().enabled = on;
  statement_queue().reorder = on && reorder;
}
inline bool statement_grouping() { return statement_queue().enabled; }

} // namespace B200

class DeviceVector;

/// By-value proxy returned by DeviceVector::operator(): explicit read only. Not arithmetic, not
/// assignable: the generic per-element Bittern algorithms cannot be instantiated with it.
struct DeviceElement {
  double value;
  double get() const noexcept { return value; }
};

/// Lazy vector expression: a postfix program over <= 4 distinct device vectors and <= 4 scalars
/// (the sb_expr of include/stormb200.h). Stands in for Bittern's MapMatrixView trees
/// (MatrixMath.hpp:44-87); evaluation order is exactly the order the C++ operators were applied.
class DevExpr {
public:

  DevExpr() = default;
  DevExpr(const DeviceVector& v); // NOLINT: implicit on purpose (a vector is an expression)

  static DevExpr scalar(double s) {
    DevExpr e;
    e._e.scal[0] = s;
    e._n_scal = 1;
    e.push(SB_OP_SCAL0);
    return e;
  }

  /// `lhs (op) rhs` with op in SB_OP_ADD..SB_OP_DIV.
  static DevExpr binary(const DevExpr& lhs, const DevExpr& rhs, uint8_t op) {
    DevExpr out = lhs;
    if (out._ctx == nullptr) out._ctx = rhs._ctx, out._n = rhs._n;
    if (lhs._ctx != nullptr && rhs._ctx != nullptr && (lhs._ctx != rhs._ctx || lhs._n != rhs._n)) {
      throw std::runtime_error("stormb200: expression mixes vectors of different size or context");
    }
    int vmap[SB_EXPR_MAX_VEC], smap[SB_EXPR_MAX_SCAL];
    for (int k = 0; k < rhs._n_vec; ++k) {
      int slot = -1;
      for (int q = 0; q < out._n_vec; ++q)
        if (out._e.vec[q] == rhs._e.vec[k]) slot = q;
      if (slot < 0) {
        if (out._n_vec == SB_EXPR_MAX_VEC) throw std::runtime_error("stormb200: expression uses more than 4 vectors");
        slot = out._n_vec++;
        out._e.vec[slot] = rhs._e.vec[k];
      }
      vmap[k] = slot;
    }
    for (int k = 0; k < rhs._n_scal; ++k) {
      int slot = -1; // a constant that occurs twice (the 2.0 and 1.0 of dF/dc, Playground.cpp:142-144) takes one slot
      for (int q = 0; q < out._n_scal; ++q)
        if (std::memcmp(&out._e.scal[q], &rhs._e.scal[k], sizeof(double)) == 0) slot = q;
      if (slot < 0) {
        if (out._n_scal == SB_EXPR_MAX_SCAL) throw std::runtime_error("stormb200: expression uses more than 4 scalars");
        slot = out._n_scal++;
        out._e.scal[slot] = rhs._e.scal[k];
      }
      smap[k] = slot;
    }
    for (int k = 0; k < rhs._e.n_ops; ++k) {
      const uint8_t o = rhs._e.ops[k];
      if (o <= SB_OP_VEC3) out.push(uint8_t(SB_OP_VEC0 + vmap[o - SB_OP_VEC0]));
      else if (o >= SB_OP_SCAL0 && o <= SB_OP_SCAL3) out.push(uint8_t(SB_OP_SCAL0 + smap[o - SB_OP_SCAL0]));
      else out.push(o);
    }
    out.push(op);
    return out;
  }

  DevExpr negated() const {
    DevExpr out = *this;
    out.push(SB_OP_NEG);
    return out;
  }

  const sb_expr& program() const noexcept { return _e; }
  sb_ctx* context() const noexcept { return _ctx; }
  size_t size() const noexcept { return _n; }

private:

  void push(uint8_t op) {
    if (_e.n_ops == SB_EXPR_MAX_OPS) throw std::runtime_error("stormb200: expression too long");
    _e.ops[_e.n_ops++] = op;
  }

  sb_expr _e{};
  int _n_vec = 0, _n_scal = 0;
  sb_ctx* _ctx = nullptr;
  size_t _n = 0;
};

/// Device-resident fp64 vector: the `Vector` of the solver templates on a B200. Replaces
/// CellField (Feathers/Field.hpp:60-114): contiguous doubles, shape {n, 1} (rank 2, :77-79),
/// assign() allocates ZERO-filled storage of the same shape (:82-84 -- IDR(s) relies on it).
/// Move-only; std::swap is a pointer swap (SolverBiCgStab.hpp:84 etc.).
class DeviceVector final {
public:

  DeviceVector() = default;
  DeviceVector(sb_ctx* ctx, size_t n) { allocate(ctx, n); }
  DeviceVector(sb_ctx* ctx, const std::vector<double>& host) {
    allocate(ctx, host.size());
    upload(host.data());
  }
  DeviceVector(const DeviceVector&) = delete;
  DeviceVector& operator=(const DeviceVector&) = delete;
  DeviceVector(DeviceVector&& o) noexcept : _ctx{o._ctx}, _d{o._d}, _n{o._n}, _owned{o._owned} {
    o._d = nullptr, o._n = 0, o._owned = false;
  }
  DeviceVector& operator=(DeviceVector&& o) noexcept {
    if (this != &o) {
      release();
      _ctx = o._ctx, _d = o._d, _n = o._n, _owned = o._owned;
      o._d = nullptr, o._n = 0, o._owned = false;
    }
    return *this;
  }
  ~DeviceVector() { release(); }

  /// Non-owning view of device memory that came from sb_vec_alloc elsewhere (e.g. Python).
  static DeviceVector view(sb_ctx* ctx, double* d, size_t n) {
    DeviceVector v;
    v._ctx = ctx, v._d = d, v._n = n, v._owned = false;
    return v;
  }

  // -- the legacy_vector_like surface (Operator.hpp:39-45) ---------------------------------------
  auto shape() const noexcept { return std::array<size_t, 2>{_n, 1}; }
  /// Debug accessor (one-element D2H copy). Exists so that Storm::matrix<DeviceVector> holds.
  DeviceElement operator()(size_t row, size_t = 0) const {
    double v = 0.0;
    B200::flush_for({_d}, {});
    B200::check(sb_vec_download(_ctx, _d + row, &v, 1), "sb_vec_download");
    return DeviceElement{v};
  }
  void assign(const DeviceVector& other, bool copy = true) {
    if (&other == this) return;
    release();
    allocate(other._ctx, other._n); // zero-filled
    if (copy && _n > 0) {
      B200::flush_for({other._d}, {_d});
      B200::check(sb_copy(_ctx, _d, other._d, _n), "sb_copy");
    }
  }

  // -- storage -----------------------------------------------------------------------------------
  sb_ctx* context() const noexcept { return _ctx; }
  double* data() noexcept { return _d; }
  const double* data() const noexcept { return _d; }
  size_t size() const noexcept { return _n; }
  void upload(const double* host) {
    B200::flush_for({}, {_d});
    B200::check(sb_vec_upload(_ctx, _d, host, _n), "sb_vec_upload");
  }
  void download(double* host) const {
    B200::flush_for({_d}, {});
    B200::check(sb_vec_download(_ctx, _d, host, _n), "sb_vec_download");
  }
  std::vector<double> to_host() const {
    std::vector<double> h(_n);
    if (_n > 0) download(h.data());
    return h;
  }

  // -- assignment operators (MatrixTarget.hpp:96-119) ---------------------------------------------
  DeviceVector& eval(int assign_op, const DevExpr& e) {
    if (e.context() != nullptr && (e.context() != _ctx || e.size() != _n)) {
      throw std::runtime_error("stormb200: assignment between vectors of different size or context");
    }
    if (B200::statement_queue().try_enqueue(_ctx, _d, _n, assign_op, e.program())) return *this;
    {
      const sb_expr& pr = e.program(); // reads: the expression's operands (and y for += -= *= /=); writes: y
      const double* rd[SB_EXPR_MAX_VEC + 1];
      int nr = 0;
      for (int k = 0; k < SB_EXPR_MAX_VEC; ++k)
        if (pr.vec[k] != nullptr) rd[nr++] = pr.vec[k];
      rd[nr++] = _d;
      const double* wr = _d;
      B200::statement_queue().flush_for(rd, nr, &wr, 1);
    }
    B200::check(sb_eval(_ctx, _d, _n, assign_op, &e.program()), "sb_eval");
    return *this;
  }
  DeviceVector& operator+=(const DevExpr& e) { return eval(SB_ADD_ASSIGN, e); }
  DeviceVector& operator-=(const DevExpr& e) { return eval(SB_SUB_ASSIGN, e); }
  DeviceVector& operator*=(double s) { return eval(SB_MUL_ASSIGN, DevExpr::scalar(s)); }
  DeviceVector& operator/=(double s) { return eval(SB_DIV_ASSIGN, DevExpr::scalar(s)); }

private:

  void allocate(sb_ctx* ctx, size_t n) {
    if (ctx == nullptr) throw std::runtime_error("stormb200: DeviceVector needs a context");
    _ctx = ctx, _n = n, _owned = true;
    B200::check(sb_vec_alloc(ctx, n, &_d), "sb_vec_alloc");
  }
  void release() noexcept {
    if (_d != nullptr && (!B200::statement_queue().stmts.empty() || B200::statement_queue().apply.active)) {
      try { // queued statements may read or write this storage
        B200::flush_for({_d}, {_d});
      } catch (...) { // release() runs in destructors
      }
    }
    if (_owned && _d != nullptr) sb_vec_free(_ctx, _d);
    _d = nullptr, _n = 0, _owned = false;
  }

  sb_ctx* _ctx = nullptr;
  double* _d = nullptr;
  size_t _n = 0;
  bool _owned = false;
};

inline DevExpr::DevExpr(const DeviceVector& v) : _n_vec{1}, _ctx{v.context()}, _n{v.size()} {
  _e.vec[0] = v.data();
  _e.ops[0] = SB_OP_VEC0;
  _e.n_ops = 1;
}

// ---- y <<= expr (MatrixAlgorithms.hpp:120-124) ---------------------------------------------------
inline DeviceVector& operator<<=(DeviceVector& y, const DevExpr& e) { return y.eval(SB_ASSIGN, e); }
inline DeviceVector& operator<<=(DeviceVector& y, DeviceVector& x) { return y.eval(SB_ASSIGN, DevExpr{x}); }
inline DeviceVector& operator<<=(DeviceVector& y, const DeviceVector& x) { return y.eval(SB_ASSIGN, DevExpr{x}); }

// ---- expression builders (MatrixMath.hpp:233-301), all cv combinations ---------------------------
#define STORM_B200_BINARY(OPNAME, CODE)                                                                          \
  inline DevExpr OPNAME(const DevExpr& a, const DevExpr& b) { return DevExpr::binary(a, b, CODE); }              \
  inline DevExpr OPNAME(DeviceVector& a, DeviceVector& b) { return DevExpr::binary(a, b, CODE); }                \
  inline DevExpr OPNAME(DeviceVector& a, const DeviceVector& b) { return DevExpr::binary(a, b, CODE); }          \
  inline DevExpr OPNAME(const DeviceVector& a, DeviceVector& b) { return DevExpr::binary(a, b, CODE); }          \
  inline DevExpr OPNAME(const DeviceVector& a, const DeviceVector& b) { return DevExpr::binary(a, b, CODE); }    \
  inline DevExpr OPNAME(DeviceVector& a, const DevExpr& b) { return DevExpr::binary(a, b, CODE); }               \
  inline DevExpr OPNAME(const DeviceVector& a, const DevExpr& b) { return DevExpr::binary(a, b, CODE); }         \
  inline DevExpr OPNAME(const DevExpr& a, DeviceVector& b) { return DevExpr::binary(a, b, CODE); }               \
  inline DevExpr OPNAME(const DevExpr& a, const DeviceVector& b) { return DevExpr::binary(a, b, CODE); }
STORM_B200_BINARY(operator+, SB_OP_ADD)
STORM_B200_BINARY(operator-, SB_OP_SUB)
#undef STORM_B200_BINARY

// scal * mat, mat * scal (MatrixMath.hpp:247-257): element-wise scal * x (commutative in IEEE-754).
inline DevExpr operator*(double s, const DevExpr& a) { return DevExpr::binary(DevExpr::scalar(s), a, SB_OP_MUL); }
inline DevExpr operator*(double s, DeviceVector& a) { return s * DevExpr{a}; }
inline DevExpr operator*(double s, const DeviceVector& a) { return s * DevExpr{a}; }
inline DevExpr operator*(const DevExpr& a, double s) { return DevExpr::binary(a, DevExpr::scalar(s), SB_OP_MUL); }
inline DevExpr operator*(DeviceVector& a, double s) { return DevExpr{a} * s; }
inline DevExpr operator*(const DeviceVector& a, double s) { return DevExpr{a} * s; }
// mat / scal (MatrixMath.hpp:266-270)
inline DevExpr operator/(const DevExpr& a, double s) { return DevExpr::binary(a, DevExpr::scalar(s), SB_OP_DIV); }
inline DevExpr operator/(DeviceVector& a, double s) { return DevExpr{a} / s; }
inline DevExpr operator/(const DeviceVector& a, double s) { return DevExpr{a} / s; }
// unary minus (MatrixMath.hpp:241-243)
inline DevExpr operator-(const DevExpr& a) { return a.negated(); }
inline DevExpr operator-(DeviceVector& a) { return DevExpr{a}.negated(); }
inline DevExpr operator-(const DeviceVector& a) { return DevExpr{a}.negated(); }

// ---- map(func, vec) (MatrixMath.hpp:101-105) ------------------------------------------------------
// The reference applies `func` to every element on the host. A C ABI cannot carry a C++ callable, so the
// function is TRACED instead: it is called once with a symbolic element (B200::Sym), whose arithmetic
// operators record the postfix program that sb_eval then runs per element on the device, in the association
// order the function's body was written in. The callable must therefore be generic in its argument
// (`[](auto c) { return 2.0 * c * (c - 1.0) * (2.0 * c - 1.0); }` -- the playground's dF_dc,
// Playground.cpp:142-144, with `real_t c` spelled `auto c`) and use + - * / only.  `f <<= B200::map(dF_dc, c);`
namespace B200 {
struct Sym {
  DevExpr e;
};
inline Sym operator+(const Sym& a, const Sym& b) { return {DevExpr::binary(a.e, b.e, SB_OP_ADD)}; }
inline Sym operator-(const Sym& a, const Sym& b) { return {DevExpr::binary(a.e, b.e, SB_OP_SUB)}; }
inline Sym operator*(const Sym& a, const Sym& b) { return {DevExpr::binary(a.e, b.e, SB_OP_MUL)}; }
inline Sym operator/(const Sym& a, const Sym& b) { return {DevExpr::binary(a.e, b.e, SB_OP_DIV)}; }
inline Sym operator+(const Sym& a, double s) { return {DevExpr::binary(a.e, DevExpr::scalar(s), SB_OP_ADD)}; }
inline Sym operator-(const Sym& a, double s) { return {DevExpr::binary(a.e, DevExpr::scalar(s), SB_OP_SUB)}; }
inline Sym operator*(const Sym& a, double s) { return {DevExpr::binary(a.e, DevExpr::scalar(s), SB_OP_MUL)}; }
inline Sym operator/(const Sym& a, double s) { return {DevExpr::binary(a.e, DevExpr::scalar(s), SB_OP_DIV)}; }
inline Sym operator+(double s, const Sym& a) { return {DevExpr::binary(DevExpr::scalar(s), a.e, SB_OP_ADD)}; }
inline Sym operator-(double s, const Sym& a) { return {DevExpr::binary(DevExpr::scalar(s), a.e, SB_OP_SUB)}; }
inline Sym operator*(double s, const Sym& a) { return {DevExpr::binary(DevExpr::scalar(s), a.e, SB_OP_MUL)}; }
inline Sym operator/(double s, const Sym& a) { return {DevExpr::binary(DevExpr::scalar(s), a.e, SB_OP_DIV)}; }
inline Sym operator-(const Sym& a) { return {a.e.negated()}; }
inline Sym operator+(const Sym& a) { return a; }
/// B200::map(func, vec): called QUALIFIED. An unqualified `map(...)` would also consider the reference's generic
/// template (found by ADL), whose constraint instantiates `func` with the host element proxy -- a hard error inside a
/// generic lambda's body, not a substitution failure.
template<class Func>
  requires requires(Func f, Sym s) { { f(s) } -> std::same_as<Sym>; }
inline DevExpr map(Func func, const DeviceVector& v) {
  return func(Sym{DevExpr{v}}).e;
}
} // namespace B200

// ---- reductions (MatrixAlgorithms.hpp:262-270, 310-317): fixed tree SB_TREE v1, result on the host
namespace B200 {
inline double dot_impl(const DeviceVector& a, const DeviceVector& b) {
  if (a.context() != b.context() || a.size() != b.size()) {
    throw std::runtime_error("stormb200: dot_product of vectors of different size or context");
  }
  const double v = statement_queue().reduce(a.context(), a.data(), b.data(), a.size());
  observe(v);
  return v;
}
inline double norm_impl(const DeviceVector& a) {
  double v = 0.0;
  if (statement_queue().enabled && (!statement_queue().stmts.empty() || statement_queue().apply.active)) {
    // norm_2 = sqrt(sum |a_i|^2) (MatrixAlgorithms.hpp:262-270), the sum riding on the queued statements
    v = std::sqrt(statement_queue().reduce(a.context(), a.data(), a.data(), a.size()));
  } else {
    check(sb_norm2(a.context(), a.data(), a.size(), &v), "sb_norm2");
  }
  observe(v);
  return v;
}
} // namespace B200
inline double dot_product(DeviceVector& a, DeviceVector& b) { return B200::dot_impl(a, b); }
inline double dot_product(DeviceVector& a, const DeviceVector& b) { return B200::dot_impl(a, b); }
inline double dot_product(const DeviceVector& a, DeviceVector& b) { return B200::dot_impl(a, b); }
inline double dot_product(const DeviceVector& a, const DeviceVector& b) { return B200::dot_impl(a, b); }
inline double norm_2(DeviceVector& a) { return B200::norm_impl(a); }
inline double norm_2(const DeviceVector& a) { return B200::norm_impl(a); }

// ---- fills ---------------------------------------------------------------------------------------
inline DeviceVector& fill_with(DeviceVector& y, double s) {
  B200::flush_for({}, {y.data()});
  B200::check(sb_fill(y.context(), y.data(), y.size(), s), "sb_fill");
  return y;
}
inline DeviceVector& fill(DeviceVector& y, double s) { return fill_with(y, s); }
/// Uniform(0,1) numbers from the reference's engine construction, generated on the host in row
/// order and uploaded: integer RNG state must be reproduced exactly.
inline DeviceVector& fill_randomly(DeviceVector& y) {
  std::uniform_real_distribution<double> distribution{0.0, 1.0};
  std::vector<double> h(y.size());
  for (double& v : h) v = distribution(B200::random_engine());
  if (!h.empty()) y.upload(h.data());
  return y;
}

} // namespace Storm

#include <Storm/Solvers/Operator.hpp>

namespace Storm {

static_assert(matrix<DeviceVector>);
static_assert(legacy_vector_like<DeviceVector>);
static_assert(!output_matrix<DeviceVector>, "the generic per-element Bittern path must stay unreachable");

/// The uploaded matrix-free FVM operator as an Operator<DeviceVector> (Operator.hpp:66-74): what the
/// playground wraps in a FunctionalOperator around stormDivGrad (Playground.cpp:115-131,153-167).
class FvmOperator final : public Operator<DeviceVector> {
public:

  /// Takes ownership of nothing: `op` must outlive this object (sb_op_create / sb_op_destroy).
  FvmOperator(sb_ctx* ctx, const sb_op* op) : _ctx{ctx}, _op{op} {}

  /// Upload a face-list mesh (sb_mesh_soa) and own the resulting device operator.
  FvmOperator(sb_ctx* ctx, const sb_mesh_soa& mesh, const sb_op_desc& desc) : _ctx{ctx} {
    B200::check(sb_op_create(ctx, &mesh, &desc, &_owned), "sb_op_create");
    _op = _owned;
  }
  ~FvmOperator() override {
    if (_owned != nullptr) sb_op_destroy(_ctx, _owned);
  }

  void mul(DeviceVector& y, const DeviceVector& x) const override {
    if (B200::statement_queue().enabled) { // launched with the next statement, or together with the dot that follows
      B200::statement_queue().defer_apply(_ctx, _op, x.data(), y.data(), x.size());
    } else {
      B200::flush();
      B200::check(sb_apply(_ctx, _op, x.data(), y.data()), "sb_apply");
    }
    ++_num_applies;
  }

  sb_ctx* context() const noexcept { return _ctx; }
  const sb_op* handle() const noexcept { return _op; }
  size_t num_applies() const noexcept { return _num_applies; }

private:

  sb_ctx* _ctx = nullptr;
  const sb_op* _op = nullptr;
  sb_op* _owned = nullptr;
  mutable size_t _num_applies = 0;
};

namespace B200 {
/// `stormDivGrad(mesh, u, dt, c)` (Playground.cpp:115-131) on the device: u += dt * div grad c over the faces of the
/// uploaded mesh, each cell adding its face terms to the value the caller left in u, in ascending face order
/// (bit-identical to the face loop). `op` must hold a faithful-form operator (sb_op_desc::form = SB_FORM_FAITHFUL).
inline void div_grad(const FvmOperator& op, DeviceVector& u, double dt, const DeviceVector& c) {
  if (u.context() != c.context() || u.size() != c.size()) {
    throw std::runtime_error("stormb200: div_grad on vectors of different size or context");
  }
  flush_for({c.data(), u.data()}, {u.data()});
  check(sb_apply_accumulate(op.context(), op.handle(), dt, c.data(), u.data()), "sb_apply_accumulate");
}
} // namespace B200

} // namespace Storm

#include <Storm/Solvers/Preconditioner.hpp>

namespace Storm {

/// Jacobi (point-diagonal) preconditioner for the device path: fills the slot every reference solver
/// already branches on (`IterativeSolver::pre_op` / `pre_side`, Solver.hpp:74-75). y = D^-1 x with D the
/// diagonal of the coefficient-form FVM operator (sb_op_jacobi). The reference ships only
/// IdentityPreconditioner (Preconditioner.hpp:84-97), so this is new functionality behind its interface:
///     solver.pre_op = std::make_unique<Storm::JacobiPreconditioner>(fvm_op);
class JacobiPreconditioner final : public Preconditioner<DeviceVector> {
public:

  explicit JacobiPreconditioner(const FvmOperator& op) : _ctx{op.context()}, _op{op.handle()} {}
  JacobiPreconditioner(sb_ctx* ctx, const sb_op* op) : _ctx{ctx}, _op{op} {}

  void mul(DeviceVector& y, const DeviceVector& x) const override {
    B200::flush_for({x.data()}, {y.data()});
    B200::check(sb_op_jacobi(_ctx, _op, x.data(), y.data()), "sb_op_jacobi");
  }
  void conj_mul(DeviceVector& x, const DeviceVector& y) const override { mul(x, y); } // D is real

private:

  sb_ctx* _ctx = nullptr;
  const sb_op* _op = nullptr;
};

} // namespace Storm
