// Storm/B200/GroupedSolvers.hpp -- IDR(s) and BiCGStab(l) with their vector statements issued in groups.
//
// The reference's IdrsSolver (Solvers/SolverIdrs.hpp:42-292) and BiCgStabLSolver (SolverBiCgStab.hpp:185-383) spend
// most of their vector passes in runs of consecutive linear-combination statements (`v <<= r - gamma_k*g_k;
// v -= gamma_i*g_i; ...`, `u_i <<= r_i - beta*u_i` for every i, the 3(l-1)+3 updates that close a BiCGStab(l) cycle)
// and in runs of dot products whose results are only needed together. On the generic drop-in path every one of them
// is a kernel of its own and every reduction a host round trip. The classes below are the same algorithms,
// statement for statement, behind the same base class (InnerOuterIterativeSolver<DeviceVector>: the reference's own
// driver loop, stopping rule and public fields are used unchanged), with each run handed to the device as ONE
// sb_eval_group launch: one pass over the union of the operands, one synchronisation for all the reductions behind it.
// Per-element arithmetic is untouched -- a chain `((base - c0*x0) - c1*x1) ...` rounds exactly like the statements it
// stands for -- so the iterates, the residual history and the sequence of reduction values are bit-identical to the
// reference templates on the same vector type (tests/test_dropin_emulated.py on the host emulator of the C ABI,
// tests/test_gpu_playground.py on the device). Scalars stay on the host, like the reference's.
//
// Vector passes per iteration (V = one read or write of one vector; oracle/statement_trace.py counts them):
//   IDR(4)       43.25 V in 18.25 launches as written  ->  31.0 V in 6.5 launches
//   BiCGStab(2)  31.5  V in 15    launches as written  ->  25.5 V in 8.5 launches
//
// Any Operator<DeviceVector> works (the groups only touch vectors). A preconditioner is not supported: use the
// reference templates for that.
#pragma once

#include <Storm/B200/DeviceVector.hpp>

#include <Storm/Solvers/Solver.hpp>

#include <cmath>
#include <utility>
#include <vector>

namespace Storm::B200 {

/// Builder of one statement group (sb_eval_group). Statements are appended with chain()/plus()/minus(), reductions
/// with dot()/norm(); run() launches and returns the reduction values in the order they were appended, reporting each
/// to the reduction observer exactly as dot_product / norm_2 would. Chains longer than the ABI's term limit continue as
/// `y = y +- ...` (the same arithmetic); more statements than fit in one launch are issued as several launches with
/// the reductions behind the last one.
class Group {
public:

  Group(sb_ctx* ctx, size_t n) : _ctx{ctx}, _n{n} {}

  /// Start the statement `y = base ...` (base == nullptr: the chain starts from its first term).
  Group& chain(DeviceVector& y, const DeviceVector* base) {
    sb_chain c{};
    c.y = checked(y), c.base = base != nullptr ? checked(*base) : nullptr, c.n_terms = 0;
    _stmts.push_back(c);
    return *this;
  }
  Group& plus(double c, const DeviceVector& x) { return term(c, x, 0); }
  Group& minus(double c, const DeviceVector& x) { return term(c, x, 1); }
  Group& dot(const DeviceVector& a, const DeviceVector& b) {
    _da.push_back(checked(a)), _db.push_back(checked(b)), _is_norm.push_back(false);
    return *this;
  }
  Group& norm(const DeviceVector& a) {
    _da.push_back(checked(a)), _db.push_back(checked(a)), _is_norm.push_back(true);
    return *this;
  }

  std::vector<double> run() {
    for (const sb_chain& c : _stmts) {
      if (c.n_terms == 0) throw std::runtime_error("stormb200: statement group holds a chain without terms");
    }
    std::vector<double> out(_da.size(), 0.0);
    if (_stmts.empty() && _da.empty()) return out;
    size_t s0 = 0, d0 = 0;
    // Under statement grouping an operator apply may still be pending (FvmOperator::mul defers it): a leading dot with
    // its output rides on it (sb_apply_dot), and <y,x> right behind <y,y> comes from the same kernel.
    StatementQueue& q = statement_queue();
    if (q.enabled && q.apply.active && _stmts.empty()) {
      out[0] = q.reduce(_ctx, _da[0], _db[0], _n);
      d0 = 1;
      if (_da.size() > 1 && q.spare.valid) out[1] = q.reduce(_ctx, _da[1], _db[1], _n), d0 = 2;
    }
    flush(); // statements queued by the automatic grouping of DeviceVector come first
    if (d0 == _da.size() && _stmts.empty()) { // everything rode on the apply
      finish(out);
      return out;
    }
    // statements in launches of at most SB_GROUP_MAX_STMT; reductions ride on the last statement launch, at most
    // SB_GROUP_MAX_DOTS per call
    do {
      const size_t ns = std::min<size_t>(_stmts.size() - s0, SB_GROUP_MAX_STMT);
      const bool last = s0 + ns == _stmts.size();
      const size_t nd = last ? std::min<size_t>(_da.size() - d0, SB_GROUP_MAX_DOTS) : 0;
      check(sb_eval_group(_ctx, _n, (int) ns, ns > 0 ? _stmts.data() + s0 : nullptr, (int) nd,
                          nd > 0 ? _da.data() + d0 : nullptr, nd > 0 ? _db.data() + d0 : nullptr,
                          nd > 0 ? out.data() + d0 : nullptr),
            "sb_eval_group");
      s0 += ns, d0 += nd;
    } while (s0 < _stmts.size() || d0 < _da.size());
    finish(out);
    return out;
  }

private:

  void finish(std::vector<double>& out) {
    for (size_t d = 0; d < out.size(); ++d) {
      if (_is_norm[d]) out[d] = std::sqrt(out[d]); // norm_2 = sqrt(sum |a_i|^2), MatrixAlgorithms.hpp:262-270
      observe(out[d]);
    }
    _stmts.clear(), _da.clear(), _db.clear(), _is_norm.clear();
  }

  const double* checked(const DeviceVector& v) const {
    if (v.context() != _ctx || v.size() != _n) {
      throw std::runtime_error("stormb200: statement group mixes vectors of different size or context");
    }
    return v.data();
  }
  double* checked(DeviceVector& v) const { return const_cast<double*>(checked(std::as_const(v))); }
  Group& term(double c, const DeviceVector& x, uint8_t sub) {
    if (_stmts.empty()) throw std::runtime_error("stormb200: term without a statement");
    if (_stmts.back().n_terms == SB_GROUP_MAX_TERMS) { // continue the chain: y = y +- ...
      sb_chain next{};
      next.y = _stmts.back().y, next.base = _stmts.back().y, next.n_terms = 0;
      _stmts.push_back(next);
    }
    sb_chain& ch = _stmts.back();
    if (ch.base == nullptr && ch.n_terms == 0 && sub != 0) {
      throw std::runtime_error("stormb200: a chain without a base cannot start with a subtraction");
    }
    ch.x[ch.n_terms] = checked(x), ch.c[ch.n_terms] = c, ch.sub[ch.n_terms] = sub;
    ch.n_terms++;
    return *this;
  }

  sb_ctx* _ctx;
  size_t _n;
  std::vector<sb_chain> _stmts;
  std::vector<const double*> _da, _db;
  std::vector<bool> _is_norm;
};

namespace detail {
inline void no_preconditioner(const void* pre_op) {
  if (pre_op != nullptr) {
    throw std::runtime_error("stormb200: the grouped solvers do not take a preconditioner; "
                             "use the reference solver templates");
  }
}
} // namespace detail

/// IDR(s) (SolverIdrs.hpp:42-292), statements grouped. Same defaults (s = 4).
class IdrsSolver final : public InnerOuterIterativeSolver<DeviceVector> {
public:

  IdrsSolver() { this->num_inner_iterations = 4; }

private:

  real_t _omega{};
  std::vector<real_t> _phi, _gamma, _mu; // _mu: s x s, row-major
  DeviceVector _r_vec, _v_vec;
  std::vector<DeviceVector> _p_vecs, _u_vecs, _g_vecs;

  real_t& mu(size_t i, size_t j) { return _mu[i * this->num_inner_iterations + j]; }

  real_t outer_init(const DeviceVector& x_vec, const DeviceVector& b_vec, const Operator<DeviceVector>& lin_op,
                    const Preconditioner<DeviceVector>* pre_op) override {
    detail::no_preconditioner(pre_op);
    const size_t s = this->num_inner_iterations;
    _phi.assign(s, 0.0), _gamma.assign(s, 0.0), _mu.assign(s * s, 0.0);
    _r_vec.assign(x_vec, false), _v_vec.assign(x_vec, false);
    _p_vecs.resize(s), _u_vecs.resize(s), _g_vecs.resize(s);
    for (DeviceVector& p_vec : _p_vecs) p_vec.assign(x_vec, false);
    for (DeviceVector& u_vec : _u_vecs) u_vec.assign(x_vec, false);
    for (DeviceVector& g_vec : _g_vecs) g_vec.assign(x_vec, false);
    // r <- b - A x, phi_0 <- ||r||   (:94-99)
    lin_op.mul(_r_vec, x_vec);
    Group grp{x_vec.context(), x_vec.size()};
    grp.chain(_r_vec, &b_vec).minus(1.0, _r_vec).norm(_r_vec); // b - r: 1.0*r is exact
    _phi[0] = grp.run()[0];
    return _phi[0];
  }

  void inner_init(const DeviceVector& x_vec, const DeviceVector&, const Operator<DeviceVector>&,
                  const Preconditioner<DeviceVector>*) override {
    const size_t s = this->num_inner_iterations;
    if (this->iteration == 0) {
      // the shadow space (:131-141): one-time set-up, the reference's statements as they are
      _omega = mu(0, 0) = 1.0;
      _p_vecs[0] <<= _r_vec / _phi[0];
      for (size_t i = 1; i < s; ++i) {
        mu(i, i) = 1.0, _phi[i] = 0.0;
        fill_randomly(_p_vecs[i]);
        for (size_t j = 0; j < i; ++j) {
          mu(i, j) = 0.0;
          _p_vecs[i] -= dot_product(_p_vecs[i], _p_vecs[j]) * _p_vecs[j];
        }
        _p_vecs[i] /= norm_2(_p_vecs[i]);
      }
    } else {
      // phi_i <- <p_i, r> for every i (:143-145): one batch, one synchronisation
      Group grp{x_vec.context(), x_vec.size()};
      for (size_t i = 0; i < s; ++i) grp.dot(_p_vecs[i], _r_vec);
      const std::vector<double> d = grp.run();
      for (size_t i = 0; i < s; ++i) _phi[i] = d[i];
    }
  }

  real_t inner_iterate(DeviceVector& x_vec, const DeviceVector&, const Operator<DeviceVector>& lin_op,
                       const Preconditioner<DeviceVector>*) override {
    const size_t s = this->num_inner_iterations;
    const size_t k = this->inner_iteration;
    Group grp{x_vec.context(), x_vec.size()};

    // gamma_{k:s-1} <- (mu_{k:s-1,k:s-1})^-1 phi_{k:s-1}   (:184-190), host scalars
    for (size_t i = k; i < s; ++i) {
      _gamma[i] = _phi[i];
      for (size_t j = k; j < i; ++j) _gamma[i] -= mu(i, j) * _gamma[j];
      _gamma[i] /= mu(i, i);
    }

    // v <- r - gamma_k g_k - sum_{i>k} gamma_i g_i ;  u_k <- omega v + gamma_k u_k + sum_{i>k} gamma_i u_i   (:195-206)
    grp.chain(_v_vec, &_r_vec);
    for (size_t i = k; i < s; ++i) grp.minus(_gamma[i], _g_vecs[i]);
    grp.chain(_u_vecs[k], nullptr).plus(_omega, _v_vec);
    for (size_t i = k; i < s; ++i) grp.plus(_gamma[i], _u_vecs[i]);
    grp.run();
    lin_op.mul(_g_vecs[k], _u_vecs[k]); // (:210)

    // bi-orthogonalise g_k and u_k against p_0..p_{k-1} (:221-226): alpha_i needs g_k as updated by step i-1, so the
    // dot for step i+1 rides on the update of step i; the new column of mu (:234-236) rides on the last update
    if (k > 0) {
      grp.dot(_p_vecs[0], _g_vecs[k]);
      double pg = grp.run()[0];
      for (size_t i = 0; i < k; ++i) {
        const real_t alpha = safe_divide(pg, mu(i, i));
        grp.chain(_u_vecs[k], &_u_vecs[k]).minus(alpha, _u_vecs[i]);
        grp.chain(_g_vecs[k], &_g_vecs[k]).minus(alpha, _g_vecs[i]);
        if (i + 1 < k) {
          grp.dot(_p_vecs[i + 1], _g_vecs[k]);
          pg = grp.run()[0];
        }
      }
    }
    for (size_t i = k; i < s; ++i) grp.dot(_p_vecs[i], _g_vecs[k]);
    {
      const std::vector<double> d = grp.run();
      for (size_t i = k; i < s; ++i) mu(i, k) = d[i - k];
    }

    // beta <- phi_k / mu_kk ; x += beta u_k ; r -= beta g_k ; phi_{k+1:} -= beta mu_{k+1:,k}   (:244-256)
    const real_t beta = safe_divide(_phi[k], mu(k, k));
    grp.chain(x_vec, &x_vec).plus(beta, _u_vecs[k]);
    grp.chain(_r_vec, &_r_vec).minus(beta, _g_vecs[k]);
    for (size_t i = k + 1; i < s; ++i) _phi[i] -= beta * mu(i, k);

    if (k == s - 1) {
      // enter the next G subspace (:272-279): v <- A r ; omega <- <v,r>/<v,v> ; x += omega r ; r -= omega v
      grp.run();
      lin_op.mul(_v_vec, _r_vec);
      grp.dot(_v_vec, _v_vec).dot(_v_vec, _r_vec); // g++ evaluates safe_divide's arguments right to left: <v,v> first
      const std::vector<double> d = grp.run();
      _omega = safe_divide(d[1], d[0]);
      grp.chain(x_vec, &x_vec).plus(_omega, _r_vec);
      grp.chain(_r_vec, &_r_vec).minus(_omega, _v_vec);
    }
    grp.norm(_r_vec); // (:282) rides on the last update
    return grp.run()[0];
  }
};

/// BiCGStab(l) (SolverBiCgStab.hpp:185-383), statements grouped. Same defaults (l = 2).
class BiCgStabLSolver final : public InnerOuterIterativeSolver<DeviceVector> {
public:

  BiCgStabLSolver() { this->num_inner_iterations = 2; }

private:

  real_t _alpha{}, _rho{}, _omega{};
  std::vector<real_t> _gamma, _gamma_bar, _gamma_bbar, _sigma, _tau; // _tau: (l+1) x (l+1), row-major
  DeviceVector _r_tilde_vec;
  std::vector<DeviceVector> _r_vecs, _u_vecs;

  real_t& tau(size_t i, size_t j) { return _tau[i * (this->num_inner_iterations + 1) + j]; }

  real_t outer_init(const DeviceVector& x_vec, const DeviceVector& b_vec, const Operator<DeviceVector>& lin_op,
                    const Preconditioner<DeviceVector>* pre_op) override {
    detail::no_preconditioner(pre_op);
    const size_t l = this->num_inner_iterations;
    _gamma.assign(l + 1, 0.0), _gamma_bar.assign(l + 1, 0.0), _gamma_bbar.assign(l + 1, 0.0);
    _sigma.assign(l + 1, 0.0), _tau.assign((l + 1) * (l + 1), 0.0);
    _r_tilde_vec.assign(x_vec, false);
    _r_vecs.resize(l + 1), _u_vecs.resize(l + 1);
    for (DeviceVector& r_vec : _r_vecs) r_vec.assign(x_vec, false);
    for (DeviceVector& u_vec : _u_vecs) u_vec.assign(x_vec, false);
    // u_0 <- 0 ; r_0 <- b - A x ; r~ <- r_0 ; rho <- <r~, r_0>   (:224-231)
    fill_with(_u_vecs[0], 0.0);
    lin_op.Residual(_r_vecs[0], b_vec, x_vec);
    _r_tilde_vec <<= _r_vecs[0];
    _rho = dot_product(_r_tilde_vec, _r_vecs[0]);
    return std::sqrt(_rho);
  }

  real_t inner_iterate(DeviceVector& x_vec, const DeviceVector&, const Operator<DeviceVector>& lin_op,
                       const Preconditioner<DeviceVector>*) override {
    const size_t l = this->num_inner_iterations;
    const size_t j = this->inner_iteration;
    Group grp{x_vec.context(), x_vec.size()};

    // BiCG part (:264-283)
    if (this->iteration == 0) {
      _u_vecs[0] <<= _r_vecs[0];
    } else {
      grp.dot(_r_tilde_vec, _r_vecs[j]);
      const real_t rho_bar = std::exchange(_rho, grp.run()[0]);
      const real_t beta = safe_divide(_alpha * _rho, rho_bar);
      for (size_t i = 0; i <= j; ++i) grp.chain(_u_vecs[i], &_r_vecs[i]).minus(beta, _u_vecs[i]);
      grp.run();
    }
    lin_op.mul(_u_vecs[j + 1], _u_vecs[j]);
    grp.dot(_r_tilde_vec, _u_vecs[j + 1]);
    _alpha = safe_divide(_rho, grp.run()[0]);
    for (size_t i = 0; i <= j; ++i) grp.chain(_r_vecs[i], &_r_vecs[i]).minus(_alpha, _u_vecs[i + 1]);
    // x += alpha u_0 ; r_{j+1} <- A r_j   (:294-299)
    grp.chain(x_vec, &x_vec).plus(_alpha, _u_vecs[0]);
    grp.run();
    lin_op.mul(_r_vecs[j + 1], _r_vecs[j]);

    if (j == l - 1) {
      // minimal-residual part (:312-323): modified Gram-Schmidt over r_1..r_l; tau_ij needs r_j as updated by the
      // previous i, so each dot rides on the update before it; sigma_j and <r_0, r_j> ride on the last update of r_j
      for (size_t jj = 1; jj <= l; ++jj) {
        if (jj > 1) {
          grp.dot(_r_vecs[1], _r_vecs[jj]);
          double rr = grp.run()[0];
          for (size_t i = 1; i < jj; ++i) {
            tau(i, jj) = safe_divide(rr, _sigma[i]);
            grp.chain(_r_vecs[jj], &_r_vecs[jj]).minus(tau(i, jj), _r_vecs[i]);
            if (i + 1 < jj) {
              grp.dot(_r_vecs[i + 1], _r_vecs[jj]);
              rr = grp.run()[0];
            }
          }
        }
        grp.dot(_r_vecs[jj], _r_vecs[jj]).dot(_r_vecs[0], _r_vecs[jj]);
        const std::vector<double> d = grp.run();
        _sigma[jj] = d[0];
        _gamma_bar[jj] = safe_divide(d[1], _sigma[jj]);
      }
      // (:339-351) host scalars
      _omega = _gamma[l] = _gamma_bar[l], _rho *= -_omega;
      for (size_t q = l - 1; q != 0; --q) {
        _gamma[q] = _gamma_bar[q];
        for (size_t i = q + 1; i <= l; ++i) _gamma[q] -= tau(q, i) * _gamma[i];
      }
      for (size_t q = 1; q < l; ++q) {
        _gamma_bbar[q] = _gamma[q + 1];
        for (size_t i = q + 1; i < l; ++i) _gamma_bbar[q] += tau(q, i) * _gamma[i + 1];
      }
      // (:364-371) x, r_0 and u_0 each collect their 1 + (l-1) updates in one chain; the chain of x comes first: its
      // first term reads r_0 before r_0's own chain overwrites it, exactly as in the reference's statement order
      grp.chain(x_vec, &x_vec).plus(_gamma[1], _r_vecs[0]);
      for (size_t q = 1; q < l; ++q) grp.plus(_gamma_bbar[q], _r_vecs[q]);
      grp.chain(_r_vecs[0], &_r_vecs[0]).minus(_gamma_bar[l], _r_vecs[l]);
      for (size_t q = 1; q < l; ++q) grp.minus(_gamma_bar[q], _r_vecs[q]);
      grp.chain(_u_vecs[0], &_u_vecs[0]).minus(_gamma[l], _u_vecs[l]);
      for (size_t q = 1; q < l; ++q) grp.minus(_gamma[q], _u_vecs[q]);
    }
    grp.norm(_r_vecs[0]); // (:374)
    return grp.run()[0];
  }
};

} // namespace Storm::B200
