/// Chebyshev polynomial preconditioner in the reference's `Preconditioner<Vector>` slot
/// (source/Storm/Solvers/Preconditioner.hpp:63-77). The reference ships only the identity preconditioner and lists
/// polynomial ones as planned (README); every solver already branches on `pre_op` / `pre_side`
/// (e.g. SolverCg.hpp:64-84, SolverBiCgStab.hpp:135-137, SolverGmres.hpp:149-156,233-248), so this is the interface a
/// maintainer would fill in.
///
/// Written ONLY in the reference's own vector vocabulary -- `assign`, `<<=`, `+=`, `scalar * v`, `a - b`,
/// `a + b`, `norm_2`, `Operator::mul` -- so the same template runs on the reference's host vectors (the oracle compiles
/// it on its host vector, oracle/ref_build/ref_solvers.cpp) and on `Storm::DeviceVector`, where every statement is an
/// element-wise kernel or joins a statement group and every `mul` is the matrix-free operator apply. It needs no
/// matrix entries, no triangular solves and no reductions in its application: one application of degree k costs k - 1
/// operator applies and ~5 (k - 1) vector passes, all on the HBM roofline, and no synchronisation -- which is what a
/// polynomial preconditioner buys on a GPU (and across GPUs: k - 1 halo exchanges, zero all-reduces).
///
/// z = p(A) r with p the degree-(k-1) Chebyshev polynomial that minimises max |1 - lambda p(lambda)| over
/// [lambda_min, lambda_max]: the classical three-term recurrence of the Chebyshev iteration for A z = r started from
/// z = 0 (Saad, Iterative Methods for Sparse Linear Systems, 2nd ed., Alg. 12.1). p(A) is symmetric positive definite
/// for symmetric positive definite A, so it is admissible for CG. `build()` estimates lambda_max by a few power
/// iterations (x <- A x / ||A x||, started from the right-hand side) times a safety factor and sets
/// lambda_min = lambda_max / eig_ratio: the polynomial damps the upper part of the spectrum, the Krylov method takes
/// care of the rest.
#pragma once

#include <Storm/Solvers/Preconditioner.hpp>

#include <cstddef>

namespace Storm {

template<legacy_vector_like Vector>
class ChebyshevPreconditioner final : public Preconditioner<Vector> {
public:

  size_t degree{4};               ///< terms of the polynomial: degree - 1 operator applies per application
  real_t eig_ratio{30.0};         ///< lambda_min = lambda_max / eig_ratio
  size_t num_power_iterations{10};
  real_t safety_factor{1.1};      ///< lambda_max = safety_factor * (power-iteration estimate)
  real_t lambda_max{0.0};         ///< > 0 before build(): used as given (no power iterations)
  real_t lambda_min{0.0};

  void build(const Vector& x_vec, const Vector& b_vec, const Operator<Vector>& any_op) override {
    _op = &any_op;
    _d_vec.assign(x_vec, false);
    _r_vec.assign(x_vec, false);
    if (!(lambda_max > 0.0)) {
      // power iteration: d <- A d / ||A d||
      _d_vec <<= b_vec;
      real_t nrm = norm_2(_d_vec);
      if (nrm == 0.0) {
        fill_with(_d_vec, 1.0);
        nrm = norm_2(_d_vec);
      }
      real_t estimate = 0.0;
      for (size_t k = 0; k < num_power_iterations; ++k) {
        _d_vec <<= (1.0 / nrm) * _d_vec;
        _op->mul(_r_vec, _d_vec);
        estimate = norm_2(_r_vec);
        if (estimate == 0.0) break;
        std::swap(_d_vec, _r_vec);
        nrm = estimate;
      }
      lambda_max = safety_factor * estimate;
    }
    lambda_min = lambda_max / eig_ratio;
  }

  void mul(Vector& z_vec, const Vector& r_vec) const override {
    const real_t theta = 0.5 * (lambda_max + lambda_min), delta = 0.5 * (lambda_max - lambda_min);
    const real_t sigma1 = theta / delta;
    real_t rho = 1.0 / sigma1;
    // d <- r / theta, z <- d
    _d_vec <<= (1.0 / theta) * r_vec;
    z_vec <<= _d_vec;
    for (size_t k = 1; k < degree; ++k) {
      // residual of the inner system: w <- r - A z
      _op->mul(_r_vec, z_vec);
      _r_vec <<= r_vec - _r_vec;
      const real_t rho_new = 1.0 / (2.0 * sigma1 - rho);
      // d <- rho_new rho d + (2 rho_new / delta) w ; z <- z + d
      _d_vec <<= (rho_new * rho) * _d_vec + (2.0 * rho_new / delta) * _r_vec;
      z_vec += _d_vec;
      rho = rho_new;
    }
  }

  void conj_mul(Vector& x_vec, const Vector& y_vec) const override { mul(x_vec, y_vec); } // p(A) is self-adjoint with A

private:

  const Operator<Vector>* _op = nullptr;
  mutable Vector _d_vec, _r_vec;

}; // class ChebyshevPreconditioner

} // namespace Storm
