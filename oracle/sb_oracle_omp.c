/*
 * sb_oracle_omp.c -- TEST / MEASUREMENT INFRASTRUCTURE, not product code and not a parity oracle.
 *
 * A multi-threaded (OpenMP) port of the two benchmark solvers on the host, for bench.py's
 * `cpu_baseline_all_cores` figure only: the reference itself is single-threaded by construction,
 * so the CPU number next to the GPU line that uses every host core has to come from a port. Same algorithms and
 * statement sequence as SolverCg.hpp:54-126 / SolverBiCgStab.hpp:59-165 (no preconditioner) on the operator's
 * coefficient rows (oracle row contract, orc_rows_coef), with the passes merged the way the GPU path merges them
 * (each dot computed in the pass that produces its operand). Reductions are OpenMP reductions: the summation order
 * differs from the reference's, so the iterates agree with it to rounding, not bit for bit -- bench.py checks the
 * residual against the single-threaded run before quoting the time.
 * Compile: gcc -O2 -fopenmp -ffp-contract=off.
 */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define COL_PAD INT32_MIN

static double safe_divide(double x, double y) { return y == 0.0 ? 0.0 : x / y; }

/* y = A x (coefficient rows, column-major ELL [width][ld]); returns sum_i u_i * y_i (u may be x, y (pass NULL -> y) or a third vector) */
static double apply_dot(int64_t n, int width, int64_t ld, const int32_t* col, const double* a, const double* diag,
                        const double* x, double* y, const double* u, double* second /* optional: sum y_i * x_i */) {
  double s = 0.0, s2 = 0.0;
#pragma omp parallel for reduction(+ : s, s2) schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    double acc = diag[i] * x[i];
    for (int k = 0; k < width; ++k) {
      const int32_t c = col[(int64_t) k * ld + i];
      if (c >= 0) acc += a[(int64_t) k * ld + i] * x[c];
    }
    y[i] = acc;
    s += (u != NULL ? u[i] : acc) * acc;
    s2 += acc * x[i];
  }
  if (second != NULL) *second = s2;
  return s;
}

int orc_omp_threads(void) { return omp_get_max_threads(); }

/* solver: 0 = CG, 1 = BiCGStab. Runs exactly `iters` iterations from x (in/out); returns the residual norm after
 * the last one. *seconds = wall time of the iteration loop only (initial residual excluded). */
double orc_omp_solve(int solver, int64_t n, int width, int64_t ld, const int32_t* col, const double* a,
                     const double* diag, const double* b, double* x, int64_t iters, double* seconds) {
  double* r = (double*) malloc(sizeof(double) * (size_t) n);
  double* p = (double*) malloc(sizeof(double) * (size_t) n);
  double* v = (double*) malloc(sizeof(double) * (size_t) n);
  double* t = (double*) malloc(sizeof(double) * (size_t) n);
  double* rt = (double*) malloc(sizeof(double) * (size_t) n);
  apply_dot(n, width, ld, col, a, diag, x, r, NULL, NULL);
  double rho = 0.0;
#pragma omp parallel for reduction(+ : rho) schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    r[i] = b[i] - r[i];
    p[i] = r[i], rt[i] = r[i];
    rho += r[i] * r[i];
  }
  double err = sqrt(rho), alpha = 0.0, omega = 0.0, gamma = rho;
  const double t0 = omp_get_wtime();
  for (int64_t it = 0; it < iters; ++it) {
    if (solver == 0) {
      const double pz = apply_dot(n, width, ld, col, a, diag, p, v, p, NULL); /* z = A p, <p,z> */
      alpha = safe_divide(gamma, pz);
      double g2 = 0.0;
#pragma omp parallel for reduction(+ : g2) schedule(static)
      for (int64_t i = 0; i < n; ++i) {
        x[i] += alpha * p[i];
        r[i] -= alpha * v[i];
        g2 += r[i] * r[i];
      }
      const double beta = safe_divide(g2, gamma);
      gamma = g2, err = sqrt(g2);
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < n; ++i) p[i] = r[i] + beta * p[i];
    } else {
      if (it != 0) {
        double rho2 = 0.0;
#pragma omp parallel for reduction(+ : rho2) schedule(static)
        for (int64_t i = 0; i < n; ++i) rho2 += rt[i] * r[i];
        const double beta = safe_divide(alpha * rho2, omega * rho);
        rho = rho2;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) p[i] = r[i] + beta * (p[i] - omega * v[i]);
      }
      const double rv = apply_dot(n, width, ld, col, a, diag, p, v, rt, NULL); /* v = A p, <r~,v> */
      alpha = safe_divide(rho, rv);
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < n; ++i) r[i] -= alpha * v[i];
      double tr = 0.0;
      const double tt = apply_dot(n, width, ld, col, a, diag, r, t, NULL, &tr); /* t = A r, <t,t>, <t,r> */
      omega = safe_divide(tr, tt);
      double rr = 0.0;
#pragma omp parallel for reduction(+ : rr) schedule(static)
      for (int64_t i = 0; i < n; ++i) {
        x[i] = (x[i] + alpha * p[i]) + omega * r[i];
        r[i] -= omega * t[i];
        rr += r[i] * r[i];
      }
      err = sqrt(rr);
    }
  }
  *seconds = omp_get_wtime() - t0;
  free(r), free(p), free(v), free(t), free(rt);
  return err;
}
