// sb_solver_bodies.cuh -- the device-side pieces of the fused CG / BiCGStab solvers: the solver state, the scalar
// updates that follow each reduction ("Final" functors) and the element-wise update bodies. Shared by the two
// schedules that run them: one kernel per step (sb_solvers.cu) and the persistent whole-solve kernel (sb_mega.cuh).
//
// Reference: CgSolver (source/Storm/Solvers/SolverCg.hpp:54-126), BiCgStabSolver (SolverBiCgStab.hpp:59-165), driven
// as IterativeSolver::solve (Solver.hpp:116-147), no preconditioner. Every statement keeps the reference's
// per-element operation order.
#pragma once

#include "sb_op.cuh"

namespace sb {

// (Recorder and the scalar updates behind the reductions live in sb_finals.cuh, where the in-kernel reducer needs them.)

// ---- CG -------------------------------------------------------------------------------------------
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
  // staged inputs (persistent kernel: the runs of the NV input vectors arrive by bulk copy)
  static constexpr int NV = 4;
  __device__ __forceinline__ const double* in(int k) const { return k == 0 ? x : (k == 1 ? p : (k == 2 ? r : z)); }
  __device__ __forceinline__ void fill(Regs& g, const double2 (&v)[NV]) const { g.x = v[0], g.p = v[1], g.r = v[2], g.z = v[3]; }
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
  static constexpr int NV = 2;
  __device__ __forceinline__ const double* in(int k) const { return k == 0 ? p : r; }
  __device__ __forceinline__ void fill(Regs& g, const double2 (&v)[NV]) const { g.p = v[0], g.r = v[1]; }
  __device__ __forceinline__ void run(int64_t e0, int64_t, Regs& g, double (&)[1]) const {
    const double beta = st->beta;
    double2 pn;
    pn.x = __dadd_rn(g.r.x, __dmul_rn(beta, g.p.x));
    pn.y = __dadd_rn(g.r.y, __dmul_rn(beta, g.p.y));
    st2(p, e0, pn);
  }
};

// ---- BiCGStab -------------------------------------------------------------------------------------
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
  static constexpr int NV = 3; // the staged variant fetches p and v in iteration 0 too (unused there)
  __device__ __forceinline__ const double* in(int k) const { return k == 0 ? r : (k == 1 ? p : v); }
  __device__ __forceinline__ void fill(Regs& g, const double2 (&w)[NV]) const { g.r = w[0], g.p = w[1], g.v = w[2]; }
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
  static constexpr int NV = 2;
  __device__ __forceinline__ const double* in(int k) const { return k == 0 ? r : v; }
  __device__ __forceinline__ void fill(Regs& g, const double2 (&w)[NV]) const { g.r = w[0], g.v = w[1]; }
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
  static constexpr int NV = 5;
  __device__ __forceinline__ const double* in(int k) const { return k == 0 ? x : (k == 1 ? p : (k == 2 ? r : (k == 3 ? t : rt))); }
  __device__ __forceinline__ void fill(Regs& g, const double2 (&v)[NV]) const {
    g.x = v[0], g.p = v[1], g.r = v[2], g.t = v[3], g.rt = v[4];
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

enum class Kind { Cg, BiCgStab };

// The persistent whole-solve schedule (sb_mega.cu): the iteration loop of an initialised solve in ONE cooperative
// launch. The vectors are the solver's workspaces, `st` the device solver state the initialisation kernels have filled.
struct MegaLaunch {
  Kind kind;
  double *x, *r, *p, *v, *t, *rt; // CG: v = z; t, rt unused
  SolveBlock* blk;        // version 0 holds the initialised state; the kernel publishes final_ / done
  double* hist;
  double* trace;
  int32_t timeline_iters; // > 0: record the in-kernel timeline of the first iterations into ctx->d_timeline
};
bool mega_supported(const sb_ctx* ctx, const sb_op* op);
int launch_mega(sb_ctx* ctx, const sb_op* op, const MegaLaunch& L);

} // namespace sb
