// sb_op.cuh -- device layout of the FVM operator and the operator-apply kernels.
//
// Reference semantics: the face loop of source_apps/playground/Playground.cpp:115-131 (interior
// faces) plus the boundary-ghost pattern of Feathers/ConvectionScheme.hpp:95-106, restated as a
// cell-row gather so no atomics are needed and every cell sums its face contributions in ascending
// face index, the order the CPU loop produces.
//
// HBM layout (column-major ELL, leading dimension ld = n padded to the 2048-row CTA tile):
//   col  [W][ld] int32   neighbour cell | ~i (Dirichlet mirror ghost of cell i) | kColPad
//   v0   [W][ld] fp64    coef form: a = ((area/vol_i)*dt)/dist ; faithful form: g = area/vol_i
//   v1   [W][ld] fp64    faithful form only: dist
//   diag [ld]    fp64    coef form only: prefill - sum(a) - sum_ghost(a+a)
// A lane owns two consecutive rows, so entry k of both rows is one int2 + one double2 load and a warp
// reads 256 B + 512 B contiguous per k: fully coalesced 64/128-bit streams. Algorithmic bytes per
// apply (coef form): 24*N + 12*entries = 72 B/cell on tets.
#pragma once

#include "sb_kernels.cuh"

namespace sb {

struct OpDev {
  int64_t n = 0, ld = 0;
  int32_t width = 0, form = 0, prefill = 0;
  double dt = 0.0;
  const int32_t* col = nullptr;
  const double* v0 = nullptr;
  const double* v1 = nullptr;
  const double* diag = nullptr;
};

// Epilogues fuse dot products into the apply: they see the input pair x[e0..e0+1] and the freshly
// computed output pair, so <x,Ax>-style reductions cost no extra vector pass.
struct NoEpi {
  struct Regs {};
  __device__ __forceinline__ void load(int64_t, Regs&) const {}
  __device__ __forceinline__ void run(int64_t, int64_t, double2, double2, Regs&, double (&)[1]) const {}
};

// acc[0] += x.y
struct EpiXY {
  struct Regs {};
  __device__ __forceinline__ void load(int64_t, Regs&) const {}
  __device__ __forceinline__ void run(int64_t e0, int64_t n, double2 x, double2 y, Regs&, double (&acc)[1]) const {
    acc_pair(acc[0], e0, n, __dmul_rn(x.x, y.x), __dmul_rn(x.y, y.y));
  }
};

// acc[0] += u.y  (u: a third vector, e.g. r~ in BiCGStab)
struct EpiUY {
  const double* u;
  struct Regs {
    double2 u;
  };
  __device__ __forceinline__ void load(int64_t e0, Regs& r) const { r.u = ld2(u, e0); }
  __device__ __forceinline__ void run(int64_t e0, int64_t n, double2, double2 y, Regs& r, double (&acc)[1]) const {
    acc_pair(acc[0], e0, n, __dmul_rn(r.u.x, y.x), __dmul_rn(r.u.y, y.y));
  }
};

// acc[0] += y.y ; acc[1] += y.x   (BiCGStab: <t,t>, <t,r>)
struct EpiYYandYX {
  struct Regs {};
  __device__ __forceinline__ void load(int64_t, Regs&) const {}
  __device__ __forceinline__ void run(int64_t e0, int64_t n, double2 x, double2 y, Regs&, double (&acc)[2]) const {
    acc_pair(acc[0], e0, n, __dmul_rn(y.x, y.x), __dmul_rn(y.y, y.y));
    acc_pair(acc[1], e0, n, __dmul_rn(y.x, x.x), __dmul_rn(y.y, x.y));
  }
};

// r = b - A x fused into the apply (Operator::Residual, Operator.hpp:95-99), with acc[0] += r.r.
// The kernel stores the value returned through `out`.
struct EpiResidual {
  const double* b;
  struct Regs {
    double2 b;
  };
  __device__ __forceinline__ void load(int64_t e0, Regs& r) const { r.b = ld2(b, e0); }
};

template<int FORM, int W>
__device__ __forceinline__ double2 apply_rows(const OpDev& op, const double* __restrict__ x, int64_t e0, double2 xo) {
  const int64_t h = e0 >> 1, ldh = op.ld >> 1;
  const int2* __restrict__ col2 = reinterpret_cast<const int2*>(op.col);
  const double2* __restrict__ a2 = reinterpret_cast<const double2*>(op.v0);
  int2 c[W];
  double2 a[W];
#pragma unroll
  for (int k = 0; k < W; ++k) c[k] = col2[k * ldh + h], a[k] = a2[k * ldh + h];
  double2 out;
  if constexpr (FORM == SB_FORM_COEF) {
    const double2 dg = ld2(op.diag, e0);
    double g0[W], g1[W];
#pragma unroll
    for (int k = 0; k < W; ++k) {
      g0[k] = (c[k].x >= 0) ? __ldg(x + c[k].x) : 0.0;
      g1[k] = (c[k].y >= 0) ? __ldg(x + c[k].y) : 0.0;
    }
    double u0 = __dmul_rn(dg.x, xo.x), u1 = __dmul_rn(dg.y, xo.y);
#pragma unroll
    for (int k = 0; k < W; ++k) {
      const double t0 = __dadd_rn(u0, __dmul_rn(a[k].x, g0[k]));
      const double t1 = __dadd_rn(u1, __dmul_rn(a[k].y, g1[k]));
      u0 = (c[k].x >= 0) ? t0 : u0;
      u1 = (c[k].y >= 0) ? t1 : u1;
    }
    out = make_double2(u0, u1);
  } else {
    const double2* __restrict__ d2 = reinterpret_cast<const double2*>(op.v1);
    double2 d[W];
#pragma unroll
    for (int k = 0; k < W; ++k) d[k] = d2[k * ldh + h];
    double g0[W], g1[W];
#pragma unroll
    for (int k = 0; k < W; ++k) {
      // ghost entry (col == ~i): mirror state -x[i]; padding: skipped below
      g0[k] = (c[k].x >= 0) ? __ldg(x + c[k].x) : -xo.x;
      g1[k] = (c[k].y >= 0) ? __ldg(x + c[k].y) : -xo.y;
    }
    double u0 = op.prefill ? xo.x : 0.0, u1 = op.prefill ? xo.y : 0.0;
#pragma unroll
    for (int k = 0; k < W; ++k) {
      // flux = dt*(x_nbr - x_i)/dist ; u += (area/vol)*flux      (Playground.cpp:125-128)
      const double f0 = __ddiv_rn(__dmul_rn(op.dt, __dsub_rn(g0[k], xo.x)), d[k].x);
      const double f1 = __ddiv_rn(__dmul_rn(op.dt, __dsub_rn(g1[k], xo.y)), d[k].y);
      const double t0 = __dadd_rn(u0, __dmul_rn(a[k].x, f0));
      const double t1 = __dadd_rn(u1, __dmul_rn(a[k].y, f1));
      u0 = (c[k].x != kColPad) ? t0 : u0;
      u1 = (c[k].y != kColPad) ? t1 : u1;
    }
    out = make_double2(u0, u1);
  }
  return out;
}

// y <- A x with a fused reduction epilogue. RESID: store b - A x instead (and reduce <r,r>).
template<int FORM, int W, int ND, bool RESID, class Epi, class Final>
__global__ void __launch_bounds__(kThreads) apply_kernel(OpDev op, const double* __restrict__ x, double* __restrict__ y,
                                                         Epi epi, RedPtrs red, Final fin,
                                                         const int* __restrict__ done) {
  if (done != nullptr && *done != 0) return;
  double acc[ND > 0 ? ND : 1];
#pragma unroll
  for (int d = 0; d < (ND > 0 ? ND : 1); ++d) acc[d] = 0.0;
#pragma unroll 1
  for (int j = 0; j < kSub; ++j) {
    const int64_t e0 = lane_elem(j);
    const double2 xo = ld2(x, e0);
    typename Epi::Regs er;
    epi.load(e0, er);
    double2 out = apply_rows<FORM, W>(op, x, e0, xo);
    if constexpr (RESID) {
      out.x = __dsub_rn(er.b.x, out.x);
      out.y = __dsub_rn(er.b.y, out.y);
      acc_pair(acc[0], e0, op.n, __dmul_rn(out.x, out.x), __dmul_rn(out.y, out.y));
    }
    st2(y, e0, out);
    if constexpr (!RESID) epi.run(e0, op.n, xo, out, er, acc);
  }
  if constexpr (ND > 0) block_reduce_finalize<ND>(acc, red, fin);
}

} // namespace sb

struct sb_op {
  sb::OpDev d;
  int64_t n_entries = 0;
  int64_t device_bytes = 0;
  void* buffers[4] = {nullptr, nullptr, nullptr, nullptr};
};

namespace sb {

// Launch y <- A x (+ epilogue) on the context's stream, dispatching on form and ELL width.
template<int ND, bool RESID, class Epi, class Final>
int launch_apply(sb_ctx* ctx, const sb_op* op, const double* x, double* y, const Epi& epi, const Final& fin,
                 const int* done) {
  const OpDev& d = op->d;
  const unsigned grid = (unsigned) num_tiles(d.n);
  if constexpr (ND > 0) {
    SB_TRY(ensure_red_scratch(ctx, d.n));
  }
  const RedPtrs red{ctx->red.partials, ctx->red.cap_tiles, ctx->red.ticket};
#define SB_LAUNCH(FORM, W)                                                                       \
  apply_kernel<FORM, W, ND, RESID, Epi, Final><<<grid, kThreads, 0, ctx->stream>>>(d, x, y, epi, red, fin, done)
#define SB_WIDTHS(FORM)                   \
  switch (d.width) {                      \
    case 0: case 1: SB_LAUNCH(FORM, 1); break; \
    case 2: SB_LAUNCH(FORM, 2); break;    \
    case 3: SB_LAUNCH(FORM, 3); break;    \
    case 4: SB_LAUNCH(FORM, 4); break;    \
    case 5: SB_LAUNCH(FORM, 5); break;    \
    case 6: SB_LAUNCH(FORM, 6); break;    \
    case 7: SB_LAUNCH(FORM, 7); break;    \
    case 8: SB_LAUNCH(FORM, 8); break;    \
    default:                              \
      set_error("operator width %d not supported (max 8)", d.width); \
      return SB_ERR_INVALID;              \
  }
  if (d.form == SB_FORM_COEF) {
    SB_WIDTHS(SB_FORM_COEF)
  } else {
    SB_WIDTHS(SB_FORM_FAITHFUL)
  }
#undef SB_WIDTHS
#undef SB_LAUNCH
  ctx->launches++;
  SB_CUDA(cudaGetLastError());
  return SB_OK;
}

} // namespace sb
