// sb_apply_rows.cuh -- the per-row arithmetic of the operator apply (register-staged kernel): OpDev, gather() and
// apply_rows<FORM, W>. Included by sb_op.cuh inside namespace sb, after ld2 / kColPad / SB_FORM_*; kept in a file of its
// own, free of anything CUDA-specific beyond the __device__ qualifiers, the cache-hinted loads and the rounding
// intrinsics, so that the test infrastructure can compile this very code for the host (oracle/emu/apply_rows_host.cpp)
// and check it against the oracle's face loop without a GPU.
#pragma once

struct OpDev {
  int64_t n = 0, ld = 0;
  int32_t width = 0, form = 0, prefill = 0; // prefill: 0 y = A x from 0, 1 from x, 2 (faithful form, per call) from the old y
  double dt = 0.0;
  const int32_t* col = nullptr;
  const double* v0 = nullptr;
  const double* v1 = nullptr;
  const double* diag = nullptr;
  // coef form, blocked layout ("SELL-64"): one record per 64-row slice,
  //   [ col[W][64] int32 | coef[W][64] fp64 | diag[64] fp64 ]  = 768*W + 512 bytes, contiguous,
  // so a warp stage is ONE bulk copy and the whole operator is a single sequential HBM stream.
  const unsigned char* blk = nullptr;
  int32_t slice_bytes = 0;
  int32_t debug = 0; // experiments only (SB_DEBUG env): bit0 = skip the gathers
  int32_t stream_hint = 0; // per call: the bulk copies of the slice records carry an L2 evict-first policy (SB_TUNE_STREAM_OPERATOR)
  int32_t zero = 0;  // always 0, but only known at run time: lets the kernel tie an instruction to a value it
                     // must wait for without changing the arithmetic (see apply_kernel_tma)
};

// Gathered x: read-only path on one GPU; in P2P mode the halo tail is written by peers while the
// kernel runs, so the coherent path is used (the flag acquire above orders it).
__device__ __forceinline__ double gather(const double* p, int coherent) {
  return coherent ? __ldca(p) : __ldg(p);
}

// Rows wider than 8 entries (polyhedral cells) are processed in chunks of 8: the sum over a row is
// sequential in k either way, so chunking changes the register footprint, not the result.
template<int FORM, int W>
__device__ __forceinline__ double2 apply_rows(const OpDev& op, const double* __restrict__ x, int64_t e0, double2 xo,
                                              double2 yo, int coh) {
  constexpr int C = W <= 8 ? W : 8;
  const int64_t h = e0 >> 1, ldh = op.ld >> 1;
  const int2* __restrict__ col2 = reinterpret_cast<const int2*>(op.col);
  const double2* __restrict__ a2 = reinterpret_cast<const double2*>(op.v0);
  double u0, u1;
  if constexpr (FORM == SB_FORM_COEF) {
    const double2 dg = ld2(op.diag, e0);
    u0 = __dmul_rn(dg.x, xo.x), u1 = __dmul_rn(dg.y, xo.y);
  } else {
    // prefill 2: the face terms are added to what the caller left in y (stormDivGrad as the playground calls it,
    // Playground.cpp:157-165); the row sum then starts from the old y, like the face loop's `u[cell] +=`
    u0 = op.prefill == 2 ? yo.x : (op.prefill ? xo.x : 0.0), u1 = op.prefill == 2 ? yo.y : (op.prefill ? xo.y : 0.0);
  }
#pragma unroll
  for (int k0 = 0; k0 < W; k0 += C) {
    int2 c[C];
    double2 a[C];
#pragma unroll
    for (int k = 0; k < C; ++k)
      if (k0 + k < W) c[k] = col2[(k0 + k) * ldh + h], a[k] = a2[(k0 + k) * ldh + h];
    if constexpr (FORM == SB_FORM_COEF) {
      double g0[C], g1[C];
#pragma unroll
      for (int k = 0; k < C; ++k)
        if (k0 + k < W) {
          g0[k] = (c[k].x >= 0) ? gather(x + c[k].x, coh) : 0.0;
          g1[k] = (c[k].y >= 0) ? gather(x + c[k].y, coh) : 0.0;
        }
#pragma unroll
      for (int k = 0; k < C; ++k)
        if (k0 + k < W) {
          const double t0 = __dadd_rn(u0, __dmul_rn(a[k].x, g0[k]));
          const double t1 = __dadd_rn(u1, __dmul_rn(a[k].y, g1[k]));
          u0 = (c[k].x >= 0) ? t0 : u0;
          u1 = (c[k].y >= 0) ? t1 : u1;
        }
    } else {
      const double2* __restrict__ d2 = reinterpret_cast<const double2*>(op.v1);
      double2 d[C];
#pragma unroll
      for (int k = 0; k < C; ++k)
        if (k0 + k < W) d[k] = d2[(k0 + k) * ldh + h];
      double g0[C], g1[C];
#pragma unroll
      for (int k = 0; k < C; ++k)
        if (k0 + k < W) {
          // ghost entry (col == ~i): mirror state -x[i]; padding: skipped below
          g0[k] = (c[k].x >= 0) ? gather(x + c[k].x, coh) : -xo.x;
          g1[k] = (c[k].y >= 0) ? gather(x + c[k].y, coh) : -xo.y;
        }
#pragma unroll
      for (int k = 0; k < C; ++k)
        if (k0 + k < W) {
          // flux = dt*(x_nbr - x_i)/dist ; u += (area/vol)*flux      (Playground.cpp:125-128)
          const double f0 = __ddiv_rn(__dmul_rn(op.dt, __dsub_rn(g0[k], xo.x)), d[k].x);
          const double f1 = __ddiv_rn(__dmul_rn(op.dt, __dsub_rn(g1[k], xo.y)), d[k].y);
          const double t0 = __dadd_rn(u0, __dmul_rn(a[k].x, f0));
          const double t1 = __dadd_rn(u1, __dmul_rn(a[k].y, f1));
          u0 = (c[k].x != kColPad) ? t0 : u0;
          u1 = (c[k].y != kColPad) ? t1 : u1;
        }
    }
  }
  return make_double2(u0, u1);
}

