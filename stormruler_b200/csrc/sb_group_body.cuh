// sb_group_body.cuh -- the element-wise body of sb_eval_group. Included by sb_kernels.cuh inside namespace sb, after
// ld2 / st2 / acc_pair / kMaxDots; kept in a file of its own, free of anything CUDA-specific beyond the __device__
// qualifiers and the rounding intrinsics, so that the test infrastructure can compile this very code for the host
// (oracle/emu/group_body_host.cpp) and check it against the emulator of the C ABI without a GPU.
#pragma once

// ---- statement group: linear-combination chains + trailing dots in one pass (sb_eval_group) ------------------------
// A thread runs every statement on its own elements in order, through memory: a later statement that reads what an
// earlier one wrote re-loads the line the same thread has just stored (an L2 hit, not HBM traffic), so no forwarding
// logic is needed and aliasing between statements is correct by program order. The operand pointers are read through
// the kernel-parameter bank with uniform (per-CTA constant) indices.
struct GroupBody {
  sb_chain st[SB_GROUP_MAX_STMT];
  int32_t n_stmt, n_dots;
  const double* da[kMaxDots];
  const double* db[kMaxDots];
  struct Regs {};
  __device__ __forceinline__ void load(int64_t, Regs&) const {}
  template<int ND>
  __device__ __forceinline__ void run(int64_t e0, int64_t n, Regs&, double (&acc)[ND]) const {
#pragma unroll 1
    for (int s = 0; s < n_stmt; ++s) {
      const sb_chain& ch = st[s];
      double2 x[SB_GROUP_MAX_TERMS];
      double2 a = make_double2(0.0, 0.0);
      const bool has_base = ch.base != nullptr;
      if (has_base) a = ld2(ch.base, e0);
#pragma unroll
      for (int t = 0; t < SB_GROUP_MAX_TERMS; ++t)
        if (t < ch.n_terms) x[t] = ld2(ch.x[t], e0);   // all loads of the statement before its arithmetic
#pragma unroll
      for (int t = 0; t < SB_GROUP_MAX_TERMS; ++t)
        if (t < ch.n_terms) {
          const double px = __dmul_rn(ch.c[t], x[t].x), py = __dmul_rn(ch.c[t], x[t].y);
          if (t == 0 && !has_base) a = make_double2(px, py);
          else if (ch.sub[t]) a = make_double2(__dsub_rn(a.x, px), __dsub_rn(a.y, py));
          else a = make_double2(__dadd_rn(a.x, px), __dadd_rn(a.y, py));
        }
      st2(ch.y, e0, a);
    }
#pragma unroll
    for (int d = 0; d < ND; ++d)
      if (d < n_dots) {
        const double2 p = ld2(da[d], e0), q = ld2(db[d], e0);
        acc_pair(acc[d], e0, n, __dmul_rn(p.x, q.x), __dmul_rn(p.y, q.y));
      }
  }
};

