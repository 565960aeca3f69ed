// TEST INFRASTRUCTURE -- not product code.
//
// Compiles the product's OWN per-row device code of the operator apply -- OpDev, gather(), apply_rows<FORM, W>,
// stormruler_b200/csrc/sb_apply_rows.cuh, included verbatim -- for the host and runs it over every row pair the way
// apply_kernel does (csrc/sb_op.cuh: a lane owns rows e0, e0+1; prefill 2 reads the old y first). The host stand-ins of
// the CUDA primitives round every operation separately (built with -ffp-contract=off). tests/test_dropin_emulated.py
// feeds it the oracle's rows in the product's layout contract and compares with the oracle's FACE LOOP: the faithful
// form (incl. the accumulate mode of sb_apply_accumulate) must reproduce it bit for bit, the coefficient form its
// row oracle. The device run then only has to show that the kernel around this code loads and stores the right rows.
#include <climits>
#include <cstdint>

#include "../../include/stormb200.h"

#define __device__
#define __forceinline__ inline

struct double2 {
  double x, y;
};
struct int2 {
  int x, y;
};
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __ldca(const double* p) { return *p; }
static inline double __ldg(const double* p) { return *p; }

namespace sb {
constexpr int32_t kColPad = INT32_MIN; // csrc/sb_common.cuh
static inline double2 ld2(const double* p, int64_t e) { return double2{p[e], p[e + 1]}; }
#include "../../stormruler_b200/csrc/sb_apply_rows.cuh"

template<int FORM, int W>
static void run(const OpDev& op, const double* x, double* y) {
  for (int64_t e0 = 0; e0 < op.ld; e0 += 2) {
    const double2 xo = ld2(x, e0);
    double2 yo = make_double2(0.0, 0.0);
    if (FORM == SB_FORM_FAITHFUL && op.prefill == 2) yo = ld2(y, e0);
    const double2 out = apply_rows<FORM, W>(op, x, e0, xo, yo, 0);
    y[e0] = out.x, y[e0 + 1] = out.y;
  }
}
} // namespace sb

// col / v0 / v1: [width][ld] column-major ELL as the product uploads them; x, y: capacity ld (ld even).
extern "C" __attribute__((visibility("default"))) int apply_rows_host(int form, int width, int64_t n, int64_t ld,
                                                                       const int32_t* col, const double* v0,
                                                                       const double* v1, const double* diag, int prefill,
                                                                       double dt, const double* x, double* y) {
  sb::OpDev op;
  op.n = n, op.ld = ld, op.width = width, op.form = form, op.prefill = prefill, op.dt = dt;
  op.col = col, op.v0 = v0, op.v1 = v1, op.diag = diag;
  if (ld % 2 != 0) return -2;
#define CASE(W)                                                         \
  case W:                                                               \
    if (form == SB_FORM_COEF) sb::run<SB_FORM_COEF, W>(op, x, y);       \
    else sb::run<SB_FORM_FAITHFUL, W>(op, x, y);                        \
    return 0;
  switch (width) {
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(10) CASE(12) CASE(14) CASE(16)
  }
#undef CASE
  return -1;
}
