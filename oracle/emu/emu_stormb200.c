/* TEST INFRASTRUCTURE -- not product code; never linked into libstormb200.so or loaded by the product.
 *
 * A HOST-EXECUTING stand-in for the C-ABI entry points that stormruler_b200/host/dropin.cpp imports
 * (include/stormb200.h). It restates, in plain C on malloc'ed arrays, the semantic contract each entry point
 * documents in the header -- element-wise statements evaluated per element in program order with every operation
 * rounded separately, reductions by the oracle (sequential like the reference, or the restated GPU tree), the
 * operator through a callback into the oracle's face loop. The drop-in TU -- the reference's unmodified solver
 * templates on Storm::DeviceVector, plus the playground's Cahn-Hilliard step -- is linked against this library
 * instead of libstormb200.so (oracle/Makefile, target `emu`). That lets the CPU test-suite check, without a GPU,
 * that the C++23 host layer (expression flattening, overload set, traced map(), statement sequencing, solver
 * plumbing) issues exactly the reference's arithmetic: with sequential reductions the result must equal the
 * reference's own run bit for bit (tests/test_dropin_emulated.py). The GPU tests then only have to show that the
 * CUDA kernels honour the same per-entry contract.
 *
 * sb_cg_solve / sb_bicgstab_solve answer with the oracle's restatements of the same contract (see below); sb_gmres_solve is
 * not emulated. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/stormb200.h"
#include "../sb_oracle.h"

#define API __attribute__((visibility("default")))

/* The emulator's operator handle: what `sb_op*` points to when the drop-in runs on it. */
struct sb_op {
  int64_t n;
  orc_apply_fn apply;       /* y <- A(x) */
  void* apply_user;
  const orc_face_op* faces; /* for sb_apply_accumulate; may be NULL */
  const double* diag;       /* for sb_op_jacobi; may be NULL */
};

static char g_error[256] = "";
static int g_mode = ORC_RED_SEQ;
static int64_t g_calls[8]; /* eval, fill, copy, dot, norm, apply, accumulate, jacobi */
static int64_t g_groups; /* sb_eval_group calls */
static int64_t g_apply_dots; /* sb_apply_dot calls */
static int g_dummy_ctx;

static int fail(int code, const char* what) {
  snprintf(g_error, sizeof g_error, "%s", what);
  return code;
}

/* ---- emulator control (called by oracle/emu.py) ------------------------------------------------------------------ */
API sb_ctx* emu_ctx(void) { return (sb_ctx*) &g_dummy_ctx; }
API void emu_set_reduction_mode(int mode) { g_mode = mode; }
API void emu_reset_counts(void) { memset(g_calls, 0, sizeof g_calls), g_groups = 0, g_apply_dots = 0; }
API int64_t emu_apply_dot_count(void) { return g_apply_dots; }
API int64_t emu_group_count(void) { return g_groups; }
API void emu_get_counts(int64_t out[8]) { memcpy(out, g_calls, sizeof g_calls); }
API sb_op* emu_op_create(int64_t n, orc_apply_fn apply, void* apply_user, const orc_face_op* faces, const double* diag) {
  struct sb_op* op = (struct sb_op*) calloc(1, sizeof *op);
  op->n = n, op->apply = apply, op->apply_user = apply_user, op->faces = faces, op->diag = diag;
  return op;
}
API void emu_op_free(sb_op* op) { free(op); }

/* ---- the C ABI --------------------------------------------------------------------------------------------------- */
API const char* sb_last_error(void) { return g_error; }

API int sb_vec_alloc(sb_ctx* ctx, size_t n, double** d_out) {
  if (ctx == NULL || d_out == NULL) return fail(SB_ERR_INVALID, "null argument");
  *d_out = (double*) calloc(n > 0 ? n : 1, sizeof(double)); /* zero-filled, like the device allocation */
  return *d_out != NULL ? SB_OK : fail(SB_ERR_CUDA, "out of memory");
}
API int sb_vec_free(sb_ctx* ctx, double* d) {
  (void) ctx;
  free(d);
  return SB_OK;
}
API int sb_vec_upload(sb_ctx* ctx, double* d, const double* h_src, size_t n) {
  (void) ctx;
  memcpy(d, h_src, n * sizeof(double));
  return SB_OK;
}
API int sb_vec_download(sb_ctx* ctx, const double* d, double* h_dst, size_t n) {
  (void) ctx;
  memcpy(h_dst, d, n * sizeof(double));
  return SB_OK;
}

/* y (op)= expr: the validation rules and the evaluation order of sb_eval (csrc/sb_api.cu, sb_kernels.cuh). */
API int sb_eval(sb_ctx* ctx, double* y, size_t n, int assign_op, const sb_expr* e) {
  if (ctx == NULL || y == NULL || e == NULL) return fail(SB_ERR_INVALID, "null argument");
  if (e->n_ops < 1 || e->n_ops > SB_EXPR_MAX_OPS) return fail(SB_ERR_INVALID, "expression length out of range");
  if (assign_op < SB_ASSIGN || assign_op > SB_DIV_ASSIGN) return fail(SB_ERR_INVALID, "unknown assign_op");
  int depth = 0;
  for (int k = 0; k < e->n_ops; ++k) {
    const int op = e->ops[k];
    if (op >= SB_OP_VEC0 && op <= SB_OP_VEC3) {
      if (e->vec[op] == NULL) return fail(SB_ERR_INVALID, "expression references a null vector operand");
      depth++;
    } else if (op >= SB_OP_SCAL0 && op <= SB_OP_SCAL3) {
      depth++;
    } else if (op == SB_OP_NEG) {
      if (depth < 1) return fail(SB_ERR_INVALID, "expression stack underflow");
    } else if (op >= SB_OP_ADD && op <= SB_OP_DIV) {
      if (depth < 2) return fail(SB_ERR_INVALID, "expression stack underflow");
      depth--;
    } else {
      return fail(SB_ERR_INVALID, "unknown opcode");
    }
    if (depth > 6) return fail(SB_ERR_INVALID, "expression too deep (max stack depth 6)");
  }
  if (depth != 1) return fail(SB_ERR_INVALID, "expression must leave exactly one value");
  g_calls[0]++;
  for (size_t i = 0; i < n; ++i) {
    volatile double st[8]; /* volatile: every intermediate is a rounded fp64, never kept in a wider register */
    int sp = 0;
    for (int k = 0; k < e->n_ops; ++k) {
      const int op = e->ops[k];
      if (op <= SB_OP_VEC3) st[sp++] = e->vec[op][i];
      else if (op <= SB_OP_SCAL3) st[sp++] = e->scal[op - SB_OP_SCAL0];
      else if (op == SB_OP_NEG) st[sp - 1] = -st[sp - 1];
      else {
        const double b = st[--sp], a = st[sp - 1];
        st[sp - 1] = op == SB_OP_ADD ? a + b : op == SB_OP_SUB ? a - b : op == SB_OP_MUL ? a * b : a / b;
      }
    }
    const double v = st[0];
    switch (assign_op) {
      case SB_ASSIGN: y[i] = v; break;
      case SB_ADD_ASSIGN: y[i] = y[i] + v; break;
      case SB_SUB_ASSIGN: y[i] = y[i] - v; break;
      case SB_MUL_ASSIGN: y[i] = y[i] * v; break;
      default: y[i] = y[i] / v; break;
    }
  }
  return SB_OK;
}

/* statement group: statements in order (each over all elements: element-wise statements commute with the element loop),
 * then the dots over the final values -- the contract of sb_eval_group in include/stormb200.h */
API int sb_eval_group(sb_ctx* ctx, size_t n, int n_stmt, const sb_chain* st, int n_dots, const double* const* da,
                      const double* const* db, double* h_out) {
  if (ctx == NULL) return fail(SB_ERR_INVALID, "null argument");
  if (n_stmt < 0 || n_stmt > SB_GROUP_MAX_STMT || n_dots < 0 || n_dots > SB_GROUP_MAX_DOTS || n_stmt + n_dots == 0)
    return fail(SB_ERR_INVALID, "statement / dot count out of range");
  for (int s = 0; s < n_stmt; ++s) {
    if (st[s].y == NULL || st[s].n_terms < 1 || st[s].n_terms > SB_GROUP_MAX_TERMS) return fail(SB_ERR_INVALID, "bad chain");
    if (st[s].base == NULL && st[s].sub[0] != 0) return fail(SB_ERR_INVALID, "a chain without a base cannot start with a subtraction");
    for (int t = 0; t < st[s].n_terms; ++t)
      if (st[s].x[t] == NULL) return fail(SB_ERR_INVALID, "null term vector");
  }
  g_calls[0] += n_stmt; /* counted as the statements they stand for */
  g_groups++;
  for (int s = 0; s < n_stmt; ++s) {
    const sb_chain* ch = &st[s];
    for (size_t i = 0; i < n; ++i) {
      volatile double a = ch->base != NULL ? ch->base[i] : 0.0;
      for (int t = 0; t < ch->n_terms; ++t) {
        volatile double p = ch->c[t] * ch->x[t][i];
        if (t == 0 && ch->base == NULL) a = p;
        else if (ch->sub[t]) a = a - p;
        else a = a + p;
      }
      ch->y[i] = a;
    }
  }
  for (int d = 0; d < n_dots; ++d) {
    g_calls[3]++;
    h_out[d] = orc_dot((int64_t) n, da[d], db[d], g_mode);
  }
  return SB_OK;
}

API int sb_fill(sb_ctx* ctx, double* y, size_t n, double value) {
  (void) ctx;
  g_calls[1]++;
  for (size_t i = 0; i < n; ++i) y[i] = value;
  return SB_OK;
}
API int sb_copy(sb_ctx* ctx, double* y, const double* x, size_t n) {
  (void) ctx;
  g_calls[2]++;
  if (x != y) memmove(y, x, n * sizeof(double));
  return SB_OK;
}
API int sb_dot(sb_ctx* ctx, const double* a, const double* b, size_t n, double* h_out) {
  (void) ctx;
  g_calls[3]++;
  *h_out = orc_dot((int64_t) n, a, b, g_mode);
  return SB_OK;
}
API int sb_norm2(sb_ctx* ctx, const double* a, size_t n, double* h_out) {
  (void) ctx;
  g_calls[4]++;
  *h_out = orc_norm2((int64_t) n, a, g_mode);
  return SB_OK;
}
API int sb_apply(sb_ctx* ctx, const sb_op* op, const double* x, double* y) {
  if (ctx == NULL || op == NULL || x == NULL || y == NULL) return fail(SB_ERR_INVALID, "null argument");
  if (x == y) return fail(SB_ERR_INVALID, "sb_apply: x and y must not alias");
  g_calls[5]++;
  op->apply(op->apply_user, y, x, (size_t) op->n);
  return SB_OK;
}
API int sb_apply_dot(sb_ctx* ctx, const sb_op* op, const double* x, double* y, const double* u, double* h_out) {
  if (ctx == NULL || op == NULL || x == NULL || y == NULL || h_out == NULL) return fail(SB_ERR_INVALID, "null argument");
  if (x == y || u == y) return fail(SB_ERR_INVALID, "sb_apply_dot: y must not alias x or u");
  g_calls[5]++, g_calls[3]++;
  g_apply_dots++;
  op->apply(op->apply_user, y, x, (size_t) op->n);
  *h_out = orc_dot(op->n, u != NULL ? u : x, y, g_mode);
  return SB_OK;
}
API int sb_apply_dot_yy_yx(sb_ctx* ctx, const sb_op* op, const double* x, double* y, double* h_out) {
  if (ctx == NULL || op == NULL || x == NULL || y == NULL || h_out == NULL) return fail(SB_ERR_INVALID, "null argument");
  if (x == y) return fail(SB_ERR_INVALID, "sb_apply_dot_yy_yx: x and y must not alias");
  g_calls[5]++, g_calls[3] += 2;
  g_apply_dots++;
  op->apply(op->apply_user, y, x, (size_t) op->n);
  h_out[0] = orc_dot(op->n, y, y, g_mode);
  h_out[1] = orc_dot(op->n, y, x, g_mode);
  return SB_OK;
}
API int sb_apply_accumulate(sb_ctx* ctx, const sb_op* op, double dt, const double* x, double* y) {
  if (ctx == NULL || op == NULL || x == NULL || y == NULL) return fail(SB_ERR_INVALID, "null argument");
  if (x == y) return fail(SB_ERR_INVALID, "sb_apply_accumulate: x and y must not alias");
  if (op->faces == NULL) return fail(SB_ERR_INVALID, "sb_apply_accumulate needs a faithful-form operator");
  g_calls[6]++;
  orc_divgrad_accumulate(op->faces, dt, x, y);
  return SB_OK;
}
API int sb_op_jacobi(sb_ctx* ctx, const sb_op* op, const double* x, double* y) {
  if (ctx == NULL || op == NULL || x == NULL || y == NULL) return fail(SB_ERR_INVALID, "null argument");
  if (op->diag == NULL) return fail(SB_ERR_INVALID, "sb_op_jacobi needs a coefficient-form operator");
  g_calls[7]++;
  for (int64_t i = 0; i < op->n; ++i) y[i] = x[i] / op->diag[i];
  return SB_OK;
}
API int sb_op_destroy(sb_ctx* ctx, sb_op* op) {
  (void) ctx, (void) op;
  return SB_OK;
}

/* The fused CG / BiCGStab are CUDA schedules; their CONTRACT (same statements, same stopping rule, same reduction
 * values as the reference) is what the oracle's plain-C restatements implement, so the emulator answers with those:
 * enough to exercise the C++ classes in front of them (Storm/B200/FusedSolvers.hpp: option and report mapping). */
static int emu_fused(int bicgstab, sb_ctx* c, const sb_op* o, double* x, const double* b, const sb_solver_opts* p,
                     sb_solver_report* r, double* h, int64_t hc, double* t, int64_t tc) {
  if (c == NULL || o == NULL || x == NULL || b == NULL || p == NULL || r == NULL) return fail(SB_ERR_INVALID, "null argument");
  if (x == b) return fail(SB_ERR_INVALID, "x and b must not alias");
  orc_solver_opts oo;
  oo.num_iterations = p->num_iterations, oo.abs_tol = p->abs_tol, oo.rel_tol = p->rel_tol, oo.reduction_mode = g_mode;
  orc_solver_report rep;
  memset(&rep, 0, sizeof rep);
  const int rc = (bicgstab ? orc_bicgstab : orc_cg)(o->apply, o->apply_user, o->n, b, x, &oo, &rep, h, h != NULL ? hc : 0, t,
                                                    t != NULL ? tc : 0);
  if (rc != 0) return fail(SB_ERR_INVALID, "oracle solver failed");
  memset(r, 0, sizeof *r);
  r->converged = rep.converged, r->iterations = rep.iterations;
  r->abs_err = rep.abs_err, r->rel_err = rep.rel_err;
  r->initial_err = (h != NULL && hc > 0) ? h[0] : 0.0;
  r->n_hist = h != NULL ? (rep.n_hist < hc ? rep.n_hist : hc) : 0;
  r->n_trace = t != NULL ? (rep.n_trace < tc ? rep.n_trace : tc) : 0;
  r->n_kernel_slots = bicgstab ? 5 : 3;
  return SB_OK;
}
API int sb_cg_solve(sb_ctx* c, const sb_op* o, double* x, const double* b, const sb_solver_opts* p, sb_solver_report* r,
                    double* h, int64_t hc, double* t, int64_t tc) {
  return emu_fused(0, c, o, x, b, p, r, h, hc, t, tc);
}
API int sb_bicgstab_solve(sb_ctx* c, const sb_op* o, double* x, const double* b, const sb_solver_opts* p,
                          sb_solver_report* r, double* h, int64_t hc, double* t, int64_t tc) {
  return emu_fused(1, c, o, x, b, p, r, h, hc, t, tc);
}
API int sb_gmres_solve(sb_ctx* c, const sb_op* o, double* x, const double* b, const sb_gmres_opts* p, sb_solver_report* r,
                       double* h, int64_t hc, double* t, int64_t tc) {
  (void) c, (void) o, (void) x, (void) b, (void) p, (void) r, (void) h, (void) hc, (void) t, (void) tc;
  return fail(SB_ERR_STATE, "host emulator: the fused solvers are CUDA schedules and are not emulated");
}
