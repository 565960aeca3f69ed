/* TEST / ANALYSIS INFRASTRUCTURE -- not product code, computes nothing.
 *
 * A logging stand-in for the 20 C-ABI entry points that stormruler_b200/host/dropin.cpp imports (include/stormb200.h).
 * The drop-in TU -- the reference's unmodified solver templates on Storm::DeviceVector -- is linked against this
 * library instead of libstormb200.so (oracle/Makefile, target `trace`), so every vector statement, reduction and
 * operator apply the reference's solvers issue shows up as one line of a log, with vector identities instead of data:
 *     eval  y aop n_vec v0 v1 ...     y (op)= expr over these distinct vector operands
 *     fill  y | copy y x | dot a b | norm a | apply y x | accum y x | jacobi y x | alloc v | free v | upload v
 * Vectors are numbered in order of first appearance. Reductions return 1.0 (solvers run with tolerances 0, so the
 * statement stream does not depend on values). oracle/statement_trace.py turns the log into vector-pass counts per
 * iteration: as written, and for the schedule a statement-fusing backend would execute. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/stormb200.h"

#define API __attribute__((visibility("default")))

static char* g_log = NULL;
static size_t g_len = 0, g_cap = 0;
static const void* g_ids[4096];
static int g_n_ids = 0;

static int id_of(const void* p) {
  for (int k = 0; k < g_n_ids; ++k)
    if (g_ids[k] == p) return k;
  if (g_n_ids < 4096) g_ids[g_n_ids] = p;
  return g_n_ids++;
}

static void put(const char* s) {
  const size_t n = strlen(s);
  if (g_len + n + 1 > g_cap) {
    g_cap = (g_cap + n + 1) * 2;
    g_log = (char*) realloc(g_log, g_cap);
  }
  memcpy(g_log + g_len, s, n + 1);
  g_len += n;
}

API void sbtrace_reset(void) {
  g_len = 0, g_n_ids = 0;
  if (g_log) g_log[0] = 0;
}
API const char* sbtrace_log(void) { return g_log ? g_log : ""; }
API void sbtrace_mark(const char* what) {
  char line[128];
  snprintf(line, sizeof line, "mark %s\n", what);
  put(line);
}

API const char* sb_last_error(void) { return "statement tracer: entry point not simulated"; }

API int sb_vec_alloc(sb_ctx* ctx, size_t n, double** d_out) {
  (void) ctx, (void) n;
  *d_out = (double*) malloc(16); /* a unique address; never dereferenced */
  char line[64];
  snprintf(line, sizeof line, "alloc %d\n", id_of(*d_out));
  put(line);
  return SB_OK;
}
API int sb_vec_free(sb_ctx* ctx, double* d) {
  (void) ctx;
  if (d == NULL) return SB_OK;
  char line[64];
  snprintf(line, sizeof line, "free %d\n", id_of(d));
  put(line);
  /* not returned to malloc: the address stays unique for the whole trace */
  return SB_OK;
}
API int sb_vec_upload(sb_ctx* ctx, double* d, const double* h_src, size_t n) {
  (void) ctx, (void) h_src, (void) n;
  char line[64];
  snprintf(line, sizeof line, "upload %d\n", id_of(d));
  put(line);
  return SB_OK;
}
API int sb_vec_download(sb_ctx* ctx, const double* d, double* h_dst, size_t n) {
  (void) ctx;
  char line[64];
  snprintf(line, sizeof line, "download %d\n", id_of(d));
  put(line);
  memset(h_dst, 0, n * sizeof(double));
  return SB_OK;
}
API int sb_eval(sb_ctx* ctx, double* y, size_t n, int assign_op, const sb_expr* e) {
  (void) ctx, (void) n;
  char line[256];
  int used[SB_EXPR_MAX_VEC] = {0, 0, 0, 0}, nv = 0;
  for (int k = 0; k < e->n_ops; ++k)
    if (e->ops[k] <= SB_OP_VEC3) used[e->ops[k]] = 1;
  for (int k = 0; k < SB_EXPR_MAX_VEC; ++k) nv += used[k];
  int off = snprintf(line, sizeof line, "eval %d %d %d", id_of(y), assign_op, nv);
  for (int k = 0; k < SB_EXPR_MAX_VEC; ++k)
    if (used[k]) off += snprintf(line + off, sizeof line - (size_t) off, " %d", id_of(e->vec[k]));
  snprintf(line + off, sizeof line - (size_t) off, "\n");
  put(line);
  if (getenv("SBTRACE_OPS") != NULL) { /* debugging aid: the postfix program itself */
    fprintf(stderr, "[sbtrace] eval ops:");
    for (int k = 0; k < e->n_ops; ++k) fprintf(stderr, " %d", e->ops[k]);
    fprintf(stderr, "\n");
  }
  return SB_OK;
}
/* group n_reads r... n_writes w... n_dots : one launch; reads = distinct vectors read before the group wrote them
 * (operands of the statements and of the dots), writes = distinct targets */
API int sb_eval_group(sb_ctx* ctx, size_t n, int n_stmt, const sb_chain* st, int n_dots, const double* const* da,
                      const double* const* db, double* h_out) {
  (void) ctx, (void) n;
  const void* rd[128];
  const void* wr[16];
  int nr = 0, nw = 0;
#define SEEN(arr, cnt, p) ({ int f_ = 0; for (int q_ = 0; q_ < (cnt); ++q_) f_ |= (arr)[q_] == (const void*) (p); f_; })
  for (int s = 0; s < n_stmt; ++s) {
    if (st[s].base != NULL && !SEEN(wr, nw, st[s].base) && !SEEN(rd, nr, st[s].base)) rd[nr++] = st[s].base;
    for (int t = 0; t < st[s].n_terms; ++t)
      if (!SEEN(wr, nw, st[s].x[t]) && !SEEN(rd, nr, st[s].x[t])) rd[nr++] = st[s].x[t];
    if (!SEEN(wr, nw, st[s].y)) wr[nw++] = st[s].y;
  }
  for (int d = 0; d < n_dots; ++d) {
    if (!SEEN(wr, nw, da[d]) && !SEEN(rd, nr, da[d])) rd[nr++] = da[d];
    if (!SEEN(wr, nw, db[d]) && !SEEN(rd, nr, db[d])) rd[nr++] = db[d];
    h_out[d] = 1.0;
  }
#undef SEEN
  char line[1024];
  int off = snprintf(line, sizeof line, "group %d", nr);
  for (int q = 0; q < nr; ++q) off += snprintf(line + off, sizeof line - (size_t) off, " %d", id_of(rd[q]));
  off += snprintf(line + off, sizeof line - (size_t) off, " %d", nw);
  for (int q = 0; q < nw; ++q) off += snprintf(line + off, sizeof line - (size_t) off, " %d", id_of(wr[q]));
  snprintf(line + off, sizeof line - (size_t) off, " %d %d\n", n_stmt, n_dots);
  put(line);
  return SB_OK;
}

API int sb_fill(sb_ctx* ctx, double* y, size_t n, double value) {
  (void) ctx, (void) n, (void) value;
  char line[64];
  snprintf(line, sizeof line, "fill %d\n", id_of(y));
  put(line);
  return SB_OK;
}
API int sb_copy(sb_ctx* ctx, double* y, const double* x, size_t n) {
  (void) ctx, (void) n;
  char line[64];
  snprintf(line, sizeof line, "copy %d %d\n", id_of(y), id_of(x));
  put(line);
  return SB_OK;
}
API int sb_dot(sb_ctx* ctx, const double* a, const double* b, size_t n, double* h_out) {
  (void) ctx, (void) n;
  char line[64];
  snprintf(line, sizeof line, "dot %d %d\n", id_of(a), id_of(b));
  put(line);
  *h_out = 1.0;
  return SB_OK;
}
API int sb_norm2(sb_ctx* ctx, const double* a, size_t n, double* h_out) {
  (void) ctx, (void) n;
  char line[64];
  snprintf(line, sizeof line, "norm %d\n", id_of(a));
  put(line);
  *h_out = 1.0;
  return SB_OK;
}
API int sb_apply(sb_ctx* ctx, const sb_op* op, const double* x, double* y) {
  (void) ctx, (void) op;
  char line[64];
  snprintf(line, sizeof line, "apply %d %d\n", id_of(y), id_of(x));
  put(line);
  return SB_OK;
}
/* applydot y x u: an apply with one dot riding on it (u: the dot's operand besides y; == x when it is the input) */
API int sb_apply_dot(sb_ctx* ctx, const sb_op* op, const double* x, double* y, const double* u, double* h_out) {
  (void) ctx, (void) op;
  char line[96];
  snprintf(line, sizeof line, "applydot %d %d %d\n", id_of(y), id_of(x), id_of(u != NULL ? u : x));
  put(line);
  *h_out = 1.0;
  return SB_OK;
}
/* applydot2 y x: an apply with <y,y> and <y,x> riding on it */
API int sb_apply_dot_yy_yx(sb_ctx* ctx, const sb_op* op, const double* x, double* y, double* h_out) {
  (void) ctx, (void) op;
  char line[96];
  snprintf(line, sizeof line, "applydot2 %d %d\n", id_of(y), id_of(x));
  put(line);
  h_out[0] = h_out[1] = 1.0;
  return SB_OK;
}
API int sb_apply_accumulate(sb_ctx* ctx, const sb_op* op, double dt, const double* x, double* y) {
  (void) ctx, (void) op, (void) dt;
  char line[64];
  snprintf(line, sizeof line, "accum %d %d\n", id_of(y), id_of(x));
  put(line);
  return SB_OK;
}
API int sb_op_jacobi(sb_ctx* ctx, const sb_op* op, const double* x, double* y) {
  (void) ctx, (void) op;
  char line[64];
  snprintf(line, sizeof line, "jacobi %d %d\n", id_of(y), id_of(x));
  put(line);
  return SB_OK;
}
API int sb_op_destroy(sb_ctx* ctx, sb_op* op) {
  (void) ctx, (void) op;
  return SB_OK;
}
/* the fused solvers are not statement streams: not simulated */
API int sb_cg_solve(sb_ctx* c, const sb_op* o, double* x, const double* b, const sb_solver_opts* p, sb_solver_report* r,
                    double* h, int64_t hc, double* t, int64_t tc) {
  (void) c, (void) o, (void) x, (void) b, (void) p, (void) r, (void) h, (void) hc, (void) t, (void) tc;
  return SB_ERR_STATE;
}
API int sb_bicgstab_solve(sb_ctx* c, const sb_op* o, double* x, const double* b, const sb_solver_opts* p,
                          sb_solver_report* r, double* h, int64_t hc, double* t, int64_t tc) {
  (void) c, (void) o, (void) x, (void) b, (void) p, (void) r, (void) h, (void) hc, (void) t, (void) tc;
  return SB_ERR_STATE;
}
API int sb_gmres_solve(sb_ctx* c, const sb_op* o, double* x, const double* b, const sb_gmres_opts* p, sb_solver_report* r,
                       double* h, int64_t hc, double* t, int64_t tc) {
  (void) c, (void) o, (void) x, (void) b, (void) p, (void) r, (void) h, (void) hc, (void) t, (void) tc;
  return SB_ERR_STATE;
}
