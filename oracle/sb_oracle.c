/*
 * sb_oracle.c -- TEST INFRASTRUCTURE. CPU restatement of the StormRuler Krylov hot path.
 * See sb_oracle.h for scope, reference anchors and the build-flag contract.
 * Plain C99, single thread, no dependencies. Compile: gcc -O2 -ffp-contract=off.
 */
#include "sb_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------
 * Operator: face loop.  Follows source_apps/playground/Playground.cpp:115-131 statement by
 * statement; the boundary loop follows the shape of Feathers/ConvectionScheme.hpp:95-106
 * (ghost state, flux into the inner cell only).
 * ---------------------------------------------------------------------------------------------- */
void orc_apply_faces(const orc_face_op* op, const double* x, double* y) {
  const int64_t n = op->n_cells;
  if (op->prefill) {
    for (int64_t i = 0; i < n; ++i) y[i] = x[i]; /* `c_hat <<= c_in` (Playground.cpp:162) */
  } else {
    for (int64_t i = 0; i < n; ++i) y[i] = 0.0;
  }
  const double dt = op->dt;
  for (int64_t f = 0; f < op->n_faces; ++f) {
    const int32_t ci = op->face_cell[2 * f + 0]; /* face.inner_cell() */
    const int32_t co = op->face_cell[2 * f + 1]; /* face.outer_cell() */
    /* flux = dt * (c[outer] - c[inner]) / length(center_outer - center_inner)   (:125-126) */
    const double flux = dt * (x[co] - x[ci]) / op->face_dist[f];
    /* u[inner] += (area / vol_inner) * flux ; u[outer] -= (area / vol_outer) * flux   (:127-128) */
    y[ci] += (op->face_area[f] / op->cell_vol[ci]) * flux;
    y[co] -= (op->face_area[f] / op->cell_vol[co]) * flux;
  }
  for (int64_t b = 0; b < op->n_bfaces; ++b) {
    const int32_t ci = op->bface_cell[b];
    const double ghost = -x[ci]; /* homogeneous Dirichlet mirror state */
    const double flux = dt * (ghost - x[ci]) / op->bface_dist[b];
    y[ci] += (op->bface_area[b] / op->cell_vol[ci]) * flux;
  }
}

/* stormDivGrad(mesh, u, dt, c) as the playground calls it (Playground.cpp:115-131): the face terms are ADDED
 * to whatever the caller left in u (`w_hat <<= f + sigma*(c_in - c)` then stormDivGrad(mesh, w_hat, -Gamma, c_in),
 * :157-159). op->prefill and op->dt are ignored. */
void orc_divgrad_accumulate(const orc_face_op* op, double dt, const double* c, double* u) {
  for (int64_t f = 0; f < op->n_faces; ++f) {
    const int32_t ci = op->face_cell[2 * f + 0];
    const int32_t co = op->face_cell[2 * f + 1];
    const double flux = dt * (c[co] - c[ci]) / op->face_dist[f];
    u[ci] += (op->face_area[f] / op->cell_vol[ci]) * flux;
    u[co] -= (op->face_area[f] / op->cell_vol[co]) * flux;
  }
  for (int64_t b = 0; b < op->n_bfaces; ++b) {
    const int32_t ci = op->bface_cell[b];
    const double ghost = -c[ci];
    const double flux = dt * (ghost - c[ci]) / op->bface_dist[b];
    u[ci] += (op->bface_area[b] / op->cell_vol[ci]) * flux;
  }
}

void orc_apply_faces_cb(void* user, double* y, const double* x, size_t n) {
  (void) n;
  orc_apply_faces((const orc_face_op*) user, x, y);
}

/* ------------------------------------------------------------------------------------------------
 * Cell-row form.
 * ---------------------------------------------------------------------------------------------- */
static int32_t* row_degrees(const orc_face_op* op) {
  int32_t* deg = (int32_t*) calloc((size_t) op->n_cells + 1, sizeof(int32_t));
  for (int64_t f = 0; f < op->n_faces; ++f) {
    deg[op->face_cell[2 * f + 0]]++;
    deg[op->face_cell[2 * f + 1]]++;
  }
  for (int64_t b = 0; b < op->n_bfaces; ++b) deg[op->bface_cell[b]]++;
  return deg;
}

int orc_rows_width(const orc_face_op* op) {
  int32_t* deg = row_degrees(op);
  int w = 0;
  for (int64_t i = 0; i < op->n_cells; ++i)
    if (deg[i] > w) w = deg[i];
  free(deg);
  return w;
}

void orc_build_rows(const orc_face_op* op, int width, int64_t ld, int32_t* col, int64_t* face) {
  const int64_t n = op->n_cells;
  for (int64_t k = 0; k < (int64_t) width * ld; ++k) {
    col[k] = ORC_COL_PAD;
    face[k] = -1;
  }
  int32_t* fill = (int32_t*) calloc((size_t) n + 1, sizeof(int32_t));
  /* Visiting faces in ascending index appends to each row in the order the face loop touches
   * the cell, which is the order its contributions are summed in. */
  for (int64_t f = 0; f < op->n_faces; ++f) {
    const int32_t ci = op->face_cell[2 * f + 0], co = op->face_cell[2 * f + 1];
    col[(int64_t) fill[ci] * ld + ci] = co;
    face[(int64_t) fill[ci] * ld + ci] = f;
    fill[ci]++;
    col[(int64_t) fill[co] * ld + co] = ci;
    face[(int64_t) fill[co] * ld + co] = f;
    fill[co]++;
  }
  for (int64_t b = 0; b < op->n_bfaces; ++b) {
    const int32_t ci = op->bface_cell[b];
    col[(int64_t) fill[ci] * ld + ci] = ~ci;
    face[(int64_t) fill[ci] * ld + ci] = op->n_faces + b;
    fill[ci]++;
  }
  free(fill);
}

void orc_rows_faithful(const orc_face_op* op, int width, int64_t ld, const int64_t* face, double* g,
                       double* d) {
  for (int k = 0; k < width; ++k) {
    for (int64_t i = 0; i < ld; ++i) {
      const int64_t e = (int64_t) k * ld + i;
      g[e] = 0.0;
      d[e] = 1.0;
      if (i >= op->n_cells || face[e] < 0) continue;
      const int64_t f = face[e];
      if (f < op->n_faces) {
        g[e] = op->face_area[f] / op->cell_vol[i];
        d[e] = op->face_dist[f];
      } else {
        g[e] = op->bface_area[f - op->n_faces] / op->cell_vol[i];
        d[e] = op->bface_dist[f - op->n_faces];
      }
    }
  }
}

void orc_apply_rows_faithful(int64_t n, int width, int64_t ld, const int32_t* col, const double* g,
                             const double* d, int32_t prefill, double dt, const double* x,
                             double* y) {
  for (int64_t i = 0; i < n; ++i) {
    const double xi = x[i];
    double u = prefill == 2 ? y[i] : (prefill ? xi : 0.0); /* 2: accumulate onto the old y (stormDivGrad as called) */
    for (int k = 0; k < width; ++k) {
      const int64_t e = (int64_t) k * ld + i;
      const int32_t c = col[e];
      if (c == ORC_COL_PAD) continue;
      const double xn = (c >= 0) ? x[c] : -x[~c];
      /* inner cell: u += g*(dt*(x_o-x_i)/d); outer cell: u -= g*(dt*(x_i-x_o)/d) == same bits */
      const double flux = dt * (xn - xi) / d[e];
      u += g[e] * flux;
    }
    y[i] = u;
  }
}

void orc_rows_coef(const orc_face_op* op, int width, int64_t ld, const int32_t* col,
                   const int64_t* face, int32_t* col_out, double* a, double* diag) {
  const int64_t n = op->n_cells;
  for (int64_t k = 0; k < (int64_t) width * ld; ++k) {
    col_out[k] = ORC_COL_PAD;
    a[k] = 0.0;
  }
  for (int64_t i = 0; i < ld; ++i) diag[i] = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    double dg = op->prefill ? 1.0 : 0.0;
    int kk = 0;
    for (int k = 0; k < width; ++k) {
      const int64_t e = (int64_t) k * ld + i;
      const int32_t c = col[e];
      if (c == ORC_COL_PAD) continue;
      const int64_t f = face[e];
      double area, dist;
      if (f < op->n_faces) {
        area = op->face_area[f];
        dist = op->face_dist[f];
      } else {
        area = op->bface_area[f - op->n_faces];
        dist = op->bface_dist[f - op->n_faces];
      }
      const double coef = ((area / op->cell_vol[i]) * op->dt) / dist;
      if (c >= 0) {
        dg = dg - coef;
        col_out[(int64_t) kk * ld + i] = c;
        a[(int64_t) kk * ld + i] = coef;
        kk++;
      } else {
        dg = dg - (coef + coef);
      }
    }
    diag[i] = dg;
  }
}

void orc_apply_rows_coef(int64_t n, int width, int64_t ld, const int32_t* col, const double* a,
                         const double* diag, const double* x, double* y) {
  for (int64_t i = 0; i < n; ++i) {
    double acc = diag[i] * x[i];
    for (int k = 0; k < width; ++k) {
      const int64_t e = (int64_t) k * ld + i;
      const int32_t c = col[e];
      if (c == ORC_COL_PAD) continue;
      acc = acc + a[e] * x[c];
    }
    y[i] = acc;
  }
}

/* ------------------------------------------------------------------------------------------------
 * Convection-diffusion (config 3): face loop and its row form.
 * ---------------------------------------------------------------------------------------------- */
void orc_apply_convdiff_faces(const orc_convdiff_op* op, const double* x, double* y) {
  const orc_face_op* g = &op->base;
  for (int64_t i = 0; i < g->n_cells; ++i) y[i] = 0.0;
  for (int64_t f = 0; f < g->n_faces; ++f) {
    const int32_t ci = g->face_cell[2 * f + 0], co = g->face_cell[2 * f + 1];
    const double un = op->face_un[f];
    const double up = un > 0.0 ? un : 0.0, um = un < 0.0 ? un : 0.0;
    /* flux = flux_scheme(normal, u[outer], u[inner])   (ConvectionScheme.hpp:87-88) */
    const double flux = (up * x[ci] + um * x[co]) - op->nu * (x[co] - x[ci]) / g->face_dist[f];
    y[ci] += (g->face_area[f] / g->cell_vol[ci]) * flux; /* :89 */
    y[co] -= (g->face_area[f] / g->cell_vol[co]) * flux; /* :90 */
  }
  for (int64_t b = 0; b < g->n_bfaces; ++b) {
    const int32_t ci = g->bface_cell[b];
    const double ghost = -x[ci]; /* bc->get_ghost_state (:98-99): homogeneous Dirichlet mirror */
    const double un = op->bface_un[b];
    const double up = un > 0.0 ? un : 0.0, um = un < 0.0 ? un : 0.0;
    const double flux = (up * x[ci] + um * ghost) - op->nu * (ghost - x[ci]) / g->bface_dist[b];
    y[ci] += (g->bface_area[b] / g->cell_vol[ci]) * flux; /* :103 */
  }
}

void orc_apply_convdiff_faces_cb(void* user, double* y, const double* x, size_t n) {
  (void) n;
  orc_apply_convdiff_faces((const orc_convdiff_op*) user, x, y);
}

void orc_rows_convdiff(const orc_convdiff_op* op, int width, int64_t ld, int32_t* col, double* a, double* diag) {
  const orc_face_op* g = &op->base;
  const int64_t n = g->n_cells;
  for (int64_t k = 0; k < (int64_t) width * ld; ++k) {
    col[k] = ORC_COL_PAD;
    a[k] = 0.0;
  }
  for (int64_t i = 0; i < ld; ++i) diag[i] = 0.0;
  int32_t* fill = (int32_t*) calloc((size_t) n + 1, sizeof(int32_t));
  for (int64_t f = 0; f < g->n_faces; ++f) {
    const int32_t ci = g->face_cell[2 * f + 0], co = g->face_cell[2 * f + 1];
    const double un = op->face_un[f];
    const double up = un > 0.0 ? un : 0.0, um = un < 0.0 ? un : 0.0;
    const double kd = op->nu / g->face_dist[f];
    const double gi = g->face_area[f] / g->cell_vol[ci], go = g->face_area[f] / g->cell_vol[co];
    int64_t e = (int64_t) fill[ci] * ld + ci;
    col[e] = co;
    a[e] = gi * (um - kd);
    diag[ci] = diag[ci] + gi * (up + kd);
    fill[ci]++;
    e = (int64_t) fill[co] * ld + co;
    col[e] = ci;
    a[e] = go * ((-up) - kd);
    diag[co] = diag[co] + go * (kd - um);
    fill[co]++;
  }
  for (int64_t b = 0; b < g->n_bfaces; ++b) {
    const int32_t ci = g->bface_cell[b];
    const double un = op->bface_un[b];
    const double up = un > 0.0 ? un : 0.0, um = un < 0.0 ? un : 0.0;
    const double kd = op->nu / g->bface_dist[b];
    const double gc = g->bface_area[b] / g->cell_vol[ci];
    diag[ci] = diag[ci] + gc * ((up - um) + (kd + kd));
  }
  free(fill);
}

void orc_apply_rows_cb(void* user, double* y, const double* x, size_t n) {
  const orc_rows_op* r = (const orc_rows_op*) user;
  (void) n;
  orc_apply_rows_coef(r->n, r->width, r->ld, r->col, r->a, r->diag, x, y);
}

/* ------------------------------------------------------------------------------------------------
 * Reductions.
 * ORC_RED_SEQ: Bittern `reduce` (MatrixAlgorithms.hpp:191-205): init = Result{} = 0.0, then
 *              init = init + a_i*b_i for ascending i.
 * ORC_RED_TREE: "SB_TREE v1", the product's fixed-shape GPU reduction tree:
 *   element e belongs to CTA tile c = e / 2048, warp w = (e % 2048) / 256, sub-iteration
 *   j = (e % 256) / 64, lane l = (e % 64) / 2. Lane accumulates its 8 products sequentially
 *   from +0.0 (j ascending, even element first); lanes combine with an xor butterfly (16,8,4,2,1);
 *   the 8 warp sums are added left to right -> partial[c]. Final stage: one CTA of 256 threads,
 *   thread t sums partial[t], partial[t+256], ... sequentially from +0.0; butterfly inside each
 *   warp; the 8 warp sums are added left to right. Out-of-range elements contribute +0.0.
 * ---------------------------------------------------------------------------------------------- */
static void butterfly32(double v[32]) {
  for (int m = 16; m >= 1; m >>= 1) {
    double t[32];
    for (int l = 0; l < 32; ++l) t[l] = v[l] + v[l ^ m];
    memcpy(v, t, sizeof(t));
  }
}

static double tree_dot(int64_t n, const double* a, const double* b) {
  const int64_t n_tiles = (n + 2047) / 2048;
  double* partial = (double*) malloc(sizeof(double) * (size_t) (n_tiles > 0 ? n_tiles : 1));
  for (int64_t c = 0; c < n_tiles; ++c) {
    double ws[8];
    for (int w = 0; w < 8; ++w) {
      double lane[32];
      for (int l = 0; l < 32; ++l) {
        double acc = 0.0;
        for (int j = 0; j < 4; ++j) {
          const int64_t e0 = c * 2048 + w * 256 + j * 64 + 2 * l;
          const double p0 = (e0 < n) ? a[e0] * b[e0] : 0.0;
          const double p1 = (e0 + 1 < n) ? a[e0 + 1] * b[e0 + 1] : 0.0;
          acc = acc + p0;
          acc = acc + p1;
        }
        lane[l] = acc;
      }
      butterfly32(lane);
      ws[w] = lane[0];
    }
    double s = ws[0];
    for (int w = 1; w < 8; ++w) s = s + ws[w];
    partial[c] = s;
  }
  /* final stage: one 256-thread CTA */
  double wsum[8];
  for (int w = 0; w < 8; ++w) {
    double lane[32];
    for (int l = 0; l < 32; ++l) {
      const int t = w * 32 + l;
      double s = 0.0;
      for (int64_t q = t; q < n_tiles; q += 256) s = s + partial[q];
      lane[l] = s;
    }
    butterfly32(lane);
    wsum[w] = lane[0];
  }
  double total = wsum[0];
  for (int w = 1; w < 8; ++w) total = total + wsum[w];
  free(partial);
  return total;
}

/* ORC_RED_TREE_SEG: the multi-GPU reduction. The vector is the concatenation of the ranks' owned
 * blocks (segment r = rank r's cells in its local order); each rank reduces its block with SB_TREE v1
 * (tiles restart at the block start) and the rank sums are added in rank order 0..P-1
 * (sb_comm.cuh: allreduce_p2p). */
static int g_n_seg = 0;
static int64_t g_seg_ptr[65];

void orc_set_segments(int n_seg, const int64_t* seg_ptr) {
  g_n_seg = n_seg < 64 ? n_seg : 64;
  for (int k = 0; k <= g_n_seg; ++k) g_seg_ptr[k] = seg_ptr[k];
}

static double seg_tree_dot(int64_t n, const double* a, const double* b) {
  if (g_n_seg <= 0 || g_seg_ptr[g_n_seg] != n) return tree_dot(n, a, b);
  double total = tree_dot(g_seg_ptr[1] - g_seg_ptr[0], a + g_seg_ptr[0], b + g_seg_ptr[0]);
  for (int r = 1; r < g_n_seg; ++r)
    total = total + tree_dot(g_seg_ptr[r + 1] - g_seg_ptr[r], a + g_seg_ptr[r], b + g_seg_ptr[r]);
  return total;
}

double orc_dot(int64_t n, const double* a, const double* b, int mode) {
  if (mode == ORC_RED_TREE) return tree_dot(n, a, b);
  if (mode == ORC_RED_TREE_SEG) return seg_tree_dot(n, a, b);
  double init = 0.0;
  for (int64_t i = 0; i < n; ++i) init = init + a[i] * b[i];
  return init;
}

double orc_norm2(int64_t n, const double* a, int mode) {
  /* norm_2 = sqrt(reduce(0, Add, AbsSquared)) (MatrixAlgorithms.hpp:262-270); no scaling. */
  return sqrt(orc_dot(n, a, a, mode));
}

double orc_safe_divide(double x, double y) {
  /* Crow/MathUtils.hpp:49-52: zero when the divisor is exactly zero. */
  return (y == 0.0) ? 0.0 : (x / y);
}

/* ------------------------------------------------------------------------------------------------
 * Solver driver (Solver.hpp:116-147) shared by the restated solvers.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  orc_apply_fn apply;
  void* user;
  int64_t n;
  int mode;
  double* trace;
  int64_t trace_cap, n_trace;
} orc_env;

static double env_dot(orc_env* E, const double* a, const double* b) {
  const double v = orc_dot(E->n, a, b, E->mode);
  if (E->trace && E->n_trace < E->trace_cap) E->trace[E->n_trace] = v;
  E->n_trace++;
  return v;
}

static double env_norm2(orc_env* E, const double* a) {
  const double v = orc_norm2(E->n, a, E->mode);
  if (E->trace && E->n_trace < E->trace_cap) E->trace[E->n_trace] = v;
  E->n_trace++;
  return v;
}

/* Operator::Residual (Operator.hpp:95-99): mul(r, x); r <<= b - r. */
static void env_residual(orc_env* E, double* r, const double* b, const double* x) {
  E->apply(E->user, r, x, (size_t) E->n);
  for (int64_t i = 0; i < E->n; ++i) r[i] = b[i] - r[i];
}

typedef double (*iterate_fn)(orc_env* E, void* state, int64_t iteration, double* x, const double* b);

static int run_solve(orc_env* E, void* state, double initial_error, iterate_fn iterate, double* x,
                     const double* b, const orc_solver_opts* opts, orc_solver_report* rep,
                     double* hist, int64_t hist_cap) {
  int64_t n_hist = 0;
  double absolute_error = initial_error, relative_error = 0.0;
  if (hist && n_hist < hist_cap) hist[n_hist] = absolute_error;
  n_hist++;
  int converged = 0;
  int64_t iteration = 0;
  if (opts->abs_tol > 0.0 && absolute_error < opts->abs_tol) {
    converged = 1; /* early exit (Solver.hpp:124-128) */
  } else {
    for (iteration = 0; !converged && iteration < opts->num_iterations; ++iteration) {
      absolute_error = iterate(E, state, iteration, x, b);
      relative_error = absolute_error / initial_error; /* no zero guard (g4) */
      if (hist && n_hist < hist_cap) hist[n_hist] = absolute_error;
      n_hist++;
      converged |= (opts->abs_tol > 0.0) && (absolute_error < opts->abs_tol);
      converged |= (opts->rel_tol > 0.0) && (relative_error < opts->rel_tol);
    }
  }
  rep->converged = converged;
  rep->iterations = iteration;
  rep->abs_err = absolute_error;
  rep->rel_err = relative_error;
  rep->n_hist = n_hist;
  rep->n_trace = E->n_trace;
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * CG, no preconditioner (SolverCg.hpp:54-126).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  double gamma;
  double *p, *r, *z;
} cg_state;

static double cg_iterate(orc_env* E, void* st, int64_t iteration, double* x, const double* b) {
  (void) iteration;
  (void) b;
  cg_state* S = (cg_state*) st;
  const int64_t n = E->n;
  E->apply(E->user, S->z, S->p, (size_t) n);                          /* z <- A p       (:95) */
  const double alpha = orc_safe_divide(S->gamma, env_dot(E, S->p, S->z)); /*             (:96) */
  for (int64_t i = 0; i < n; ++i) x[i] += alpha * S->p[i];            /* x += alpha*p   (:97) */
  for (int64_t i = 0; i < n; ++i) S->r[i] -= alpha * S->z[i];         /* r -= alpha*z   (:98) */
  const double gamma_bar = S->gamma;                                  /*                (:109) */
  S->gamma = env_dot(E, S->r, S->r);                                  /*                (:114) */
  const double beta = orc_safe_divide(S->gamma, gamma_bar);           /*                (:121) */
  for (int64_t i = 0; i < n; ++i) S->p[i] = S->r[i] + beta * S->p[i]; /* p <- r + beta*p (:122) */
  return sqrt(S->gamma);                                              /*                (:124) */
}

int orc_cg(orc_apply_fn apply, void* user, int64_t n, const double* b, double* x,
           const orc_solver_opts* opts, orc_solver_report* rep, double* hist, int64_t hist_cap,
           double* trace, int64_t trace_cap) {
  orc_env E = {apply, user, n, opts->reduction_mode, trace, trace_cap, 0};
  cg_state S;
  S.p = (double*) calloc((size_t) n + 1, sizeof(double));
  S.r = (double*) calloc((size_t) n + 1, sizeof(double));
  S.z = (double*) calloc((size_t) n + 1, sizeof(double));
  env_residual(&E, S.r, b, x);                        /* (:73) */
  memcpy(S.p, S.r, sizeof(double) * (size_t) n);      /* p <- r (:79) */
  S.gamma = env_dot(&E, S.r, S.r);                    /* (:80) */
  const int rc = run_solve(&E, &S, sqrt(S.gamma), cg_iterate, x, b, opts, rep, hist, hist_cap);
  free(S.p);
  free(S.r);
  free(S.z);
  return rc;
}

/* ------------------------------------------------------------------------------------------------
 * BiCGStab, no preconditioner (SolverBiCgStab.hpp:59-165).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  double alpha, rho, omega;
  double *p, *r, *r_tilde, *t, *v;
} bicgstab_state;

static double bicgstab_iterate(orc_env* E, void* st, int64_t iteration, double* x, const double* b) {
  (void) b;
  bicgstab_state* S = (bicgstab_state*) st;
  const int64_t n = E->n;
  if (iteration == 0) { /* global iteration counter (g2) (:111-113) */
    memcpy(S->p, S->r, sizeof(double) * (size_t) n);
  } else {
    const double rho_bar = S->rho;                                     /* (:115-116) */
    S->rho = env_dot(E, S->r_tilde, S->r);
    const double beta = orc_safe_divide(S->alpha * S->rho, S->omega * rho_bar); /* (:117) */
    for (int64_t i = 0; i < n; ++i)                                    /* p <- r + beta*(p - omega*v) (:118) */
      S->p[i] = S->r[i] + beta * (S->p[i] - S->omega * S->v[i]);
  }
  E->apply(E->user, S->v, S->p, (size_t) n);                           /* v <- A p (:137) */
  S->alpha = orc_safe_divide(S->rho, env_dot(E, S->r_tilde, S->v));    /* (:139) */
  for (int64_t i = 0; i < n; ++i) x[i] += S->alpha * S->p[i];          /* (:140) */
  for (int64_t i = 0; i < n; ++i) S->r[i] -= S->alpha * S->v[i];       /* (:141) */
  E->apply(E->user, S->t, S->r, (size_t) n);                           /* t <- A r (:158) */
  {
    /* (:159-160). The two dot_product calls are function arguments of safe_divide; g++ on
     * x86-64 evaluates them right to left, so <t.t> is computed (and traced) before <t.r>.
     * The values do not depend on the order; only the trace order does. */
    const double tt = env_dot(E, S->t, S->t);
    const double tr = env_dot(E, S->t, S->r);
    S->omega = orc_safe_divide(tr, tt);
  }
  for (int64_t i = 0; i < n; ++i) x[i] += S->omega * S->r[i];          /* (:161) */
  for (int64_t i = 0; i < n; ++i) S->r[i] -= S->omega * S->t[i];       /* (:162) */
  return env_norm2(E, S->r);                                           /* (:164) */
}

int orc_bicgstab(orc_apply_fn apply, void* user, int64_t n, const double* b, double* x,
                 const orc_solver_opts* opts, orc_solver_report* rep, double* hist, int64_t hist_cap,
                 double* trace, int64_t trace_cap) {
  orc_env E = {apply, user, n, opts->reduction_mode, trace, trace_cap, 0};
  bicgstab_state S;
  S.alpha = S.rho = S.omega = 0.0;
  S.p = (double*) calloc((size_t) n + 1, sizeof(double));
  S.r = (double*) calloc((size_t) n + 1, sizeof(double));
  S.r_tilde = (double*) calloc((size_t) n + 1, sizeof(double));
  S.t = (double*) calloc((size_t) n + 1, sizeof(double));
  S.v = (double*) calloc((size_t) n + 1, sizeof(double));
  env_residual(&E, S.r, b, x);                             /* (:83) */
  memcpy(S.r_tilde, S.r, sizeof(double) * (size_t) n);     /* (:88) */
  S.rho = env_dot(&E, S.r_tilde, S.r);                     /* (:89) */
  const int rc =
      run_solve(&E, &S, sqrt(S.rho), bicgstab_iterate, x, b, opts, rep, hist, hist_cap);
  free(S.p);
  free(S.r);
  free(S.r_tilde);
  free(S.t);
  free(S.v);
  return rc;
}
