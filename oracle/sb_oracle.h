/*
 * sb_oracle.h -- TEST INFRASTRUCTURE. CPU restatement of the StormRuler Krylov hot path.
 *
 * This is the parity oracle, NOT product code: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference leg may load it. The product
 * (stormruler_b200/) never links or calls anything declared here.
 *
 * Parity status: PINNED. Every function below is checked in tests/ against the
 * reference's own headers compiled verbatim (oracle/_ref, built by
 * oracle/Makefile from /root/reference) and against the golden vectors those
 * builds produced (tests/golden/, generator: tests/golden/make_golden.py).
 *
 * Reference anchors (paths relative to /root/reference):
 *   face-loop operator      source_apps/playground/Playground.cpp:115-131
 *   boundary-face pattern   source/Storm/Feathers/ConvectionScheme.hpp:95-106
 *   inner/outer convention  source/Storm/Mallard/Mesh.hpp:269-280,
 *                           source/Storm/Mallard/MeshUnstructured.hpp:509-554
 *   dot_product / norm_2    source/Storm/Bittern/MatrixAlgorithms.hpp:162-205,262-270,310-317
 *   safe_divide             source/Storm/Crow/MathUtils.hpp:49-52
 *   solve loop              source/Storm/Solvers/Solver.hpp:116-147
 *   CG                      source/Storm/Solvers/SolverCg.hpp:54-126
 *   BiCGStab                source/Storm/Solvers/SolverBiCgStab.hpp:59-165
 *   Residual                source/Storm/Solvers/Operator.hpp:95-99
 *
 * Build flags are part of the contract: -O2 -ffp-contract=off, no
 * -march, no -ffast-math: every a*b+c is rounded twice, sums are sequential.
 */
#ifndef SB_ORACLE_H
#define SB_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Reduction order used by dot/norm. */
enum {
  ORC_RED_SEQ = 0,  /* the reference's order: left-to-right from 0.0 (MatrixAlgorithms.hpp:191-205) */
  ORC_RED_TREE = 1, /* the GPU's fixed tree ("SB_TREE v1"), restated on the CPU */
  ORC_RED_TREE_SEG = 2 /* multi-GPU: SB_TREE v1 per rank block, rank sums added in rank order */
};
/* Segment boundaries for ORC_RED_TREE_SEG: seg_ptr[0..n_seg], seg_ptr[n_seg] = vector length. */
void orc_set_segments(int n_seg, const int64_t* seg_ptr);

/* Padding / ghost encodings of the ELL column array (shared with the product's documented layout). */
#define ORC_COL_PAD INT32_MIN
/* boundary (Dirichlet mirror ghost) entry of cell i: col = ~i  (negative, != ORC_COL_PAD) */

/* Face-list description of the FVM operator
 *   y = prefill(x) ; stormDivGrad(mesh, y, dt, x) ; [boundary ghost loop]
 * prefill: 0 -> y = 0, 1 -> y = x  (Playground.cpp:153-167 uses `c_hat <<= c_in` before the face loop). */
typedef struct {
  int64_t n_cells;
  int64_t n_faces;           /* interior faces (label 0), in reference face order */
  const int32_t* face_cell;  /* [2*n_faces]: inner, outer */
  const double* face_area;   /* [n_faces] */
  const double* face_dist;   /* [n_faces]  ||x_outer - x_inner|| */
  const double* cell_vol;    /* [n_cells] */
  int64_t n_bfaces;          /* boundary faces with a homogeneous-Dirichlet mirror ghost (0 = pure Neumann) */
  const int32_t* bface_cell; /* [n_bfaces] inner cell */
  const double* bface_area;  /* [n_bfaces] */
  const double* bface_dist;  /* [n_bfaces] ||x_ghost - x_inner|| = 2 * distance(cell centre, face centre) */
  int32_t prefill;           /* 0 or 1 */
  double dt;                 /* the `dt` argument of stormDivGrad, verbatim (negative for x - |dt| lap x) */
} orc_face_op;

/* y <- A(x), face loop in ascending face index; boundary faces after all interior faces. */
void orc_apply_faces(const orc_face_op* op, const double* x, double* y);

/* u += dt * div grad c over the same faces: stormDivGrad exactly as the playground calls it (the caller pre-fills
 * u; Playground.cpp:115-131, call sites :159,:165). Ignores op->prefill / op->dt. */
void orc_divgrad_accumulate(const orc_face_op* op, double dt, const double* c, double* u);

/* Same signature as the callback taken by oracle/_ref's ref_solve(): user = const orc_face_op*. */
void orc_apply_faces_cb(void* user, double* y, const double* x, size_t n);

/* Convection-diffusion operator  y = -nu div grad x + div(beta x), first-order upwind, in the face-loop
 * pattern of UpwindConvectionScheme::operator() (Feathers/ConvectionScheme.hpp:83-106): interior faces
 * add the flux to the inner cell and subtract it from the outer cell; boundary faces (ghost state from the
 * boundary condition, here the homogeneous-Dirichlet mirror -x[c]) add it to the inner cell only.
 * The reference's flux scheme is a functor argument (ConvectionScheme.hpp:64,87); for the scalar linear
 * problem of SURVEY.md 8d config 3 it is  flux(n, u_outer, u_inner) = max(un,0)*u_inner + min(un,0)*u_outer
 * - nu*(u_outer - u_inner)/dist,  un = beta . n.  face_un / bface_un hold un per face. */
typedef struct {
  orc_face_op base;       /* geometry (prefill and dt unused) */
  double nu;
  const double* face_un;  /* [n_faces] */
  const double* bface_un; /* [n_bfaces] */
} orc_convdiff_op;
void orc_apply_convdiff_faces(const orc_convdiff_op* op, const double* x, double* y);
void orc_apply_convdiff_faces_cb(void* user, double* y, const double* x, size_t n);
/* Row (coefficient) form of the same operator, entry order from orc_build_rows(&op->base) with
 * base.n_bfaces = 0 for the structure and the boundary faces folded into the diagonal. Operation order
 * is part of the layout contract (include/stormb200.h: sb_convdiff_desc). */
void orc_rows_convdiff(const orc_convdiff_op* op, int width, int64_t ld, int32_t* col, double* a, double* diag);
/* orc_apply_rows_coef with a context: the callback form used by the solvers. */
typedef struct {
  int64_t n;
  int width;
  int64_t ld;
  const int32_t* col;
  const double* a;
  const double* diag;
} orc_rows_op;
void orc_apply_rows_cb(void* user, double* y, const double* x, size_t n);

/* Cell-row (ELL) form of the same operator.
 * width = max entries per row; ld = leading dimension (>= n_cells); entry k of row i at [k*ld + i].
 * Entries of a row are ordered exactly as the face loop visits the cell: interior faces by
 * ascending face index, then boundary faces by ascending index.
 *   col  : neighbour cell, or ~i for a Dirichlet ghost, or ORC_COL_PAD
 *   face : face index (interior: f, boundary: n_faces + b), -1 for padding */
int orc_rows_width(const orc_face_op* op);
void orc_build_rows(const orc_face_op* op, int width, int64_t ld, int32_t* col, int64_t* face);

/* "faithful" per-entry data: g = area/vol_i, d = dist, evaluated as in Playground.cpp:126-129. */
void orc_rows_faithful(const orc_face_op* op, int width, int64_t ld, const int64_t* face, double* g,
                       double* d);
/* y_i = prefill(x_i) (prefill 2: the old y_i); for k: y_i += g_k * (dt * (xn_k - x_i) / d_k), xn_k = x[col] or -x[i] for a ghost.
 * Bit-identical to orc_apply_faces (u - g*F == u + g*(-F) exactly in IEEE-754). */
void orc_apply_rows_faithful(int64_t n, int width, int64_t ld, const int32_t* col, const double* g,
                             const double* d, int32_t prefill, double dt, const double* x, double* y);

/* "coefficient" per-entry data: a = ((area/vol_i)*dt)/dist; diag = prefill - sum(a) - sum_ghost(a+a),
 * accumulated in row order. Ghost entries fold into diag and are dropped (col_out has no ~i). */
void orc_rows_coef(const orc_face_op* op, int width, int64_t ld, const int32_t* col,
                   const int64_t* face, int32_t* col_out, double* a, double* diag);
/* y_i = diag_i*x_i; for k: y_i += a_k * x[col_k]  (mul and add rounded separately). */
void orc_apply_rows_coef(int64_t n, int width, int64_t ld, const int32_t* col, const double* a,
                         const double* diag, const double* x, double* y);

/* BLAS-1 reductions. */
double orc_dot(int64_t n, const double* a, const double* b, int mode);
double orc_norm2(int64_t n, const double* a, int mode);
double orc_safe_divide(double x, double y);

/* Generic operator handle for the solver restatements. */
typedef void (*orc_apply_fn)(void* user, double* y, const double* x, size_t n);

typedef struct {
  int64_t num_iterations;  /* Solver.hpp:67, default 2000 */
  double abs_tol;          /* Solver.hpp:71, default 1e-6; <= 0 disables */
  double rel_tol;          /* Solver.hpp:72, default 1e-6; <= 0 disables */
  int32_t reduction_mode;  /* ORC_RED_* */
} orc_solver_opts;

typedef struct {
  int32_t converged;
  int64_t iterations;   /* value of IterativeSolver::iteration after solve() */
  double abs_err;       /* IterativeSolver::absolute_error */
  double rel_err;       /* IterativeSolver::relative_error */
  int64_t n_hist;       /* hist[0] = initial residual, hist[k] = value returned by iterate() #k */
  int64_t n_trace;      /* every dot_product / norm_2 result in call order */
} orc_solver_report;

int orc_cg(orc_apply_fn apply, void* user, int64_t n, const double* b, double* x,
           const orc_solver_opts* opts, orc_solver_report* rep, double* hist, int64_t hist_cap,
           double* trace, int64_t trace_cap);
int orc_bicgstab(orc_apply_fn apply, void* user, int64_t n, const double* b, double* x,
                 const orc_solver_opts* opts, orc_solver_report* rep, double* hist, int64_t hist_cap,
                 double* trace, int64_t trace_cap);

#ifdef __cplusplus
}
#endif
#endif /* SB_ORACLE_H */
