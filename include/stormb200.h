/*
 * stormb200.h -- C ABI of the B200-native Krylov hot path for StormRuler.
 *
 * This is the drop-in boundary. The reference (Jhuighuy/StormRuler) is a header-only C++
 * template library with no FFI of its own; its extension points for this path are
 *   - the `legacy_vector_like` vector concept          source/Storm/Solvers/Operator.hpp:39-45
 *   - `Operator<Vector>::mul` (virtual)                source/Storm/Solvers/Operator.hpp:66-74
 *   - the Bittern free functions / operators found by overload resolution on the vector type
 *                                                      source/Storm/Bittern/MatrixAlgorithms.hpp,
 *                                                      MatrixMath.hpp, MatrixTarget.hpp
 * The C++23 host header stormruler_b200/host/Storm/B200/DeviceVector.hpp implements those on top
 * of the functions below; g++ compiles it together with the unmodified reference solver headers,
 * nvcc compiles this library for sm_100a (nvcc 12.9 cannot parse C++23, hence the C boundary).
 * INTEGRATION.md shows the binding a StormRuler maintainer would add.
 *
 * Conventions: opaque handles; every function returns an int status (0 = SB_OK, negative =
 * error) and never throws; sb_last_error() returns a thread-local message for the last failure.
 * Pointers are DEVICE pointers unless their name starts with h_. All device vectors must come
 * from sb_vec_alloc (it pads the capacity so kernels can use unguarded 128-bit accesses).
 * One context = one device = one host thread at a time (the reference is single-threaded,
 * SURVEY.md 8b). There is no CPU fallback: without a CUDA device every entry point that touches
 * the device fails with SB_ERR_CUDA.
 */
#ifndef STORMB200_H
#define STORMB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB_API __attribute__((visibility("default")))

#define SB_OK 0
#define SB_ERR_INVALID (-1) /* bad argument */
#define SB_ERR_CUDA (-2)    /* CUDA runtime failure (sticky: the context is unusable afterwards) */
#define SB_ERR_NOMEM (-3)
#define SB_ERR_NCCL (-4)
#define SB_ERR_STATE (-5)   /* call not valid in this state (e.g. comm not initialised) */
#define SB_ERR_COMM (-6)    /* a device-side wait for a peer (or for the grid) gave up after SB_SPIN_TIMEOUT_S
                               seconds (default 120): results are meaningless, the context is unusable */

typedef struct sb_ctx sb_ctx;
typedef struct sb_op sb_op;

SB_API const char* sb_last_error(void);
SB_API int sb_version(void);

/* ---- context ---------------------------------------------------------------------------------
 * Owns the device, one compute stream, reduction scratch and pinned staging memory. */
SB_API int sb_ctx_create(int device, sb_ctx** out);
SB_API int sb_ctx_destroy(sb_ctx* ctx);
SB_API int sb_sync(sb_ctx* ctx);
/* cudaStream_t of the compute stream (for CUDA-event timing by the caller). */
SB_API void* sb_ctx_stream(sb_ctx* ctx);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
SB_API int64_t sb_ctx_launch_count(sb_ctx* ctx);

/* ---- vectors: replaces the storage of Feathers/Field.hpp:60-114 (CellField, the solvers' Vector)
 * sb_vec_alloc zero-fills, like Field::assign (Field.hpp:82-84) which IDR(s) relies on. */
SB_API int sb_vec_alloc(sb_ctx* ctx, size_t n, double** d_out);
SB_API int sb_vec_free(sb_ctx* ctx, double* d);
SB_API int sb_vec_upload(sb_ctx* ctx, double* d, const double* h_src, size_t n);
SB_API int sb_vec_download(sb_ctx* ctx, const double* d, double* h_dst, size_t n);

/* ---- operator: replaces the face loop of source_apps/playground/Playground.cpp:115-131 driven
 * through FunctionalOperator::mul (Operator.hpp:156-158), plus the boundary-ghost pattern of
 * Feathers/ConvectionScheme.hpp:95-106.
 *
 *   y = prefill(x);  for interior faces f (ascending): flux = dt*(x[outer]-x[inner])/dist_f;
 *                    y[inner] += (area_f/vol_inner)*flux;  y[outer] -= (area_f/vol_outer)*flux;
 *                    for Dirichlet boundary faces b (ascending, after all interior faces):
 *                    flux = dt*((-x[c]) - x[c])/bdist_b;  y[c] += (barea_b/vol_c)*flux.
 *
 * The face list is uploaded once and turned into a cell-row (ELL) layout on the device side:
 * rows ordered by cell, entries of a row in ascending face index (the CPU's summation order). */
typedef struct sb_mesh_soa {
  int64_t n_cells;
  int64_t n_faces;           /* interior faces, reference face order */
  const int32_t* face_cell;  /* h_ [2*n_faces] inner, outer  (Mesh.hpp:269-280) */
  const double* face_area;   /* h_ [n_faces] */
  const double* face_dist;   /* h_ [n_faces] ||centre_outer - centre_inner|| */
  const double* cell_vol;    /* h_ [n_cells] */
  int64_t n_bfaces;          /* Dirichlet (mirror ghost) boundary faces; 0 = pure Neumann */
  const int32_t* bface_cell; /* h_ [n_bfaces] */
  const double* bface_area;  /* h_ [n_bfaces] */
  const double* bface_dist;  /* h_ [n_bfaces] ||ghost centre - cell centre|| */
} sb_mesh_soa;

#define SB_FORM_FAITHFUL 0 /* per entry (col, area/vol, dist): bit-identical to the face loop */
#define SB_FORM_COEF 1     /* per entry (col, coef) + diagonal: the 12 B/entry streaming form  */

typedef struct sb_op_desc {
  int32_t form;    /* SB_FORM_* */
  int32_t prefill; /* 0: y starts from 0; 1: y starts from x (`c_hat <<= c_in`, Playground.cpp:162) */
  double dt;       /* the dt argument of stormDivGrad, verbatim */
} sb_op_desc;

typedef struct sb_op_info {
  int64_t n_cells;
  int64_t n_entries;      /* stored off-diagonal entries (ghosts included in the faithful form) */
  int32_t width;          /* ELL width */
  int64_t ld;             /* leading dimension (padded rows) */
  int32_t form;
  int64_t device_bytes;   /* bytes of operator data resident in HBM */
  int64_t algorithmic_bytes_per_apply; /* 24*N + 12*entries (coef form) */
} sb_op_info;

SB_API int sb_op_create(sb_ctx* ctx, const sb_mesh_soa* h_mesh, const sb_op_desc* desc, sb_op** out);
SB_API int sb_op_destroy(sb_ctx* ctx, sb_op* op);
SB_API int sb_op_get_info(const sb_op* op, sb_op_info* info);
/* Convection-diffusion operator  y = -nu div grad x + div(beta x)  with the first-order upwind face
 * flux, written in the inner/outer face-loop pattern of UpwindConvectionScheme::operator()
 * (Feathers/ConvectionScheme.hpp:83-106) -- SURVEY.md 8d config 3:
 *   interior face f (inner i, outer o, un = beta . n_f, n_f pointing i -> o):
 *       flux = max(un,0)*x[i] + min(un,0)*x[o] - nu*(x[o]-x[i])/dist_f
 *       y[i] += (area_f/vol_i)*flux ;  y[o] -= (area_f/vol_o)*flux
 *   Dirichlet boundary face b of cell c (mirror ghost -x[c], un = beta . n_b outward):
 *       flux = max(un,0)*x[c] + min(un,0)*(-x[c]) - nu*((-x[c])-x[c])/bdist_b ;  y[c] += (barea_b/vol_c)*flux
 * The rows are non-symmetric (two different coefficients per face); they are stored in the coefficient
 * form (SB_FORM_COEF) and applied by the same kernels. Coefficient order of operations (restated in
 * oracle/sb_oracle.c: orc_rows_convdiff): g = area/vol_row, kd = nu/dist, up = max(un,0), um = min(un,0);
 *   inner row: a = g*(um - kd), diag += g*(up + kd);   outer row: a = g*((-up) - kd), diag += g*(kd - um);
 *   boundary:  diag += g*((up - um) + (kd + kd));      diag accumulates in ascending face order from 0. */
typedef struct sb_convdiff_desc {
  double nu;
  const double* face_un;  /* h_ [n_faces]  beta . n_f per interior face (a face-flux field) */
  const double* bface_un; /* h_ [n_bfaces] beta . n_b per boundary face; may be NULL if n_bfaces == 0 */
} sb_convdiff_desc;
SB_API int sb_op_create_convdiff(sb_ctx* ctx, const sb_mesh_soa* h_mesh, const sb_convdiff_desc* desc, sb_op** out);

/* Copy the row layout back to the host for bit-exact comparison with the oracle:
 * h_col [width*ld] int32, h_val0 [width*ld] (coef, or area/vol), h_val1 [width*ld] (dist; faithful
 * form only, may be NULL), h_diag [ld] (coef form only, may be NULL). */
SB_API int sb_op_download_rows(sb_ctx* ctx, const sb_op* op, int32_t* h_col, double* h_val0,
                               double* h_val1, double* h_diag);
/* y <- A(x): Operator::mul (Operator.hpp:74). x and y must not alias. */
SB_API int sb_apply(sb_ctx* ctx, const sb_op* op, const double* x, double* y);
/* y <- A(x) and <u, y> in one kernel (the dot rides on the apply: its operands are in registers, so it costs no pass
 * over y; SB_TREE, result on the host): what `lin_op.mul(z, p); dot_product(p, z)` (SolverCg.hpp:96-97) or
 * `lin_op.mul(v, p); dot_product(r_tilde, v)` (SolverBiCgStab.hpp:137-139) cost when issued together. u == NULL or
 * u == x: <x, y>. u must not alias y; x and y must not alias. Same kernels as the fused solvers' apply + dot. */
SB_API int sb_apply_dot(sb_ctx* ctx, const sb_op* op, const double* x, double* y, const double* u, double* h_out);
/* y <- A(x) with h_out[0] = <y, y> and h_out[1] = <y, x> from the same kernel: `lin_op.mul(t, r);
 * omega = safe_divide(dot_product(t, r), dot_product(t, t))` (SolverBiCgStab.hpp:158-160; IDR(s)'s
 * `<v,r>/<v,v>`, SolverIdrs.hpp:276-277). x and y must not alias. */
SB_API int sb_apply_dot_yy_yx(sb_ctx* ctx, const sb_op* op, const double* x, double* y, double* h_out);
/* y += dt * div grad x: `stormDivGrad(mesh, u, dt, c)` exactly as the playground calls it (Playground.cpp:115-131;
 * call sites :159 `stormDivGrad(mesh, w_hat, -Gamma, c_in)` after `w_hat <<= f + sigma*(c_in - c)`, and :165
 * `stormDivGrad(mesh, c_hat, -tau, w_hat)` after `c_hat <<= c_in`). Every row starts from the old y_i and adds its
 * face terms in ascending face index, so the result is bit-identical to the face loop accumulating into a
 * pre-filled field. `op` must be a faithful-form operator (SB_FORM_FAITHFUL; its own dt and prefill are not
 * used); boundary faces of the mesh contribute their mirror-ghost term as in sb_apply. x and y must not alias. */
SB_API int sb_apply_accumulate(sb_ctx* ctx, const sb_op* op, double dt, const double* x, double* y);

/* Jacobi (point-diagonal) preconditioner apply: y_i = x_i / a_ii, a_ii the operator's own diagonal (coefficient
 * form only; the faithful form keeps no diagonal). Fills the reference's preconditioner slot
 * (Solvers/Preconditioner.hpp:66-82, `Preconditioner<Vector>::mul`), which every reference solver already
 * branches on (e.g. SolverBiCgStab.hpp:135-137,156-158; SolverGmres.hpp:149-156). The reference ships the
 * identity preconditioner only (README: all others "planned"): SURVEY.md 8f rank 2. x may alias y. */
SB_API int sb_op_jacobi(sb_ctx* ctx, const sb_op* op, const double* x, double* y);

/* ---- mesh ingestion (host side, no GPU needed): replaces what the playground gets from
 * read_mesh_from_tetgen + UnstructuredMesh (Mallard/IoTetgen.hpp:44-235, MeshUnstructured.hpp:350-425,
 * 464-500, 509-554) for the hot path, in 3-D (the reference mesh layer is 2-D only):
 * a cell soup is turned into the face-list SoA above. Conventions follow the reference:
 *   - local faces of a tetrahedron / hexahedron as in Mallard/Shape.hpp:559-606 / 803-843;
 *   - a face is created by the first cell (in cell order, local-face order) that touches it; that
 *     cell is its inner cell, the second one its outer cell (MeshUnstructured.hpp:509-554);
 *   - faces are ordered by creation, interior faces (label 0) first (MeshUnstructured.hpp:464-500);
 *   - triangle area = |cross|/2, quadrangle = two triangles (n1,n2,n3)+(n3,n4,n1), cell centre =
 *     barycentre (simple shapes: node mean; hexahedron: volume-weighted over its 5 tetrahedral
 *     pieces), Shape.hpp:170-215,310-334,395-402,845-852; tetrahedron volume = |triple product|/6.
 * All integer results (face order, inner/outer, permutations, partitions, halo maps) are
 * deterministic and checked bit for bit against the independent C restatement in oracle/. */
typedef struct sb_mesh sb_mesh;
#define SB_CELL_TET 0
#define SB_CELL_HEX 1
#define SB_CELL_FACELIST 2 /* any cell shape: the mesh is its face list (sb_mesh_from_faces) */

/* Box [0,1]^3 of nx*ny*nz hexahedra, optionally split into 6 Kuhn tetrahedra each (SB_CELL_TET).
 * Interior nodes are displaced by U(-jitter*h, jitter*h) per coordinate (std::mt19937_64(seed_jitter),
 * node order, x then y then z); if shuffle != 0 the cell order is randomly permuted
 * (Fisher-Yates, std::mt19937_64(seed_shuffle)) so that renumbering has something to do. */
SB_API int sb_mesh_generate_box(int cell_kind, int nx, int ny, int nz, double jitter, uint64_t seed_jitter,
                                int shuffle, uint64_t seed_shuffle, sb_mesh** out);
/* General ingestion: h_xyz [3*n_nodes], h_cell_nodes [n_cells * (4 | 8)]. */
SB_API int sb_mesh_from_cells(int cell_kind, int64_t n_nodes, const double* h_xyz, int64_t n_cells,
                              const int32_t* h_cell_nodes, sb_mesh** out);
/* Ingestion of a face list -- what the reference's own mesh classes export through interior_faces() /
 * faces(label) (Mallard/Mesh.hpp:426-429,453-455), FaceView::inner_cell/outer_cell (:269-280) and the geometry
 * getters area/volume/center (:254-261,304-311): any cell shape, up to 127 faces per cell (polyhedral meshes,
 * the reference's 2-D meshes). The arrays are copied and kept verbatim (face order and inner/outer as given), so
 * the handle can be renumbered (RCM, sb_mesh_permute_cells) and partitioned like a node-based mesh. A renumbering
 * re-derives the face order by the creation rule above (creating cell = lower cell id, then the face's ordinal
 * among that cell's faces in the given list). Optional: h_cell_ctr [3*n_cells]; unit normals h_face_normal
 * [3*n_faces] (inner -> outer) and h_bface_normal [3*n_bfaces] (outward) -- both or neither. */
SB_API int sb_mesh_from_faces(const sb_mesh_soa* h_soa, const double* h_cell_ctr, const double* h_face_normal,
                              const double* h_bface_normal, sb_mesh** out);
/* TetGen ingestion in 3-D: reads `<prefix>.node` and `<prefix>.ele` with the file
 * grammar of the reference's reader (Mallard/IoTetgen.hpp:44-235: '#' comments, node header
 * `count dim n_attribs has_labels`, element header `count nodes_per_cell has_attribs`, one entity per line led
 * by its index) and hands the tetrahedra to sb_mesh_from_cells. The reference reader is hard-wired to 2-D
 * at this commit -- its 3-D branch reads `.face` files it cannot build a mesh from -- so the face
 * list follows this library's own creation-order convention, like every other 3-D mesh here. Node indices
 * may start at 0 or 1 (TetGen -z); attributes and labels are skipped. */
SB_API int sb_mesh_read_tetgen(const char* path_prefix, sb_mesh** out);
/* 2-D ingestion of Triangle's `<prefix>.node`, `<prefix>.edge`, `<prefix>.ele` -- the reader the reference actually
 * has (read_mesh_from_tetgen on UnstructuredMesh<2,2>, Mallard/IoTetgen.hpp:44-235; playground: Playground.cpp:248-255)
 * and the files its own tests ship (tests/_data/mesh). Restates the reference's construction: faces numbered in
 * .edge file order (unlisted edges appended with label 0), first inserted cell = inner cell
 * (MeshUnstructured.hpp:350-425,509-554), faces stable-sorted by label with label 0 = interior first (:464-500),
 * geometry operation by operation as Mallard/Shape.hpp computes it. The result is a face-list handle
 * (SB_CELL_FACELIST, with cell centres and normals, z = 0) whose SoA is bit-identical to what the reference's mesh
 * classes export; boundary faces are ordered by label. */
SB_API int sb_mesh_read_tetgen_2d(const char* path_prefix, sb_mesh** out);
/* Labels of the boundary faces in their current order (the reference's face labels, Mesh.hpp:426-429; 1 for every
 * boundary face of a mesh that was not read from labelled files). h_labels [n_bfaces]. */
SB_API int sb_mesh_bface_labels(const sb_mesh* mesh, int32_t* h_labels);
SB_API int sb_mesh_destroy(sb_mesh* mesh);
/* Reverse Cuthill-McKee renumbering of the cells over the face adjacency graph; faces are rebuilt
 * for the new cell order. h_perm (may be NULL) receives perm[new] = old (Utils/Permutations.hpp:77-103). */
SB_API int sb_mesh_renumber_rcm(sb_mesh* mesh, int32_t* h_perm);
/* Apply an arbitrary cell permutation perm[new] = old (bijection check included). */
SB_API int sb_mesh_permute_cells(sb_mesh* mesh, const int32_t* h_perm);
/* Fill `soa` with pointers into the mesh (valid until the mesh is modified or destroyed). Boundary
 * faces: every face with a single adjacent cell; bface_dist = 2*|face centre - cell centre|. */
SB_API int sb_mesh_get_soa(const sb_mesh* mesh, sb_mesh_soa* soa);
SB_API int sb_mesh_cell_centers(const sb_mesh* mesh, double* h_xyz /* [3*n_cells] */);
/* Unit face normals (what FaceView::normal() is to the reference's flux schemes,
 * Feathers/ConvectionScheme.hpp:87,100): h_fn [3*n_faces] oriented inner -> outer, h_bn [3*n_bfaces]
 * oriented out of the domain; either may be NULL. Triangle: cross(v2-v1, v3-v1); quadrangle: cross of
 * the diagonals; normalised, flipped if it points against (outer centre - inner centre). */
SB_API int sb_mesh_face_normals(const sb_mesh* mesh, double* h_fn, double* h_bn);
/* Legacy-VTK (ASCII, unstructured grid) dump of the mesh and of n_fields per-cell scalar fields in the CURRENT cell
 * order, in the file grammar of the playground's save_vtk (source_apps/playground/Playground.cpp:65-109: header
 * lines, 16 significant digits, POINTS / CELLS / CELL_TYPES / CELL_DATA with one SCALARS block per field); 3-D
 * points, cell types VTK_TETRA (10) / VTK_HEXAHEDRON (12). Node-based meshes only. Host side, no GPU needed:
 * download the device vectors first (sb_vec_download). */
SB_API int sb_mesh_write_vtk(const sb_mesh* mesh, const char* path, int n_fields, const char* const* names,
                             const double* const* h_fields /* n_fields arrays of [n_cells] */);
/* Bandwidth of the cell graph, max |inner - outer| over interior faces (renumbering quality). */
SB_API int64_t sb_mesh_bandwidth(const sb_mesh* mesh);

/* ---- partitioning (host side, no GPU needed): new work specified by SURVEY.md 8e -- the reference has
 * no partitioning or halo code. The cell graph (cells, one edge per interior face) is split into
 * n_parts; each rank gets a local mesh in the numbering
 *     [ interior owned | boundary owned | padding to a 2048 multiple | halo (by owner rank, then global id) ]
 * whose face list keeps the GLOBAL ascending face order and inner/outer orientation, so every owned
 * row of the distributed operator is bit-identical to the single-GPU row. All arrays are checked bit
 * for bit against the numpy restatement in oracle/mesh_oracle.py. */
typedef struct sb_part sb_part;
#define SB_PART_METIS 0 /* METIS_PartGraphKway, default options (deterministic) */
#define SB_PART_SLAB 1  /* contiguous blocks of the current cell order (use after RCM): <= 2 neighbours */

typedef struct sb_part_info {
  int32_t n_parts;
  int64_t n_cells;
  int64_t edge_cut;     /* interior faces whose cells belong to different parts */
  int64_t max_owned, min_owned;
  int64_t max_halo;
  int64_t vec_capacity; /* max over ranks of pad2048(n_owned) + pad2048(n_halo): size of one pool vector */
} sb_part_info;

typedef struct sb_local_mesh {
  int32_t rank, n_parts;
  int64_t n_owned;      /* rows of this rank; local ids [0, n_owned) */
  int64_t n_interior;   /* owned cells without a neighbour on another rank: local ids [0, n_interior) */
  int64_t n_halo;       /* local ids [halo_base, halo_base + n_halo) */
  int64_t halo_base;    /* n_owned rounded up to a multiple of 2048 */
  const int32_t* local_to_global; /* h_ [n_owned + n_halo]: owned cells, then halo cells */
  sb_mesh_soa soa;      /* local face list (n_cells = halo_base + n_halo), boundary faces of owned cells */
  const int64_t* face_global;     /* h_ [soa.n_faces] global face index of each local face (ascending) */
  int32_t n_nbr;
  const int32_t* nbr_rank; /* h_ [n_nbr] ascending */
  const int64_t* send_ptr; /* h_ [n_nbr+1] */
  const int32_t* send_idx; /* h_ [send_ptr[n_nbr]] local ids of owned cells, ascending global id per neighbour */
  const int64_t* recv_ptr; /* h_ [n_nbr+1] offsets into the halo block; group k holds cells owned by nbr_rank[k] */
  const int64_t* send_dst; /* h_ [n_nbr] element offset in neighbour k's vectors where my block starts
                              (= its halo_base + its recv_ptr for me) */
  const int64_t* bface_global;    /* h_ [soa.n_bfaces] global boundary-face index of each local one */
} sb_local_mesh;

/* SB_PART_METIS: METIS_PartGraphKway over the cell graph; if it leaves a part empty (tiny or edge-less graphs) the
 * contiguous slabs of SB_PART_SLAB are used instead, so every rank always owns cells. */
SB_API int sb_part_create(const sb_mesh* mesh, int n_parts, int method, sb_part** out);
SB_API int sb_part_from_array(const sb_mesh* mesh, int n_parts, const int32_t* h_part, sb_part** out);
SB_API int sb_part_destroy(sb_part* part);
SB_API int sb_part_get_info(sb_part* part, sb_part_info* info);
SB_API int sb_part_get_array(const sb_part* part, const int32_t** h_part /* [n_cells], owned by part */);
/* Local mesh of one rank (pointers stay valid until sb_part_destroy). */
SB_API int sb_part_local(sb_part* part, int rank, sb_local_mesh* out);

/* ---- multi-GPU communicator: one process per GPU. After sb_comm_prepare the context
 * serves every vector (sb_vec_alloc and the solver workspaces) from a pool of n_vectors blocks of
 * vec_capacity doubles inside one slab, so that vectors sit at the same offset on every rank.
 *   SB_COMM_NCCL: halo exchange = grouped ncclSend/ncclRecv, reductions = ncclAllReduce (the
 *                 torch-bundled or system libnccl.so.2 is loaded at run time);
 *   SB_COMM_P2P:  the slab is shared with CUDA IPC; the pack kernel stores boundary values directly
 *                 into the neighbours' halo tails over NVLink and the one-CTA final-reduce kernel
 *                 exchanges its partial sums with all peers itself (summed in rank order: every rank
 *                 gets bit-identical scalars). No host or library call on the critical path.
 * Rendezvous: every rank calls sb_comm_prepare, the SB_COMM_BLOB_BYTES blobs are all-gathered out of
 * band (bench.py: torch.distributed), every rank calls sb_comm_connect with the world*blob array, and
 * the ranks synchronise once (a barrier) before the first collective call. */
#define SB_COMM_NCCL 0
#define SB_COMM_P2P 1
#define SB_COMM_BLOB_BYTES 256
#define SB_COMM_MAX_RANKS 8
SB_API int sb_comm_prepare(sb_ctx* ctx, int rank, int world, int mode, int64_t vec_capacity, int32_t n_vectors,
                           void* h_blob /* [SB_COMM_BLOB_BYTES] out */);
SB_API int sb_comm_connect(sb_ctx* ctx, const void* h_all_blobs /* [world * SB_COMM_BLOB_BYTES] */);
SB_API int sb_comm_destroy(sb_ctx* ctx);
/* Device-side error word of the communicator (0 = fine); spin loops that give up set it and trap. */
SB_API int sb_comm_status(sb_ctx* ctx, uint64_t* h_error);
/* Distributed operator of this rank: rows for the owned cells of `local`, halo plan on the device.
 * sb_apply / the fused solvers exchange halos inside the call; x must be a vector of this context
 * (its halo tail is written by the neighbours). Vectors have n = local->n_owned logical elements. */
SB_API int sb_dist_op_create(sb_ctx* ctx, const sb_local_mesh* local, const sb_op_desc* desc, sb_op** out);
/* Same for the convection-diffusion operator; desc->face_un / bface_un are indexed by LOCAL face
 * (gather them with local->face_global / bface_global). */
SB_API int sb_dist_op_create_convdiff(sb_ctx* ctx, const sb_local_mesh* local, const sb_convdiff_desc* desc,
                                      sb_op** out);

/* ---- BLAS-1: replaces Bittern's lazy expressions + assignment operators
 * (MatrixMath.hpp:233-301, MatrixTarget.hpp:96-119, MatrixAlgorithms.hpp:95-135) for the vector.
 * An expression is a postfix program over <= SB_EXPR_MAX_VEC vector operands and
 * <= SB_EXPR_MAX_SCAL scalars, evaluated per element in exactly the order written (each
 * operation rounded separately, no FMA contraction), e.g. r + beta*(p - omega*v) is
 *   V0 S0 V1 S1 V2 MUL SUB MUL ADD.   The target may alias any operand. */
#define SB_EXPR_MAX_OPS 24
#define SB_EXPR_MAX_VEC 4
#define SB_EXPR_MAX_SCAL 4
enum {
  SB_OP_VEC0 = 0, SB_OP_VEC1, SB_OP_VEC2, SB_OP_VEC3,
  SB_OP_SCAL0 = 8, SB_OP_SCAL1, SB_OP_SCAL2, SB_OP_SCAL3,
  SB_OP_ADD = 16, SB_OP_SUB, SB_OP_MUL, SB_OP_DIV, SB_OP_NEG
};
typedef struct sb_expr {
  int32_t n_ops;
  uint8_t ops[SB_EXPR_MAX_OPS];
  const double* vec[SB_EXPR_MAX_VEC];
  double scal[SB_EXPR_MAX_SCAL];
} sb_expr;
enum { SB_ASSIGN = 0, SB_ADD_ASSIGN = 1, SB_SUB_ASSIGN = 2, SB_MUL_ASSIGN = 3, SB_DIV_ASSIGN = 4 };
/* y (op)= expr, element-wise over n elements. */
SB_API int sb_eval(sb_ctx* ctx, double* y, size_t n, int assign_op, const sb_expr* expr);
SB_API int sb_fill(sb_ctx* ctx, double* y, size_t n, double value);
SB_API int sb_copy(sb_ctx* ctx, double* y, const double* x, size_t n);

/* A GROUP of element-wise statements and the reductions behind them in ONE kernel launch (one pass over the union of
 * the operands, one host synchronisation for all the reductions): what a solver's consecutive vector statements
 * cost when they are issued together instead of one kernel each. Every statement is a linear-combination chain, the
 * shape all of the reference solvers' updates have (e.g. SolverIdrs.hpp:195-198 `v <<= r - gamma_k*g_k; v -= gamma_i*g_i`,
 * SolverBiCgStab.hpp:271-273 `u_i <<= r_i - beta*u_i`):
 *     y = ((base (+|-) c0*x0) (+|-) c1*x1) ...      base == NULL: the chain starts from c0*x0 itself
 * evaluated per element in exactly that order, every product and every sum rounded separately -- bit-identical to
 * issuing `y <<= base - c0*x0; y -= c1*x1; ...` one statement at a time. Statements run in the order given; a later
 * statement (and the dots) see what an earlier one stored; y may alias base or any x of its own chain. After the last
 * statement the n_dots dot products <dot_a[d], dot_b[d]> are reduced with SB_TREE over the final values and
 * returned on the host in order. n_stmt and n_dots may be 0 (not both). */
#define SB_GROUP_MAX_STMT 8
#define SB_GROUP_MAX_TERMS 8
#define SB_GROUP_MAX_DOTS 8
typedef struct sb_chain {
  double* y;
  const double* base;                   /* may be NULL */
  int32_t n_terms;                      /* 1..SB_GROUP_MAX_TERMS (>= 1) */
  const double* x[SB_GROUP_MAX_TERMS];
  double c[SB_GROUP_MAX_TERMS];
  uint8_t sub[SB_GROUP_MAX_TERMS];      /* 0: + c*x, 1: - c*x (with base == NULL, sub[0] must be 0) */
} sb_chain;
SB_API int sb_eval_group(sb_ctx* ctx, size_t n, int n_stmt, const sb_chain* h_stmts, int n_dots,
                         const double* const* h_dot_a, const double* const* h_dot_b, double* h_out);

/* dot_product (MatrixAlgorithms.hpp:310-317) and norm_2 (:262-270: sqrt of the sum of squares) with the fixed
 * reduction tree "SB_TREE v1" (DESIGN.md): run-to-run and grid-size independent. Results are
 * returned on the host (one stream synchronisation per call). sb_dot_batch evaluates m dot
 * products with one synchronisation. */
SB_API int sb_dot(sb_ctx* ctx, const double* a, const double* b, size_t n, double* h_out);
SB_API int sb_norm2(sb_ctx* ctx, const double* a, size_t n, double* h_out);
SB_API int sb_dot_batch(sb_ctx* ctx, int m, const double* const* h_a, const double* const* h_b, size_t n,
                        double* h_out);

/* ---- fused solvers: CgSolver (SolverCg.hpp:54-126) and BiCgStabSolver (SolverBiCgStab.hpp:59-165)
 * driven as IterativeSolver::solve (Solver.hpp:116-147), no preconditioner. Same statements, same
 * per-element operation order and the same stopping rule as the reference; scalars stay on the
 * device and each dot/norm is fused into the kernel that produces its operand. */
typedef struct sb_solver_opts {
  int64_t num_iterations; /* Solver.hpp:67 (default 2000) */
  double abs_tol;         /* Solver.hpp:71 (default 1e-6); <= 0 disables */
  double rel_tol;         /* Solver.hpp:72 (default 1e-6); <= 0 disables */
  int32_t check_every;    /* host polls the device convergence flag every this many iterations (0 = 32) */
  int32_t use_graph;      /* 1: replay one captured CUDA graph per iteration */
  int32_t profile;        /* 1: bracket every kernel of the iteration with CUDA events (no graph) and
                             report the accumulated time per kernel slot in sb_solver_report.kernel_ms
                             (implies the stepwise schedule) */
  int32_t schedule;       /* SB_SCHEDULE_*; 0 = automatic: the stepwise schedule (the fastest one) */
  int32_t timeline_iters; /* persistent schedule: record the in-kernel timeline of the first this-many
                             iterations into h_timeline (0 = off) */
  uint64_t* h_timeline;   /* [timeline_iters][SB_TIMELINE_WORDS], globaltimer nanoseconds, written by CTA 0:
                             [0] iteration start; [1+b] time after grid barrier b of the iteration (BiCGStab:
                             b = 0 direction, 1 apply+dot, 2 half update, 3 apply+2 dots, 4 final update;
                             CG: 0 apply+dot, 1 update+dot, 2 direction); [6+b] ns CTA 0 spent in barrier b, from
                             its own arrival until it may go on (the wait for the slowest CTA, the reduction and
                             the all-reduce included); [11+b] reducing barriers: ns the CTA that ran the reduction
                             waited for the other ranks' sums after posting its own; [16+k] longest wait of any
                             warp of this rank for a neighbour's halo values in apply k */
  uint32_t tuning;        /* stepwise schedule: SB_TUNE_* bits, 0 = the library's choice (STREAM_OPERATOR | NO_ACK, plus
                             PUSH_ON_PRODUCE | PUSH_LAZY when a rank's apply kernel is at most ~two waves of tiles:
                             DESIGN.md 5d / 6). Results are bit-identical whatever the bits; they exist so that one
                             process can A/B them */
} sb_solver_opts;

/* Tuning bits of the stepwise schedule (sb_solver_opts::tuning; measurements: DESIGN.md 6):
 *   PUSH_ON_PRODUCE  multi-GPU, SB_COMM_P2P: the kernel that PRODUCES an apply's input vector (CG: the direction update;
 *                    BiCGStab: the direction and the half update) stores its boundary values straight into the
 *                    neighbours' halo tails as soon as a boundary tile is computed (boundary tiles are scheduled
 *                    first); the apply kernel has no pack CTAs and normally finds the halo flags already raised;
 *   NO_ACK           multi-GPU: the in-apply halo push skips the ack round (valid inside the fused solvers: between two
 *                    applies that write the same halo tail there is always an all-reduce);
 *   STREAM_OPERATOR  the apply kernel's bulk copies of the operator slices carry an L2 evict-first policy, so that a
 *                    rank whose vectors fit the 126 MB L2 (<= ~1.3 M cells) keeps them there between the kernels
 *                    (outside the fused solvers -- sb_apply, sb_gmres_solve, the generic path -- it is always on);
 *   PDL_FINAL        the one-CTA final stages are launched with programmatic stream serialization (resident while
 *                    the producer drains); PDL_AFTER_FINAL: so are the kernels that follow a final stage (resident,
 *                    operator slices prefetched, while the one-CTA stage runs on an otherwise idle machine);
 *   OFF              (with no other bit) switch every optional mechanism off: the round-1 behaviour. */
#define SB_TUNE_PUSH_ON_PRODUCE 1u
#define SB_TUNE_NO_ACK 2u
#define SB_TUNE_STREAM_OPERATOR 4u
#define SB_TUNE_PDL_FINAL 8u
#define SB_TUNE_PDL_AFTER_FINAL 16u
#define SB_TUNE_IN_KERNEL_REDUCER 64u /* the reductions of the ELEMENT-WISE steps are finished by the last CTA of the
                                         kernel that produces them: it takes the tile partials as they appear (the
                                         value is the flag), runs the all-reduce over the ranks and the scalar update;
                                         nobody waits in-kernel. An apply's reduction keeps its one-CTA final stage
                                         (4 / 7 launches per CG / BiCGStab iteration): carrying the role cost the apply
                                         kernel 6 % whether used or not */
#define SB_TUNE_PDL_APPLY 32u /* the apply kernels too (they follow an element-wise kernel; slice prefetch under its tail) */
#define SB_TUNE_PUSH_LAZY 128u /* with PUSH_ON_PRODUCE: the producer only issues the peer stores (posted writes that drain
                                  while the rest of it runs; its completion performs them); the flags are raised by the
                                  first CTA of the consuming apply: no fence, ticket or flag in the producer */
#define SB_TUNE_OFF 0x80000000u

/* Schedules of the fused CG / BiCGStab solvers (bit-identical results; measurements: DESIGN.md 5d):
 *   STEPWISE   one kernel per step of the iteration + a one-CTA kernel per reduction (final sum over the tile
 *              partials, all-reduce over the ranks through NVLink peer memory or NCCL, scalar update), replayed
 *              as a CUDA graph: 5 / 8 launches per CG / BiCGStab iteration. The fastest of the three on the B200
 *              at every size measured, hence the automatic choice;
 *   FOLDED     the same steps, but the last stage of every reduction is folded into the kernel that consumes the
 *              result (its CTA 0 reduces and raises a flag, the other CTAs have their loads in flight and wait
 *              for it): 3 / 5 launches per iteration. One GPU or SB_COMM_P2P. Saves the kernel boundaries but
 *              makes the whole first wave of the consumer wait for a reducer that now runs on a loaded machine:
 *              1-4 % slower than STEPWISE;
 *   PERSISTENT one cooperative kernel runs the whole iteration loop; steps are separated by grid-wide barriers
 *              in global memory, the reductions and the all-reduce happen inside the barrier. Needs the coefficient
 *              form; multi-GPU needs SB_COMM_P2P. A grid barrier (5-7 us: the arriving CTA first drains its stores)
 *              costs more than a kernel boundary inside a replayed graph (2.7 us): 5 % slower than STEPWISE for
 *              BiCGStab, on par for CG; kept for its in-kernel timeline (sb_solver_opts::h_timeline). */
#define SB_SCHEDULE_AUTO 0
#define SB_SCHEDULE_STEPWISE 1
#define SB_SCHEDULE_PERSISTENT 2
#define SB_SCHEDULE_FOLDED 3
#define SB_TIMELINE_WORDS 20

#define SB_MAX_KERNEL_SLOTS 8

typedef struct sb_solver_report {
  int32_t converged;
  int64_t iterations;  /* IterativeSolver::iteration after solve() */
  double initial_err;
  double abs_err;      /* IterativeSolver::absolute_error */
  double rel_err;      /* IterativeSolver::relative_error */
  int64_t n_hist;      /* entries written to h_hist: [0] initial, [k] after iteration k */
  int64_t n_trace;     /* entries written to h_trace: every reduction result, reference call order */
  double solve_ms;     /* device time of the whole solve incl. initialisation (CUDA events) */
  double iter_ms;      /* device time of the iteration loop only */
  int64_t launches;    /* kernels launched by this solve */
  int32_t n_kernel_slots;               /* kernels per iteration (CG 3, BiCGStab 5) */
  double kernel_ms[SB_MAX_KERNEL_SLOTS]; /* profile=1: total device time per kernel slot, in launch order
                                            (CG: apply+dot, update+dot, direction; BiCGStab: direction,
                                            apply+dot, half update, apply+2 dots, final update+2 dots) */
  int32_t schedule;    /* SB_SCHEDULE_STEPWISE, _FOLDED or _PERSISTENT: the one that ran */
  double wait_ms[SB_MAX_KERNEL_SLOTS];   /* profile=1: per kernel slot, in-kernel waits summed over the iterations.
                                            Apply slots: the longest wait of a boundary CTA for a neighbour's halo
                                            values + (STEPWISE) the wait of the one-CTA stage behind the apply for
                                            the other ranks' partial sums. Other slots: the wait for the other
                                            ranks' sums of the reduction behind / (FOLDED) in front of the step;
                                            FOLDED on one GPU: the duration of the reducer's chain instead */
  double ar_wait_ms[SB_MAX_KERNEL_SLOTS]; /* profile=1, STEPWISE on several GPUs: per kernel slot, how long the one-CTA
                                            final stage behind the slot waited for the other ranks' sums after posting
                                            its own (rank skew + one NVLink crossing), summed over the iterations */
  double final_ms[SB_MAX_KERNEL_SLOTS];  /* profile=1, STEPWISE with one-CTA final stages: the part of kernel_ms[k] that
                                            is the final stage behind the slot's kernel (from an event recorded between
                                            the two launches): kernel_ms[k] - final_ms[k] is the reducing kernel alone */
} sb_solver_report;

SB_API int sb_cg_solve(sb_ctx* ctx, const sb_op* op, double* x, const double* b,
                       const sb_solver_opts* opts, sb_solver_report* report, double* h_hist,
                       int64_t hist_cap, double* h_trace, int64_t trace_cap);
SB_API int sb_bicgstab_solve(sb_ctx* ctx, const sb_op* op, double* x, const double* b,
                             const sb_solver_opts* opts, sb_solver_report* report, double* h_hist,
                             int64_t hist_cap, double* h_trace, int64_t trace_cap);
/* Fused restarted GMRES(m): GmresSolver / FgmresSolver without a preconditioner (SolverGmres.hpp:42-310) driven
 * as InnerOuterIterativeSolver (Solver.hpp:154-259) + IterativeSolver::solve. The Arnoldi process runs on the
 * device without host round trips (each modified-Gram-Schmidt step is one kernel: subtract the previous
 * projection and reduce the next one; coefficients are read from device memory), the Givens rotations, the
 * residual estimate and the stopping rule are the reference's own scalar statements evaluated on the host a few
 * steps behind, from asynchronously copied H columns. Bit-identical to the reference headers (oracle tree
 * reductions): iteration count, every reduction scalar (h_trace), residual history, solution.
 * One deliberate deviation: when the initial residual is already below abs_tol (or num_iterations == 0) the
 * reference's finalize() divides by H(0,0) = 0 and turns x into NaN; here x is left untouched. */
typedef struct sb_gmres_opts {
  int64_t num_iterations;       /* Solver.hpp:67 (default 2000); one iteration = one inner Arnoldi step */
  double abs_tol;               /* <= 0 disables */
  double rel_tol;               /* <= 0 disables */
  int32_t num_inner_iterations; /* restart length m, Solver.hpp:159; 0 = the reference's default 50; max 128 */
  int32_t lookahead;            /* Arnoldi steps the device may run ahead of the host's stopping test (0 = 3) */
} sb_gmres_opts;
SB_API int sb_gmres_solve(sb_ctx* ctx, const sb_op* op, double* x, const double* b, const sb_gmres_opts* opts,
                          sb_solver_report* report, double* h_hist, int64_t hist_cap, double* h_trace,
                          int64_t trace_cap);

/* Same, with HOST buffers for x (in: initial guess, out: solution) and b: the copies are part
 * of the call (bench.py's e2e figure). solver: "cg" | "bicgstab". */
SB_API int sb_solve_host(sb_ctx* ctx, const sb_op* op, const char* solver, double* h_x,
                         const double* h_b, const sb_solver_opts* opts, sb_solver_report* report,
                         double* h_hist, int64_t hist_cap);

#ifdef __cplusplus
}
#endif
#endif /* STORMB200_H */
