// sb_op.cu -- operator construction (face list -> cell rows) and sb_apply.
//
// The row build is host-side integer work done once per mesh: bucket the faces by cell in ascending
// face index (Playground.cpp's loop order), pad to the ELL width, compute the per-entry
// coefficients with the same operation order as the oracle (oracle/sb_oracle.c: orc_build_rows,
// orc_rows_faithful, orc_rows_coef) so the uploaded arrays can be compared bit for bit.
#include "sb_op.cuh"

#include <algorithm>
#include <memory>
#include <thread>

using namespace sb;

namespace {

struct HostRows {
  int32_t width = 0;
  int64_t ld = 0, entries = 0;
  std::vector<int32_t> col;
  std::vector<double> v0, v1, diag;
};

// ELL width for a maximum row degree: exact up to 8 (tets 4, hexes 6); wider rows (polyhedral cells) are
// padded to the next even width up to 16, the widths the apply kernels are instantiated for.
constexpr int32_t kMaxWidth = 16;
inline int32_t ell_width(int32_t max_deg) {
  const int32_t w = std::max<int32_t>(1, max_deg);
  return w <= 8 ? w : ((w + 1) & ~1);
}

// n_rows: rows are built for cells [0, n_rows) only (distributed operator: the owned cells; columns
// may reference any cell < n_cells, i.e. the halo tail). n_rows == n_cells on a single GPU.
int build_rows(const sb_mesh_soa* m, const sb_op_desc* desc, int64_t n_rows, HostRows& R) {
  const int64_t n = m->n_cells, F = m->n_faces, B = m->n_bfaces;
  for (int64_t f = 0; f < 2 * F; ++f)
    SB_REQUIRE(m->face_cell[f] >= 0 && m->face_cell[f] < n, "face_cell index out of range");
  for (int64_t b = 0; b < B; ++b)
    SB_REQUIRE(m->bface_cell[b] >= 0 && m->bface_cell[b] < n_rows, "bface_cell index out of range");
  const bool coef = desc->form == SB_FORM_COEF;
  // degrees: interior entries, plus ghost entries in the faithful form
  std::vector<int32_t> deg((size_t) n + 1, 0);
  for (int64_t f = 0; f < F; ++f) deg[m->face_cell[2 * f]]++, deg[m->face_cell[2 * f + 1]]++;
  if (!coef)
    for (int64_t b = 0; b < B; ++b) deg[m->bface_cell[b]]++;
  int32_t max_deg = 1;
  for (int64_t i = 0; i < n_rows; ++i) max_deg = std::max(max_deg, deg[i]);
  const int32_t width = ell_width(max_deg);
  SB_REQUIRE(width <= kMaxWidth, "cells with more than 16 faces are not supported");
  const int64_t ld = pad_up(n_rows);
  R.width = width, R.ld = ld, R.entries = 0;
  R.col.assign((size_t) width * ld, kColPad);
  R.v0.assign((size_t) width * ld, 0.0);
  if (coef) R.diag.assign((size_t) ld, 0.0);
  else R.v1.assign((size_t) width * ld, 1.0);
  std::vector<int32_t> fill((size_t) n + 1, 0);
  const double dt = desc->dt;
  if (coef)
    for (int64_t i = 0; i < n_rows; ++i) R.diag[i] = desc->prefill ? 1.0 : 0.0;
  auto put = [&](int32_t row, int32_t c, double area, double dist) {
    if (row >= n_rows) return; // a halo cell: its row lives on the owning rank
    const int64_t e = (int64_t) fill[row] * ld + row;
    if (coef) {
      const double a = ((area / m->cell_vol[row]) * dt) / dist;
      R.col[e] = c, R.v0[e] = a;
      R.diag[row] = R.diag[row] - a;
    } else {
      R.col[e] = c, R.v0[e] = area / m->cell_vol[row], R.v1[e] = dist;
    }
    fill[row]++;
    R.entries++;
  };
  for (int64_t f = 0; f < F; ++f) {
    const int32_t ci = m->face_cell[2 * f], co = m->face_cell[2 * f + 1];
    put(ci, co, m->face_area[f], m->face_dist[f]);
    put(co, ci, m->face_area[f], m->face_dist[f]);
  }
  for (int64_t b = 0; b < B; ++b) {
    const int32_t ci = m->bface_cell[b];
    if (coef) {
      // mirror ghost folds into the diagonal: y_i += g*dt*(-x_i - x_i)/d  ->  diag -= (a + a)
      const double a = ((m->bface_area[b] / m->cell_vol[ci]) * dt) / m->bface_dist[b];
      R.diag[ci] = R.diag[ci] - (a + a);
    } else {
      put(ci, ~ci, m->bface_area[b], m->bface_dist[b]);
    }
  }
  return SB_OK;
}

// Convection-diffusion rows (include/stormb200.h: sb_convdiff_desc). Same entry order as build_rows
// (ascending face index per row); the coefficient arithmetic is restated operation by operation in
// oracle/sb_oracle.c: orc_rows_convdiff.
int build_rows_convdiff(const sb_mesh_soa* m, const sb_convdiff_desc* desc, int64_t n_rows, HostRows& R) {
  const int64_t n = m->n_cells, F = m->n_faces, B = m->n_bfaces;
  for (int64_t f = 0; f < 2 * F; ++f)
    SB_REQUIRE(m->face_cell[f] >= 0 && m->face_cell[f] < n, "face_cell index out of range");
  for (int64_t b = 0; b < B; ++b)
    SB_REQUIRE(m->bface_cell[b] >= 0 && m->bface_cell[b] < n_rows, "bface_cell index out of range");
  std::vector<int32_t> deg((size_t) n + 1, 0);
  for (int64_t f = 0; f < F; ++f) deg[m->face_cell[2 * f]]++, deg[m->face_cell[2 * f + 1]]++;
  int32_t max_deg = 1;
  for (int64_t i = 0; i < n_rows; ++i) max_deg = std::max(max_deg, deg[i]);
  const int32_t width = ell_width(max_deg);
  SB_REQUIRE(width <= kMaxWidth, "cells with more than 16 faces are not supported");
  const int64_t ld = pad_up(n_rows);
  R.width = width, R.ld = ld, R.entries = 0;
  R.col.assign((size_t) width * ld, kColPad);
  R.v0.assign((size_t) width * ld, 0.0);
  R.diag.assign((size_t) ld, 0.0);
  std::vector<int32_t> fill((size_t) n + 1, 0);
  const double nu = desc->nu;
  auto put = [&](int32_t row, int32_t c, double a, double dg) {
    if (row >= n_rows) return;
    const int64_t e = (int64_t) fill[row] * ld + row;
    R.col[e] = c, R.v0[e] = a;
    R.diag[row] = R.diag[row] + dg;
    fill[row]++;
    R.entries++;
  };
  for (int64_t f = 0; f < F; ++f) {
    const int32_t ci = m->face_cell[2 * f], co = m->face_cell[2 * f + 1];
    const double un = desc->face_un[f];
    const double up = un > 0.0 ? un : 0.0, um = un < 0.0 ? un : 0.0;
    const double kd = nu / m->face_dist[f];
    if (ci < n_rows) {
      const double g = m->face_area[f] / m->cell_vol[ci];
      put(ci, co, g * (um - kd), g * (up + kd));
    }
    if (co < n_rows) {
      const double g = m->face_area[f] / m->cell_vol[co];
      put(co, ci, g * ((-up) - kd), g * (kd - um));
    }
  }
  for (int64_t b = 0; b < B; ++b) {
    const int32_t ci = m->bface_cell[b];
    const double un = desc->bface_un[b];
    const double up = un > 0.0 ? un : 0.0, um = un < 0.0 ? un : 0.0;
    const double kd = nu / m->bface_dist[b];
    const double g = m->bface_area[b] / m->cell_vol[ci];
    R.diag[ci] = R.diag[ci] + g * ((up - um) + (kd + kd));
  }
  return SB_OK;
}

template<class T>
int upload(sb_ctx* ctx, const std::vector<T>& h, void** d_out, int64_t& bytes) {
  *d_out = nullptr;
  if (h.empty()) return SB_OK;
  cudaError_t e = cudaMalloc(d_out, sizeof(T) * h.size());
  if (e == cudaErrorMemoryAllocation) {
    (void) cudaGetLastError();
    vec_cache_release(ctx); // cached vector blocks (sb_vec_free keeps them for reuse) give way to the operator
    e = cudaMalloc(d_out, sizeof(T) * h.size());
  }
  SB_CUDA(e);
  SB_CUDA(cudaMemcpyAsync(*d_out, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice, ctx->stream));
  bytes += (int64_t) (sizeof(T) * h.size());
  return SB_OK;
}

} // namespace

extern "C" {

static int validate_mesh(sb_ctx* ctx, const sb_mesh_soa* m, int64_t n_rows, sb_op** out) {
  SB_REQUIRE(ctx != nullptr && m != nullptr && out != nullptr, "null argument");
  *out = nullptr;
  SB_REQUIRE(n_rows > 0 && n_rows <= m->n_cells, "row count out of range");
  SB_REQUIRE(m->n_cells > 0 && m->n_cells < (int64_t) INT32_MAX - kTile, "n_cells out of range (int32 indices)");
  SB_REQUIRE(m->n_faces >= 0 && m->n_bfaces >= 0, "negative face count");
  SB_REQUIRE(m->n_faces == 0 || (m->face_cell && m->face_area && m->face_dist), "null face arrays");
  SB_REQUIRE(m->n_bfaces == 0 || (m->bface_cell && m->bface_area && m->bface_dist), "null boundary-face arrays");
  SB_REQUIRE(m->cell_vol != nullptr, "null cell_vol");
  return SB_OK;
}

// Upload the host rows: blocked SELL-64 records for the coefficient form, plain ELL arrays otherwise.
static int finish_op(sb_ctx* ctx, HostRows& R, int form, int prefill, double dt, int64_t n_rows, sb_op** out) {
  std::unique_ptr<sb_op> op(new sb_op());
  op->d.n = n_rows, op->d.ld = R.ld, op->d.width = R.width, op->d.form = form;
  op->d.prefill = prefill, op->d.dt = dt;
  op->n_entries = R.entries;
  if (const char* dbg = std::getenv("SB_DEBUG")) op->d.debug = std::atoi(dbg) & 1;
  SB_CUDA(cudaSetDevice(ctx->device));
  std::vector<unsigned char> blk;
  if (form == SB_FORM_COEF) {
    // blocked layout: slice record = [col[W][64] | coef[W][64] | diag[64]]
    const int W = R.width;
    const int64_t slice_bytes = 768 * (int64_t) W + 512, n_slices = R.ld / 64;
    blk.resize((size_t) (slice_bytes * n_slices));
    auto pack = [&](int64_t lo, int64_t hi) {
      for (int64_t sl = lo; sl < hi; ++sl) {
        unsigned char* rec = blk.data() + sl * slice_bytes;
        for (int k = 0; k < W; ++k) {
          std::memcpy(rec + k * 256, &R.col[(size_t) k * R.ld + sl * 64], 256);
          std::memcpy(rec + W * 256 + k * 512, &R.v0[(size_t) k * R.ld + sl * 64], 512);
        }
        std::memcpy(rec + W * 768, &R.diag[(size_t) sl * 64], 512);
      }
    };
    const unsigned T = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::vector<std::thread> th;
    for (unsigned t = 0; t < T; ++t) th.emplace_back(pack, n_slices * t / T, n_slices * (t + 1) / T);
    for (auto& x : th) x.join();
    SB_TRY(upload(ctx, blk, &op->buffers[4], op->device_bytes));
    op->d.blk = (const unsigned char*) op->buffers[4];
    op->d.slice_bytes = (int32_t) slice_bytes;
  } else {
    SB_TRY(upload(ctx, R.col, &op->buffers[0], op->device_bytes));
    SB_TRY(upload(ctx, R.v0, &op->buffers[1], op->device_bytes));
    SB_TRY(upload(ctx, R.v1, &op->buffers[2], op->device_bytes));
    SB_TRY(upload(ctx, R.diag, &op->buffers[3], op->device_bytes));
  }
  SB_CUDA(cudaStreamSynchronize(ctx->stream));
  op->d.col = (const int32_t*) op->buffers[0];
  op->d.v0 = (const double*) op->buffers[1];
  op->d.v1 = (const double*) op->buffers[2];
  op->d.diag = (const double*) op->buffers[3];
  *out = op.release();
  return SB_OK;
}

static int op_create(sb_ctx* ctx, const sb_mesh_soa* m, const sb_op_desc* desc, int64_t n_rows, sb_op** out) {
  SB_REQUIRE(desc != nullptr, "null argument");
  SB_TRY(validate_mesh(ctx, m, n_rows, out));
  SB_REQUIRE(desc->form == SB_FORM_FAITHFUL || desc->form == SB_FORM_COEF, "unknown operator form");
  SB_REQUIRE(desc->prefill == 0 || desc->prefill == 1, "prefill must be 0 or 1");
  // The dependency order of the ghost fold matters for bit-exactness in the coef form: the oracle
  // subtracts interior coefficients and ghost terms in row order (interior first, then ghosts),
  // which is exactly the order of the two loops in build_rows.
  HostRows R;
  SB_TRY(build_rows(m, desc, n_rows, R));
  return finish_op(ctx, R, desc->form, desc->prefill, desc->dt, n_rows, out);
}

static int op_create_convdiff(sb_ctx* ctx, const sb_mesh_soa* m, const sb_convdiff_desc* desc, int64_t n_rows, sb_op** out) {
  SB_REQUIRE(desc != nullptr, "null argument");
  SB_TRY(validate_mesh(ctx, m, n_rows, out));
  SB_REQUIRE(m->n_faces == 0 || desc->face_un != nullptr, "null face_un");
  SB_REQUIRE(m->n_bfaces == 0 || desc->bface_un != nullptr, "null bface_un");
  SB_REQUIRE(desc->nu >= 0.0, "nu must be >= 0");
  HostRows R;
  SB_TRY(build_rows_convdiff(m, desc, n_rows, R));
  return finish_op(ctx, R, SB_FORM_COEF, 0, 0.0, n_rows, out);
}

// Halo plan of a distributed operator (shared by the diffusion and convection-diffusion forms).
static int attach_halo(sb_ctx* ctx, const sb_local_mesh* loc, sb_op* op) {
  op->distributed = true;
  op->halo_base = loc->halo_base, op->n_halo = loc->n_halo;
  HaloDev& h = op->halo;
  h.n_nbr = loc->n_nbr;
  h.first_boundary_tile = (int32_t) (loc->n_interior / kTile);
  const int64_t total = loc->n_nbr > 0 ? loc->send_ptr[loc->n_nbr] : 0;
  for (int k = 0; k < loc->n_nbr; ++k) {
    h.nbr_rank[k] = loc->nbr_rank[k], h.send_dst[k] = loc->send_dst[k];
    SB_REQUIRE(loc->nbr_rank[k] >= 0 && loc->nbr_rank[k] < ctx->comm.world && loc->nbr_rank[k] != ctx->comm.rank, "bad neighbour rank");
  }
  for (int k = 0; k <= loc->n_nbr; ++k) h.send_ptr[k] = loc->send_ptr[k], op->recv_ptr[k] = loc->recv_ptr[k];
  for (int64_t i = 0; i < total; ++i)
    SB_REQUIRE(loc->send_idx[i] >= 0 && loc->send_idx[i] < loc->n_owned, "send_idx out of range");
  if (total > 0) {
    SB_CUDA(cudaMalloc(&op->d_send_idx, sizeof(int32_t) * (size_t) total));
    SB_CUDA(cudaMemcpy(op->d_send_idx, loc->send_idx, sizeof(int32_t) * (size_t) total, cudaMemcpyHostToDevice));
  }
  h.send_idx = op->d_send_idx;
  // push-on-produce plan: the same entries sorted by the tile of their source cell (stable: within a tile they keep the
  // neighbour-major order of the send list), so that the CTA that produces a boundary tile can forward its share
  const int64_t owned_tiles = num_tiles(loc->n_owned);
  h.n_push_tiles = (int32_t) (owned_tiles - h.first_boundary_tile);
  if (total > 0 && h.n_push_tiles > 0 && loc->n_owned < (1 << 28) && total < INT32_MAX) {
    std::vector<int32_t> ptr((size_t) h.n_push_tiles + 1, 0);
    for (int64_t i = 0; i < total; ++i) {
      const int64_t q = loc->send_idx[i] / kTile - h.first_boundary_tile;
      SB_REQUIRE(q >= 0 && q < h.n_push_tiles, "send_idx names a cell in front of the boundary block of the local order");
      ptr[(size_t) q + 1]++;
    }
    for (int32_t q = 0; q < h.n_push_tiles; ++q) ptr[(size_t) q + 1] += ptr[(size_t) q];
    std::vector<int32_t> fill(ptr.begin(), ptr.end() - 1);
    std::vector<int2> entry((size_t) total);
    for (int k = 0; k < loc->n_nbr; ++k)
      for (int64_t i = loc->send_ptr[k]; i < loc->send_ptr[k + 1]; ++i) {
        const int64_t q = loc->send_idx[i] / kTile - h.first_boundary_tile;
        const int64_t dst = loc->send_dst[k] + (i - loc->send_ptr[k]);
        SB_REQUIRE(dst >= 0 && dst < INT32_MAX, "halo offset does not fit 31 bits");
        entry[(size_t) fill[(size_t) q]++] = make_int2((int32_t) ((uint32_t) loc->send_idx[i] | ((uint32_t) k << 28)), (int32_t) dst);
      }
    SB_CUDA(cudaMalloc(&op->d_push_ptr, sizeof(int32_t) * ptr.size()));
    SB_CUDA(cudaMemcpy(op->d_push_ptr, ptr.data(), sizeof(int32_t) * ptr.size(), cudaMemcpyHostToDevice));
    SB_CUDA(cudaMalloc(&op->d_push_entry, sizeof(int2) * entry.size()));
    SB_CUDA(cudaMemcpy(op->d_push_entry, entry.data(), sizeof(int2) * entry.size(), cudaMemcpyHostToDevice));
    h.push_ptr = op->d_push_ptr, h.push_entry = op->d_push_entry;
  }
  return SB_OK;
}

static int check_local(sb_ctx* ctx, const sb_local_mesh* loc) {
  if (ctx->comm.mode < 0) {
    set_error("a distributed operator needs a communicator (sb_comm_prepare / sb_comm_connect)");
    return SB_ERR_STATE;
  }
  SB_REQUIRE(loc->rank == ctx->comm.rank && loc->n_parts == ctx->comm.world, "local mesh belongs to another rank / world size");
  SB_REQUIRE(loc->n_nbr >= 0 && loc->n_nbr < kMaxRanks, "too many neighbours");
  SB_REQUIRE(loc->n_owned > 0, "a rank without cells cannot take part (reductions and the apply sequence are collective)");
  SB_REQUIRE(loc->halo_base == pad_up(loc->n_owned) && loc->soa.n_cells == loc->halo_base + loc->n_halo, "local mesh layout");
  SB_REQUIRE(loc->halo_base + loc->n_halo <= ctx->vec_capacity, "local vector does not fit the pool block (vec_capacity)");
  return SB_OK;
}

int sb_op_create(sb_ctx* ctx, const sb_mesh_soa* m, const sb_op_desc* desc, sb_op** out) {
  SB_REQUIRE(m != nullptr, "null argument");
  return op_create(ctx, m, desc, m->n_cells, out);
}

int sb_op_create_convdiff(sb_ctx* ctx, const sb_mesh_soa* m, const sb_convdiff_desc* desc, sb_op** out) {
  SB_REQUIRE(m != nullptr, "null argument");
  return op_create_convdiff(ctx, m, desc, m->n_cells, out);
}

int sb_dist_op_create(sb_ctx* ctx, const sb_local_mesh* loc, const sb_op_desc* desc, sb_op** out) {
  SB_REQUIRE(ctx != nullptr && loc != nullptr && desc != nullptr && out != nullptr, "null argument");
  *out = nullptr;
  SB_TRY(check_local(ctx, loc));
  sb_op* op = nullptr;
  SB_TRY(op_create(ctx, &loc->soa, desc, loc->n_owned, &op));
  const int rc = attach_halo(ctx, loc, op);
  if (rc != SB_OK) {
    sb_op_destroy(ctx, op);
    return rc;
  }
  *out = op;
  return SB_OK;
}

int sb_dist_op_create_convdiff(sb_ctx* ctx, const sb_local_mesh* loc, const sb_convdiff_desc* desc, sb_op** out) {
  SB_REQUIRE(ctx != nullptr && loc != nullptr && desc != nullptr && out != nullptr, "null argument");
  *out = nullptr;
  SB_TRY(check_local(ctx, loc));
  sb_op* op = nullptr;
  SB_TRY(op_create_convdiff(ctx, &loc->soa, desc, loc->n_owned, &op));
  const int rc = attach_halo(ctx, loc, op);
  if (rc != SB_OK) {
    sb_op_destroy(ctx, op);
    return rc;
  }
  *out = op;
  return SB_OK;
}

int sb_op_destroy(sb_ctx* ctx, sb_op* op) {
  SB_REQUIRE(ctx != nullptr, "ctx is null");
  if (op == nullptr) return SB_OK;
  SB_CUDA(cudaStreamSynchronize(ctx->stream));
  for (void* b : op->buffers) cudaFree(b);
  cudaFree(op->d_send_idx);
  cudaFree(op->d_push_ptr);
  cudaFree(op->d_push_entry);
  delete op;
  return SB_OK;
}

int sb_op_get_info(const sb_op* op, sb_op_info* info) {
  SB_REQUIRE(op != nullptr && info != nullptr, "null argument");
  info->n_cells = op->d.n;
  info->n_entries = op->n_entries;
  info->width = op->d.width;
  info->ld = op->d.ld;
  info->form = op->d.form;
  info->device_bytes = op->device_bytes;
  // SURVEY.md 8d contract figure: N*(8 x + 8 y + 8 diag) + entries*(4 col + 8 coef)
  info->algorithmic_bytes_per_apply = 24 * op->d.n + 12 * op->n_entries;
  return SB_OK;
}

int sb_op_download_rows(sb_ctx* ctx, const sb_op* op, int32_t* h_col, double* h_val0, double* h_val1,
                        double* h_diag) {
  SB_REQUIRE(ctx != nullptr && op != nullptr, "null argument");
  const size_t wl = (size_t) op->d.width * (size_t) op->d.ld;
  if (op->d.blk != nullptr) {
    // blocked layout: fetch the records and unpack them into the canonical [W][ld] arrays
    const int W = op->d.width;
    const int64_t sb_ = op->d.slice_bytes, n_slices = op->d.ld / 64, ld = op->d.ld;
    std::vector<unsigned char> blk((size_t) (sb_ * n_slices));
    SB_CUDA(cudaMemcpyAsync(blk.data(), op->d.blk, blk.size(), cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int64_t sl = 0; sl < n_slices; ++sl) {
      const unsigned char* rec = blk.data() + sl * sb_;
      for (int k = 0; k < W; ++k) {
        if (h_col) std::memcpy(h_col + (size_t) k * ld + sl * 64, rec + k * 256, 256);
        if (h_val0) std::memcpy(h_val0 + (size_t) k * ld + sl * 64, rec + W * 256 + k * 512, 512);
      }
      if (h_diag) std::memcpy(h_diag + sl * 64, rec + W * 768, 512);
    }
    return SB_OK;
  }
  if (h_col) SB_CUDA(cudaMemcpyAsync(h_col, op->d.col, sizeof(int32_t) * wl, cudaMemcpyDeviceToHost, ctx->stream));
  if (h_val0) SB_CUDA(cudaMemcpyAsync(h_val0, op->d.v0, sizeof(double) * wl, cudaMemcpyDeviceToHost, ctx->stream));
  if (h_val1 && op->d.v1)
    SB_CUDA(cudaMemcpyAsync(h_val1, op->d.v1, sizeof(double) * wl, cudaMemcpyDeviceToHost, ctx->stream));
  if (h_diag && op->d.diag)
    SB_CUDA(cudaMemcpyAsync(h_diag, op->d.diag, sizeof(double) * (size_t) op->d.ld, cudaMemcpyDeviceToHost, ctx->stream));
  SB_CUDA(cudaStreamSynchronize(ctx->stream));
  return SB_OK;
}

int sb_op_jacobi(sb_ctx* ctx, const sb_op* op, const double* x, double* y) {
  SB_REQUIRE(ctx != nullptr && op != nullptr && x != nullptr && y != nullptr, "null argument");
  SB_REQUIRE(op->d.form == SB_FORM_COEF, "the Jacobi preconditioner needs the coefficient form (it stores the diagonal)");
  JacobiBody body{op->d, x, y};
  SB_CUDA(launch_kernel(ctx, ew_kernel<0, JacobiBody>, (unsigned) num_tiles(op->d.n), kThreads, 0, op->d.n, body, RedPtrs{},
                        (const int*) nullptr));
  ctx->launches++;
  return SB_OK;
}

int sb_apply(sb_ctx* ctx, const sb_op* op, const double* x, double* y) {
  SB_REQUIRE(ctx != nullptr && op != nullptr && x != nullptr && y != nullptr, "null argument");
  SB_REQUIRE(x != y, "sb_apply: x and y must not alias");
  return launch_apply<0, false>(ctx, op, x, y, NoEpi{}, NoFinal{}, nullptr);
}

int sb_apply_dot(sb_ctx* ctx, const sb_op* op, const double* x, double* y, const double* u, double* h_out) {
  SB_REQUIRE(ctx != nullptr && op != nullptr && x != nullptr && y != nullptr && h_out != nullptr, "null argument");
  SB_REQUIRE(x != y, "sb_apply_dot: x and y must not alias");
  SB_REQUIRE(u != y, "sb_apply_dot: u and y must not alias (the dot reads u before y exists)");
  const StoreFinal<1> fin{ctx->red.result};
  if (u == nullptr || u == x) {
    SB_TRY((launch_apply<1, false>(ctx, op, x, y, EpiXY{}, fin, nullptr)));
  } else {
    SB_TRY((launch_apply<1, false>(ctx, op, x, y, EpiUY{u}, fin, nullptr)));
  }
  SB_CUDA(cudaMemcpyAsync(ctx->h_pinned, ctx->red.result, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  SB_CUDA(cudaStreamSynchronize(ctx->stream));
  *h_out = ctx->h_pinned[0];
  return SB_OK;
}

int sb_apply_dot_yy_yx(sb_ctx* ctx, const sb_op* op, const double* x, double* y, double* h_out) {
  SB_REQUIRE(ctx != nullptr && op != nullptr && x != nullptr && y != nullptr && h_out != nullptr, "null argument");
  SB_REQUIRE(x != y, "sb_apply_dot_yy_yx: x and y must not alias");
  SB_TRY((launch_apply<2, false>(ctx, op, x, y, EpiYYandYX{}, StoreFinal<2>{ctx->red.result}, nullptr)));
  SB_CUDA(cudaMemcpyAsync(ctx->h_pinned, ctx->red.result, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  SB_CUDA(cudaStreamSynchronize(ctx->stream));
  h_out[0] = ctx->h_pinned[0], h_out[1] = ctx->h_pinned[1];
  return SB_OK;
}

int sb_apply_accumulate(sb_ctx* ctx, const sb_op* op, double dt, const double* x, double* y) {
  SB_REQUIRE(ctx != nullptr && op != nullptr && x != nullptr && y != nullptr, "null argument");
  SB_REQUIRE(x != y, "sb_apply_accumulate: x and y must not alias");
  SB_REQUIRE(op->d.form == SB_FORM_FAITHFUL,
             "sb_apply_accumulate needs a faithful-form operator (the per-face terms, not dt-scaled coefficients)");
  OpDev d = op->d;
  d.dt = dt, d.prefill = 2;
  ApplyOpts ao;
  ao.per_call = &d;
  return launch_apply<0, false>(ctx, op, x, y, NoEpi{}, NoFinal{}, nullptr, ao);
}

} // extern "C"
