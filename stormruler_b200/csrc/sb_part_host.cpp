// sb_part_host.cpp -- partitioning of the cell graph and the per-rank local meshes / halo maps.
//
// The reference has no partitioning or halo code; this is new host-side integer
// work specified by SURVEY.md 8e. Everything here is deterministic and restated independently in
// oracle/mesh_oracle.py, against which the tests compare every array bit for bit.
//
// Local numbering of rank r (vectors of that rank are laid out this way):
//   [ interior owned | boundary owned | padding to the 2048-row tile | halo ]
//   * owned cells keep their global relative order inside each of the two groups; a cell is
//     "boundary" when at least one of its faces leads to a cell owned by another rank;
//   * halo cells are grouped by owner rank (ascending), ascending global id inside a group, and
//     start at halo_base = pad_up(n_owned), so the halo tail of a vector is tile-aligned and never
//     shares a cache line with owned cells;
//   * rank a's send list to b (a's owned cells that b reads, ascending global id) is by construction
//     the same sequence of global ids as b's halo group for owner a: message k-th value = k-th slot.
// The local face list keeps every face with an owned endpoint, in ascending GLOBAL face index and
// with the global inner/outer orientation, so the rows built from it (sb_op.cu: build_rows) sum in
// exactly the order of the single-GPU rows: each owned row is bit-identical to the global one.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <memory>
#include <numeric>
#include <vector>

#include "../../include/stormb200.h"

namespace sb {
void set_error(const char* fmt, ...);
}

#define SBP_REQUIRE(cond, msg)                                                 \
  do {                                                                         \
    if (!(cond)) {                                                             \
      ::sb::set_error("%s:%d: invalid argument: %s", __FILE__, __LINE__, msg); \
      return SB_ERR_INVALID;                                                   \
    }                                                                          \
  } while (0)

// METIS 5.x as bundled with the CUDA toolkit (libmetis_static.a): 64-bit idx_t (verified by probe,
// SURVEY.md App. A-6); real_t arrays are passed as NULL so their width does not matter.
extern "C" {
int METIS_SetDefaultOptions(int64_t* options);
int METIS_PartGraphKway(int64_t* nvtxs, int64_t* ncon, int64_t* xadj, int64_t* adjncy, int64_t* vwgt, int64_t* vsize,
                        int64_t* adjwgt, int64_t* nparts, void* tpwgts, void* ubvec, int64_t* options, int64_t* objval,
                        int64_t* part);
}

namespace {

constexpr int64_t kTileRows = 2048; // == sb::kTile (sb_common.cuh); vectors and rows are padded to it
inline int64_t pad_up(int64_t n) { return ((n + kTileRows - 1) / kTileRows) * kTileRows; }

struct Local {
  bool built = false;
  int64_t n_owned = 0, n_interior = 0, n_halo = 0, halo_base = 0;
  std::vector<int32_t> l2g;
  // local face list
  std::vector<int32_t> face_cell, bface_cell;
  std::vector<double> face_area, face_dist, cell_vol, bface_area, bface_dist;
  std::vector<int64_t> face_global, bface_global;
  // neighbours
  std::vector<int32_t> nbr_rank, send_idx;
  std::vector<int64_t> send_ptr, recv_ptr, send_dst;
};

} // namespace

struct sb_part {
  int32_t n_parts = 0;
  sb_mesh_soa g{}; // the global mesh (borrowed: the sb_mesh must outlive the partition)
  std::vector<int32_t> part;
  std::vector<int64_t> owned_count;
  std::vector<int64_t> cnt;        // [owner * n_parts + reader]: cells of `owner` in the halo of `reader`
  std::vector<int64_t> halo_count; // [reader]
  int64_t edge_cut = 0;
  std::vector<Local> local;
};

namespace {

int finish(sb_part& P) {
  const int64_t n = P.g.n_cells, F = P.g.n_faces;
  P.owned_count.assign((size_t) P.n_parts, 0);
  for (int64_t i = 0; i < n; ++i) {
    SBP_REQUIRE(P.part[(size_t) i] >= 0 && P.part[(size_t) i] < P.n_parts, "part id out of range");
    P.owned_count[(size_t) P.part[(size_t) i]]++;
  }
  for (int r = 0; r < P.n_parts; ++r) SBP_REQUIRE(P.owned_count[(size_t) r] > 0, "a part owns no cells");
  P.edge_cut = 0;
  for (int64_t f = 0; f < F; ++f)
    P.edge_cut += P.part[(size_t) P.g.face_cell[2 * f]] != P.part[(size_t) P.g.face_cell[2 * f + 1]];
  // halo census: a cell is in rank r's halo iff it has a face to a cell of r and is not owned by r
  {
    std::vector<std::pair<int32_t, int32_t>> pairs; // (reader, cell)
    for (int64_t f = 0; f < F; ++f) {
      const int32_t a = P.g.face_cell[2 * f], b = P.g.face_cell[2 * f + 1];
      const int32_t pa = P.part[(size_t) a], pb = P.part[(size_t) b];
      if (pa != pb) pairs.emplace_back(pa, b), pairs.emplace_back(pb, a);
    }
    std::sort(pairs.begin(), pairs.end());
    pairs.erase(std::unique(pairs.begin(), pairs.end()), pairs.end());
    P.cnt.assign((size_t) P.n_parts * P.n_parts, 0);
    P.halo_count.assign((size_t) P.n_parts, 0);
    for (const auto& pr : pairs) {
      P.cnt[(size_t) P.part[(size_t) pr.second] * P.n_parts + pr.first]++;
      P.halo_count[(size_t) pr.first]++;
    }
  }
  P.local.assign((size_t) P.n_parts, Local{});
  return SB_OK;
}

void build_local(const sb_part& P, int rank, Local& L) {
  const int64_t n = P.g.n_cells, F = P.g.n_faces, B = P.g.n_bfaces;
  const int32_t* part = P.part.data();
  const int32_t* fc = P.g.face_cell;
  // 1. classify owned cells, collect halo cells
  std::vector<uint8_t> is_boundary((size_t) n, 0), is_halo((size_t) n, 0);
  for (int64_t f = 0; f < F; ++f) {
    const int32_t a = fc[2 * f], b = fc[2 * f + 1];
    const bool oa = part[a] == rank, ob = part[b] == rank;
    if (oa && !ob) is_boundary[(size_t) a] = 1, is_halo[(size_t) b] = 1;
    if (ob && !oa) is_boundary[(size_t) b] = 1, is_halo[(size_t) a] = 1;
  }
  std::vector<int32_t> g2l((size_t) n, -1);
  L.l2g.clear();
  for (int64_t i = 0; i < n; ++i)
    if (part[i] == rank && !is_boundary[(size_t) i]) g2l[(size_t) i] = (int32_t) L.l2g.size(), L.l2g.push_back((int32_t) i);
  L.n_interior = (int64_t) L.l2g.size();
  for (int64_t i = 0; i < n; ++i)
    if (part[i] == rank && is_boundary[(size_t) i]) g2l[(size_t) i] = (int32_t) L.l2g.size(), L.l2g.push_back((int32_t) i);
  L.n_owned = (int64_t) L.l2g.size();
  L.halo_base = pad_up(L.n_owned);
  // halo: by owner rank ascending, then global id ascending
  std::vector<int32_t> halo;
  for (int64_t i = 0; i < n; ++i)
    if (is_halo[(size_t) i]) halo.push_back((int32_t) i);
  std::stable_sort(halo.begin(), halo.end(), [&](int32_t a, int32_t b) { return part[a] < part[b]; });
  L.n_halo = (int64_t) halo.size();
  L.nbr_rank.clear(), L.recv_ptr.assign(1, 0);
  for (int64_t h = 0; h < L.n_halo; ++h) {
    const int32_t owner = part[halo[(size_t) h]];
    if (L.nbr_rank.empty() || L.nbr_rank.back() != owner) {
      if (!L.nbr_rank.empty()) L.recv_ptr.push_back(h);
      L.nbr_rank.push_back(owner);
    }
    g2l[(size_t) halo[(size_t) h]] = (int32_t) (L.halo_base + h);
    L.l2g.push_back(halo[(size_t) h]);
  }
  if (!L.nbr_rank.empty()) L.recv_ptr.push_back(L.n_halo);
  // 2. send lists: for neighbour q, my owned cells adjacent to a cell of q, ascending global id
  const int nn = (int) L.nbr_rank.size();
  std::vector<int> slot_of_rank((size_t) P.n_parts, -1);
  for (int k = 0; k < nn; ++k) slot_of_rank[(size_t) L.nbr_rank[(size_t) k]] = k;
  std::vector<std::vector<int32_t>> send((size_t) nn);
  for (int64_t f = 0; f < F; ++f) {
    const int32_t a = fc[2 * f], b = fc[2 * f + 1];
    if (part[a] == rank && part[b] != rank) send[(size_t) slot_of_rank[(size_t) part[b]]].push_back(a);
    if (part[b] == rank && part[a] != rank) send[(size_t) slot_of_rank[(size_t) part[a]]].push_back(b);
  }
  L.send_ptr.assign(1, 0), L.send_idx.clear();
  for (int k = 0; k < nn; ++k) {
    auto& s = send[(size_t) k];
    std::sort(s.begin(), s.end());
    s.erase(std::unique(s.begin(), s.end()), s.end());
    for (int32_t gcell : s) L.send_idx.push_back(g2l[(size_t) gcell]);
    L.send_ptr.push_back((int64_t) L.send_idx.size());
  }
  // where my block starts inside neighbour q's vectors: its halo groups are ordered by owner rank
  L.send_dst.clear();
  for (int k = 0; k < nn; ++k) {
    const int q = L.nbr_rank[(size_t) k];
    int64_t off = pad_up(P.owned_count[(size_t) q]);
    for (int o = 0; o < rank; ++o) off += P.cnt[(size_t) o * P.n_parts + q];
    L.send_dst.push_back(off);
  }
  // 3. local face list (ascending global face index, global orientation), local geometry
  L.face_cell.clear(), L.face_area.clear(), L.face_dist.clear(), L.face_global.clear();
  for (int64_t f = 0; f < F; ++f) {
    const int32_t a = fc[2 * f], b = fc[2 * f + 1];
    if (part[a] != rank && part[b] != rank) continue;
    L.face_cell.push_back(g2l[(size_t) a]), L.face_cell.push_back(g2l[(size_t) b]);
    L.face_area.push_back(P.g.face_area[f]), L.face_dist.push_back(P.g.face_dist[f]);
    L.face_global.push_back(f);
  }
  const int64_t n_loc = L.halo_base + L.n_halo;
  L.cell_vol.assign((size_t) n_loc, 1.0); // padding cells: unit volume, never referenced by a face
  for (int64_t k = 0; k < L.n_owned; ++k) L.cell_vol[(size_t) k] = P.g.cell_vol[L.l2g[(size_t) k]];
  for (int64_t h = 0; h < L.n_halo; ++h)
    L.cell_vol[(size_t) (L.halo_base + h)] = P.g.cell_vol[L.l2g[(size_t) (L.n_owned + h)]];
  L.bface_cell.clear(), L.bface_area.clear(), L.bface_dist.clear(), L.bface_global.clear();
  for (int64_t b = 0; b < B; ++b) {
    const int32_t c = P.g.bface_cell[b];
    if (part[c] != rank) continue;
    L.bface_global.push_back(b);
    L.bface_cell.push_back(g2l[(size_t) c]);
    L.bface_area.push_back(P.g.bface_area[b]), L.bface_dist.push_back(P.g.bface_dist[b]);
  }
  L.built = true;
}

} // namespace

extern "C" {

int sb_part_from_array(const sb_mesh* mesh, int n_parts, const int32_t* h_part, sb_part** out) {
  SBP_REQUIRE(mesh != nullptr && h_part != nullptr && out != nullptr, "null argument");
  SBP_REQUIRE(n_parts >= 1, "n_parts must be >= 1");
  *out = nullptr;
  std::unique_ptr<sb_part> P(new sb_part());
  P->n_parts = n_parts;
  const int rc0 = sb_mesh_get_soa(mesh, &P->g);
  if (rc0 != SB_OK) return rc0;
  P->part.assign(h_part, h_part + P->g.n_cells);
  const int rc = finish(*P);
  if (rc != SB_OK) return rc;
  *out = P.release();
  return SB_OK;
}

int sb_part_create(const sb_mesh* mesh, int n_parts, int method, sb_part** out) {
  SBP_REQUIRE(mesh != nullptr && out != nullptr, "null argument");
  SBP_REQUIRE(n_parts >= 1, "n_parts must be >= 1");
  SBP_REQUIRE(method == SB_PART_METIS || method == SB_PART_SLAB, "unknown partition method");
  *out = nullptr;
  sb_mesh_soa g{};
  const int rc0 = sb_mesh_get_soa(mesh, &g);
  if (rc0 != SB_OK) return rc0;
  const int64_t n = g.n_cells, F = g.n_faces;
  SBP_REQUIRE(n_parts <= n, "more parts than cells");
  std::vector<int32_t> part((size_t) n, 0);
  if (n_parts == 1) {
    // nothing to do
  } else if (method == SB_PART_SLAB) {
    // contiguous blocks of the current (RCM) cell order: part p = cells [n*p/P, n*(p+1)/P)
    for (int p = 0; p < n_parts; ++p)
      for (int64_t i = n * p / n_parts; i < n * (p + 1) / n_parts; ++i) part[(size_t) i] = p;
  } else {
    // CSR cell graph, neighbours of a cell in ascending face index (the order METIS sees is part of
    // the determinism contract)
    std::vector<int64_t> xadj((size_t) n + 1, 0);
    for (int64_t f = 0; f < 2 * F; ++f) xadj[(size_t) g.face_cell[f] + 1]++;
    for (int64_t i = 0; i < n; ++i) xadj[(size_t) i + 1] += xadj[(size_t) i];
    std::vector<int64_t> adj((size_t) (2 * F)), fill(xadj.begin(), xadj.end() - 1);
    for (int64_t f = 0; f < F; ++f) {
      const int32_t a = g.face_cell[2 * f], b = g.face_cell[2 * f + 1];
      adj[(size_t) fill[(size_t) a]++] = b;
      adj[(size_t) fill[(size_t) b]++] = a;
    }
    int64_t options[40];
    METIS_SetDefaultOptions(options);
    int64_t nv = n, ncon = 1, np = n_parts, objval = 0;
    std::vector<int64_t> p64((size_t) n, 0);
    const int rc = METIS_PartGraphKway(&nv, &ncon, xadj.data(), adj.data(), nullptr, nullptr, nullptr, &np, nullptr,
                                       nullptr, options, &objval, p64.data());
    if (rc != 1) {
      sb::set_error("METIS_PartGraphKway failed with status %d", rc);
      return SB_ERR_INVALID;
    }
    for (int64_t i = 0; i < n; ++i) part[(size_t) i] = (int32_t) p64[(size_t) i];
    // METIS may leave a part empty on very small or edge-less graphs: every rank needs cells, so fall back to
    // contiguous slabs there (deterministic as well)
    std::vector<int64_t> count((size_t) n_parts, 0);
    for (int64_t i = 0; i < n; ++i) count[(size_t) part[(size_t) i]]++;
    if (*std::min_element(count.begin(), count.end()) == 0)
      for (int p = 0; p < n_parts; ++p)
        for (int64_t i = n * p / n_parts; i < n * (p + 1) / n_parts; ++i) part[(size_t) i] = p;
  }
  return sb_part_from_array(mesh, n_parts, part.data(), out);
}

int sb_part_destroy(sb_part* part) {
  delete part;
  return SB_OK;
}

int sb_part_get_array(const sb_part* part, const int32_t** h_part) {
  SBP_REQUIRE(part != nullptr && h_part != nullptr, "null argument");
  *h_part = part->part.data();
  return SB_OK;
}

int sb_part_local(sb_part* P, int rank, sb_local_mesh* out) {
  SBP_REQUIRE(P != nullptr && out != nullptr, "null argument");
  SBP_REQUIRE(rank >= 0 && rank < P->n_parts, "rank out of range");
  Local& L = P->local[(size_t) rank];
  if (!L.built) build_local(*P, rank, L);
  out->rank = rank, out->n_parts = P->n_parts;
  out->n_owned = L.n_owned, out->n_interior = L.n_interior, out->n_halo = L.n_halo, out->halo_base = L.halo_base;
  out->local_to_global = L.l2g.data();
  out->soa.n_cells = L.halo_base + L.n_halo;
  out->soa.n_faces = (int64_t) L.face_area.size();
  out->soa.face_cell = L.face_cell.data(), out->soa.face_area = L.face_area.data(), out->soa.face_dist = L.face_dist.data();
  out->soa.cell_vol = L.cell_vol.data();
  out->soa.n_bfaces = (int64_t) L.bface_area.size();
  out->soa.bface_cell = L.bface_cell.data(), out->soa.bface_area = L.bface_area.data();
  out->soa.bface_dist = L.bface_dist.data();
  out->face_global = L.face_global.data();
  out->n_nbr = (int32_t) L.nbr_rank.size();
  out->nbr_rank = L.nbr_rank.data();
  out->send_ptr = L.send_ptr.data(), out->send_idx = L.send_idx.data(), out->recv_ptr = L.recv_ptr.data();
  out->send_dst = L.send_dst.data();
  out->bface_global = L.bface_global.data();
  return SB_OK;
}

int sb_part_get_info(sb_part* P, sb_part_info* info) {
  SBP_REQUIRE(P != nullptr && info != nullptr, "null argument");
  info->n_parts = P->n_parts, info->n_cells = P->g.n_cells, info->edge_cut = P->edge_cut;
  info->max_owned = *std::max_element(P->owned_count.begin(), P->owned_count.end());
  info->min_owned = *std::min_element(P->owned_count.begin(), P->owned_count.end());
  const std::vector<int64_t>& halo = P->halo_count;
  info->max_halo = *std::max_element(halo.begin(), halo.end());
  int64_t cap = 0;
  for (int r = 0; r < P->n_parts; ++r)
    cap = std::max(cap, pad_up(P->owned_count[(size_t) r]) + pad_up(std::max<int64_t>(halo[(size_t) r], 1)));
  info->vec_capacity = cap;
  return SB_OK;
}

} // extern "C"
