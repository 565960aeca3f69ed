// sb_mesh_host.cpp -- host-side mesh ingestion for the hot path: cell soup -> face-list SoA,
// synthetic box meshes, RCM renumbering. Pure host C++ (no CUDA calls): usable without a GPU.
//
// The reference's mesh layer (source/Storm/Mallard) is 2-D only at this commit and
// inserts cells one by one with sorted-adjacency searches (~16 s for 80 K cells); this
// file keeps its conventions (see include/stormb200.h "mesh ingestion") but builds the face list for
// 10 M cells with one parallel sort. Every floating-point expression is written out operation by
// operation (compiled with -ffp-contract=off) so the independent C restatement in
// oracle/sb_oracle_mesh.c reproduces the arrays bit for bit.
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <numeric>
#include <random>
#include <fstream>
#include <iomanip>
#include <limits>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/stormb200.h"

namespace sb {
void set_error(const char* fmt, ...);
}

#define SBM_REQUIRE(cond, msg)                                                      \
  do {                                                                              \
    if (!(cond)) {                                                                  \
      ::sb::set_error("%s:%d: invalid argument: %s", __FILE__, __LINE__, msg);      \
      return SB_ERR_INVALID;                                                        \
    }                                                                               \
  } while (0)

namespace {

// Local faces, 0-based, in the reference's order (Mallard/Shape.hpp: Tetrahedron::faces :589-592,
// Hexahedron::faces :833-837). Triangles repeat -1 in the 4th slot.
constexpr int kTetFaces[4][4] = {{0, 2, 1, -1}, {0, 1, 3, -1}, {1, 2, 3, -1}, {2, 0, 3, -1}};
constexpr int kHexFaces[6][4] = {{0, 3, 2, 1}, {0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {0, 4, 7, 3}, {4, 5, 6, 7}};
// Hexahedron::pieces (Shape.hpp:845-852)
constexpr int kHexPieces[5][4] = {{0, 3, 1, 4}, {3, 2, 1, 6}, {4, 5, 6, 1}, {4, 6, 7, 3}, {4, 3, 1, 6}};

struct V3 {
  double x, y, z;
};
inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 scale(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 divs(V3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double dot3(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
// norm_2: sqrt of the sequential sum of squares from 0.0 (MatrixAlgorithms.hpp:262-270)
inline double length(V3 a) { return std::sqrt(((0.0 + a.x * a.x) + a.y * a.y) + a.z * a.z); }

inline double tri_area(V3 v1, V3 v2, V3 v3) { return length(cross(sub(v2, v1), sub(v3, v1))) / 2.0; }
inline V3 tri_center(V3 v1, V3 v2, V3 v3) { return divs(add(add(v1, v2), v3), 3.0); }
inline double tet_volume(V3 v1, V3 v2, V3 v3, V3 v4) {
  return std::fabs(dot3(sub(v2, v1), cross(sub(v3, v1), sub(v4, v1)))) / 6.0;
}
inline V3 tet_center(V3 v1, V3 v2, V3 v3, V3 v4) { return divs(add(add(add(v1, v2), v3), v4), 4.0); }

template<class It, class Cmp>
void parallel_sort(It first, It last, Cmp cmp) {
  const size_t n = (size_t) (last - first);
  unsigned hw = std::thread::hardware_concurrency();
  unsigned T = 1;
  while (T * 2 <= std::min(hw ? hw : 1u, 32u) && n / (T * 2) >= (1u << 16)) T *= 2;
  if (T == 1) {
    std::sort(first, last, cmp);
    return;
  }
  std::vector<size_t> cut(T + 1);
  for (unsigned t = 0; t <= T; ++t) cut[t] = n * t / T;
  {
    std::vector<std::thread> th;
    for (unsigned t = 0; t < T; ++t) th.emplace_back([&, t] { std::sort(first + cut[t], first + cut[t + 1], cmp); });
    for (auto& x : th) x.join();
  }
  for (unsigned width = 1; width < T; width *= 2) {
    std::vector<std::thread> th;
    for (unsigned t = 0; t + width < T; t += 2 * width)
      th.emplace_back([&, t, width] {
        std::inplace_merge(first + cut[t], first + cut[t + width], first + cut[std::min(t + 2 * width, T)], cmp);
      });
    for (auto& x : th) x.join();
  }
}

template<class F>
void parallel_for(int64_t n, F f) {
  unsigned hw = std::thread::hardware_concurrency();
  const unsigned T = (unsigned) std::max<int64_t>(1, std::min<int64_t>(std::min(hw ? hw : 1u, 32u), n / 65536));
  if (T <= 1) {
    f(0, n);
    return;
  }
  std::vector<std::thread> th;
  for (unsigned t = 0; t < T; ++t) th.emplace_back([&, t] { f(n * t / T, n * (t + 1) / T); });
  for (auto& x : th) x.join();
}

struct FacePair { // one face of the mesh: its (up to) two (cell, local face) sides
  int32_t cell[2];
  int8_t lf[2];
};

} // namespace

struct sb_mesh {
  int kind = SB_CELL_TET;
  int npc = 4, nfc = 4; // nodes / faces per cell
  int64_t n_nodes = 0, n_cells = 0;
  std::vector<double> xyz;
  std::vector<int32_t> cells;
  std::vector<FacePair> pairs; // matched faces, numbering-independent except for the cell ids
  std::vector<int64_t> face_order; // ordered face q (interior first, then boundary) -> index into pairs
  // derived SoA
  std::vector<double> cell_vol, cell_ctr;
  std::vector<int32_t> face_cell, bface_cell;
  std::vector<double> face_area, face_dist, bface_area, bface_dist;

  // face-list meshes (SB_CELL_FACELIST, sb_mesh_from_faces): no nodes; the geometry is carried per face,
  // indexed like `pairs`. pair_normal (optional, 3 per face) is oriented from pair_first to the other side.
  std::vector<double> pair_area, pair_dist, pair_normal;
  std::vector<int32_t> pair_first; // the cell the stored normal points away from, as an index into cell[] (0 | 1)
  std::vector<int32_t> pair_label; // optional: the reference's face labels (0 = interior), sb_mesh_read_tetgen_2d
  bool has_centers = false;

  V3 node(int32_t i) const { return {xyz[3 * (size_t) i], xyz[3 * (size_t) i + 1], xyz[3 * (size_t) i + 2]}; }
  const int* local_face(int lf) const { return kind == SB_CELL_TET ? kTetFaces[lf] : kHexFaces[lf]; }
};

namespace {

void compute_cell_geometry(sb_mesh& m) {
  m.cell_vol.assign((size_t) m.n_cells, 0.0);
  m.cell_ctr.assign(3 * (size_t) m.n_cells, 0.0);
  parallel_for(m.n_cells, [&](int64_t lo, int64_t hi) {
    for (int64_t c = lo; c < hi; ++c) {
      const int32_t* nd = &m.cells[(size_t) c * m.npc];
      double vol;
      V3 ctr;
      if (m.kind == SB_CELL_TET) {
        const V3 v1 = m.node(nd[0]), v2 = m.node(nd[1]), v3 = m.node(nd[2]), v4 = m.node(nd[3]);
        vol = tet_volume(v1, v2, v3, v4);
        ctr = tet_center(v1, v2, v3, v4);
      } else {
        // complex shape: volume = sum over pieces, barycentre volume-weighted (Shape.hpp:170-215)
        V3 vc{0, 0, 0};
        vol = 0.0;
        for (int p = 0; p < 5; ++p) {
          const V3 v1 = m.node(nd[kHexPieces[p][0]]), v2 = m.node(nd[kHexPieces[p][1]]);
          const V3 v3 = m.node(nd[kHexPieces[p][2]]), v4 = m.node(nd[kHexPieces[p][3]]);
          const double dv = tet_volume(v1, v2, v3, v4);
          const V3 w = scale(dv, tet_center(v1, v2, v3, v4));
          if (p == 0) vol = dv, vc = w;
          else vol += dv, vc = add(vc, w);
        }
        ctr = divs(vc, vol);
      }
      m.cell_vol[(size_t) c] = vol;
      m.cell_ctr[3 * (size_t) c] = ctr.x, m.cell_ctr[3 * (size_t) c + 1] = ctr.y, m.cell_ctr[3 * (size_t) c + 2] = ctr.z;
    }
  });
}

// Stage 1: match the cell faces by their sorted node tuples.
int match_faces(sb_mesh& m) {
  struct Rec {
    uint64_t hi, lo;
    int32_t cell;
    int8_t lf;
  };
  const int64_t n_rec = m.n_cells * m.nfc;
  std::vector<Rec> rec((size_t) n_rec);
  parallel_for(m.n_cells, [&](int64_t lo, int64_t hi) {
    for (int64_t c = lo; c < hi; ++c) {
      const int32_t* nd = &m.cells[(size_t) c * m.npc];
      for (int f = 0; f < m.nfc; ++f) {
        const int* lf = m.local_face(f);
        uint32_t k[4] = {(uint32_t) nd[lf[0]], (uint32_t) nd[lf[1]], (uint32_t) nd[lf[2]],
                         lf[3] >= 0 ? (uint32_t) nd[lf[3]] : 0xFFFFFFFFu};
        std::sort(k, k + 4);
        rec[(size_t) (c * m.nfc + f)] = Rec{((uint64_t) k[0] << 32) | k[1], ((uint64_t) k[2] << 32) | k[3], (int32_t) c,
                                            (int8_t) f};
      }
    }
  });
  parallel_sort(rec.begin(), rec.end(), [](const Rec& a, const Rec& b) {
    if (a.hi != b.hi) return a.hi < b.hi;
    if (a.lo != b.lo) return a.lo < b.lo;
    return a.cell < b.cell;
  });
  m.pairs.clear();
  m.pairs.reserve((size_t) n_rec / 2 + 1024);
  for (int64_t i = 0; i < n_rec;) {
    int64_t j = i + 1;
    while (j < n_rec && rec[(size_t) j].hi == rec[(size_t) i].hi && rec[(size_t) j].lo == rec[(size_t) i].lo) ++j;
    if (j - i > 2) {
      sb::set_error("non-manifold mesh: a face is shared by %lld cells", (long long) (j - i));
      return SB_ERR_INVALID;
    }
    FacePair p;
    p.cell[0] = rec[(size_t) i].cell, p.lf[0] = rec[(size_t) i].lf;
    if (j - i == 2) p.cell[1] = rec[(size_t) i + 1].cell, p.lf[1] = rec[(size_t) i + 1].lf;
    else p.cell[1] = -1, p.lf[1] = -1;
    m.pairs.push_back(p);
    i = j;
  }
  return SB_OK;
}

// Stage 2: order the faces by creation (first cell in cell order, then local face), interior first,
// fix inner = creator, and evaluate the face geometry.
// Face-list meshes: same ordering rule (creating cell = the lower cell id, then that cell's local face), the
// geometry is the stored per-face payload instead of being evaluated from nodes.
void order_faces_facelist(sb_mesh& m) {
  const int64_t nf = (int64_t) m.pairs.size();
  std::vector<uint64_t> key((size_t) nf);
  int64_t n_int = 0;
  for (int64_t f = 0; f < nf; ++f) {
    FacePair& p = m.pairs[(size_t) f];
    if (p.cell[1] >= 0 && p.cell[1] < p.cell[0]) {
      std::swap(p.cell[0], p.cell[1]), std::swap(p.lf[0], p.lf[1]);
      m.pair_first[(size_t) f] ^= 1;
    }
    n_int += p.cell[1] >= 0;
    key[(size_t) f] = ((uint64_t) (p.cell[1] < 0) << 62) | ((uint64_t) p.cell[0] * 256u + (uint64_t) (uint8_t) p.lf[0]);
  }
  std::vector<int64_t> idx((size_t) nf);
  std::iota(idx.begin(), idx.end(), 0);
  parallel_sort(idx.begin(), idx.end(), [&](int64_t a, int64_t b) { return key[(size_t) a] < key[(size_t) b]; });
  const int64_t n_b = nf - n_int;
  m.face_cell.assign(2 * (size_t) n_int, 0);
  m.face_area.assign((size_t) n_int, 0.0), m.face_dist.assign((size_t) n_int, 0.0);
  m.bface_cell.assign((size_t) n_b, 0);
  m.bface_area.assign((size_t) n_b, 0.0), m.bface_dist.assign((size_t) n_b, 0.0);
  for (int64_t q = 0; q < nf; ++q) {
    const size_t f = (size_t) idx[(size_t) q];
    const FacePair& p = m.pairs[f];
    if (q < n_int) {
      m.face_cell[2 * (size_t) q] = p.cell[0], m.face_cell[2 * (size_t) q + 1] = p.cell[1];
      m.face_area[(size_t) q] = m.pair_area[f], m.face_dist[(size_t) q] = m.pair_dist[f];
    } else {
      const size_t b = (size_t) (q - n_int);
      m.bface_cell[b] = p.cell[0], m.bface_area[b] = m.pair_area[f], m.bface_dist[b] = m.pair_dist[f];
    }
  }
  m.face_order.swap(idx);
}

void order_faces(sb_mesh& m) {
  if (m.kind == SB_CELL_FACELIST) return order_faces_facelist(m);
  const int64_t nf = (int64_t) m.pairs.size();
  std::vector<uint64_t> key((size_t) nf); // (creator cell * 8 + local face) << 1 ... sorted ascending
  int64_t n_int = 0;
  for (int64_t f = 0; f < nf; ++f) {
    FacePair& p = m.pairs[(size_t) f];
    if (p.cell[1] >= 0 && p.cell[1] < p.cell[0]) std::swap(p.cell[0], p.cell[1]), std::swap(p.lf[0], p.lf[1]);
    n_int += p.cell[1] >= 0;
  }
  // interior faces first (label 0), each group by creation order
  std::vector<int64_t> idx((size_t) nf);
  std::iota(idx.begin(), idx.end(), 0);
  for (int64_t f = 0; f < nf; ++f) {
    const FacePair& p = m.pairs[(size_t) f];
    key[(size_t) f] = ((uint64_t) (p.cell[1] < 0) << 62) | ((uint64_t) p.cell[0] * 8u + (uint64_t) p.lf[0]);
  }
  parallel_sort(idx.begin(), idx.end(), [&](int64_t a, int64_t b) { return key[(size_t) a] < key[(size_t) b]; });
  const int64_t n_b = nf - n_int;
  m.face_cell.assign(2 * (size_t) n_int, 0);
  m.face_area.assign((size_t) n_int, 0.0), m.face_dist.assign((size_t) n_int, 0.0);
  m.bface_cell.assign((size_t) n_b, 0);
  m.bface_area.assign((size_t) n_b, 0.0), m.bface_dist.assign((size_t) n_b, 0.0);
  parallel_for(nf, [&](int64_t lo, int64_t hi) {
    for (int64_t q = lo; q < hi; ++q) {
      const FacePair& p = m.pairs[(size_t) idx[(size_t) q]];
      const int32_t ci = p.cell[0];
      const int32_t* nd = &m.cells[(size_t) ci * m.npc];
      const int* lf = m.local_face(p.lf[0]);
      double area;
      V3 fc;
      const V3 v1 = m.node(nd[lf[0]]), v2 = m.node(nd[lf[1]]), v3 = m.node(nd[lf[2]]);
      if (lf[3] < 0) {
        area = tri_area(v1, v2, v3);
        fc = tri_center(v1, v2, v3);
      } else {
        // Quadrangle::pieces: (n1,n2,n3) + (n3,n4,n1)  (Shape.hpp:395-402)
        const V3 v4 = m.node(nd[lf[3]]);
        const double a1 = tri_area(v1, v2, v3), a2 = tri_area(v3, v4, v1);
        area = a1 + a2;
        fc = divs(add(scale(a1, tri_center(v1, v2, v3)), scale(a2, tri_center(v3, v4, v1))), area);
      }
      const V3 xi{m.cell_ctr[3 * (size_t) ci], m.cell_ctr[3 * (size_t) ci + 1], m.cell_ctr[3 * (size_t) ci + 2]};
      if (q < n_int) {
        const int32_t co = p.cell[1];
        const V3 xo{m.cell_ctr[3 * (size_t) co], m.cell_ctr[3 * (size_t) co + 1], m.cell_ctr[3 * (size_t) co + 2]};
        m.face_cell[2 * (size_t) q] = ci, m.face_cell[2 * (size_t) q + 1] = co;
        m.face_area[(size_t) q] = area;
        m.face_dist[(size_t) q] = length(sub(xo, xi)); // Playground.cpp:126
      } else {
        const size_t b = (size_t) (q - n_int);
        m.bface_cell[b] = ci;
        m.bface_area[b] = area;
        m.bface_dist[b] = 2.0 * length(sub(fc, xi)); // mirror ghost centre
      }
    }
  });
  m.face_order.swap(idx);
}

// Unit face normals, oriented inner -> outer (boundary faces: out of the domain), evaluated on demand.
// Triangle: cross(v2-v1, v3-v1); quadrangle: cross of the diagonals (v3-v1) x (v4-v2); normalised by
// its length; flipped when it points against (x_outer - x_inner) / (face centre - x_inner).
void face_normals(const sb_mesh& m, double* fn, double* bn) {
  const int64_t nf = (int64_t) m.face_order.size(), n_int = (int64_t) m.face_area.size();
  if (m.kind == SB_CELL_FACELIST) {
    // stored normals point away from pair_first; the inner cell is cell[0]
    for (int64_t q = 0; q < nf; ++q) {
      const size_t f = (size_t) m.face_order[(size_t) q];
      const double sgn = m.pair_first[f] == 0 ? 1.0 : -1.0;
      double* base = q < n_int ? fn : bn;
      if (base == nullptr) continue;
      double* out = base + 3 * (q < n_int ? q : q - n_int);
      for (int c = 0; c < 3; ++c) out[c] = sgn * m.pair_normal[3 * f + (size_t) c];
    }
    return;
  }
  parallel_for(nf, [&](int64_t lo, int64_t hi) {
    for (int64_t q = lo; q < hi; ++q) {
      const FacePair& p = m.pairs[(size_t) m.face_order[(size_t) q]];
      const int32_t ci = p.cell[0];
      const int32_t* nd = &m.cells[(size_t) ci * m.npc];
      const int* lf = m.local_face(p.lf[0]);
      const V3 v1 = m.node(nd[lf[0]]), v2 = m.node(nd[lf[1]]), v3 = m.node(nd[lf[2]]);
      V3 nrm, fc;
      if (lf[3] < 0) {
        nrm = cross(sub(v2, v1), sub(v3, v1));
        fc = tri_center(v1, v2, v3);
      } else {
        const V3 v4 = m.node(nd[lf[3]]);
        nrm = cross(sub(v3, v1), sub(v4, v2));
        const double a1 = tri_area(v1, v2, v3), a2 = tri_area(v3, v4, v1);
        fc = divs(add(scale(a1, tri_center(v1, v2, v3)), scale(a2, tri_center(v3, v4, v1))), a1 + a2);
      }
      nrm = divs(nrm, length(nrm));
      const V3 xi{m.cell_ctr[3 * (size_t) ci], m.cell_ctr[3 * (size_t) ci + 1], m.cell_ctr[3 * (size_t) ci + 2]};
      V3 dir;
      if (q < n_int) {
        const int32_t co = p.cell[1];
        dir = sub(V3{m.cell_ctr[3 * (size_t) co], m.cell_ctr[3 * (size_t) co + 1], m.cell_ctr[3 * (size_t) co + 2]}, xi);
      } else {
        dir = sub(fc, xi);
      }
      if (dot3(nrm, dir) < 0.0) nrm = V3{-nrm.x, -nrm.y, -nrm.z};
      double* base = q < n_int ? fn : bn;
      if (base != nullptr) {
        double* out = base + 3 * (q < n_int ? q : q - n_int);
        out[0] = nrm.x, out[1] = nrm.y, out[2] = nrm.z;
      }
    }
  });
}

int validate_cells(const sb_mesh& m) {
  for (size_t i = 0; i < m.cells.size(); ++i)
    SBM_REQUIRE(m.cells[i] >= 0 && m.cells[i] < m.n_nodes, "cell node index out of range");
  return SB_OK;
}

int derive(sb_mesh& m) {
  compute_cell_geometry(m);
  const int rc = match_faces(m);
  if (rc != SB_OK) return rc;
  order_faces(m);
  return SB_OK;
}

int apply_permutation(sb_mesh& m, const int32_t* perm) {
  const int64_t n = m.n_cells;
  std::vector<int32_t> iperm((size_t) n, -1);
  for (int64_t k = 0; k < n; ++k) {
    SBM_REQUIRE(perm[k] >= 0 && perm[k] < n && iperm[(size_t) perm[k]] < 0, "perm is not a permutation");
    iperm[(size_t) perm[k]] = (int32_t) k;
  }
  std::vector<int32_t> cells(m.cells.size());
  std::vector<double> vol((size_t) n), ctr(3 * (size_t) n);
  parallel_for(n, [&](int64_t lo, int64_t hi) {
    for (int64_t k = lo; k < hi; ++k) {
      const size_t o = (size_t) perm[k];
      if (m.npc > 0) std::memcpy(&cells[(size_t) k * m.npc], &m.cells[o * m.npc], sizeof(int32_t) * m.npc);
      vol[(size_t) k] = m.cell_vol[o];
      if (!m.cell_ctr.empty())
        ctr[3 * (size_t) k] = m.cell_ctr[3 * o], ctr[3 * (size_t) k + 1] = m.cell_ctr[3 * o + 1], ctr[3 * (size_t) k + 2] = m.cell_ctr[3 * o + 2];
    }
  });
  m.cells.swap(cells), m.cell_vol.swap(vol);
  if (!m.cell_ctr.empty()) m.cell_ctr.swap(ctr);
  for (FacePair& p : m.pairs) {
    p.cell[0] = iperm[(size_t) p.cell[0]];
    if (p.cell[1] >= 0) p.cell[1] = iperm[(size_t) p.cell[1]];
  }
  order_faces(m);
  return SB_OK;
}

// Reverse Cuthill-McKee. Deterministic rules (restated in oracle/sb_oracle_mesh.c):
//  * graph: cells, one edge per interior face; degree = number of interior faces of the cell;
//  * components are started in order of the smallest (degree, id) among unvisited cells; that seed
//    is refined once: BFS from it, take the smallest (degree, id) cell of the last level as root;
//  * BFS from the root; the unvisited neighbours of a dequeued cell are appended in ascending
//    (degree, id); the concatenated order of all components is reversed at the end.
void rcm_order(const sb_mesh& m, std::vector<int32_t>& perm) {
  const int64_t n = m.n_cells, F = (int64_t) m.face_area.size();
  std::vector<int64_t> ptr((size_t) n + 1, 0);
  for (int64_t f = 0; f < 2 * F; ++f) ptr[(size_t) m.face_cell[(size_t) f] + 1]++;
  for (int64_t i = 0; i < n; ++i) ptr[(size_t) i + 1] += ptr[(size_t) i];
  std::vector<int32_t> adj((size_t) (2 * F));
  {
    std::vector<int64_t> fill(ptr.begin(), ptr.end() - 1);
    for (int64_t f = 0; f < F; ++f) {
      const int32_t a = m.face_cell[2 * (size_t) f], b = m.face_cell[2 * (size_t) f + 1];
      adj[(size_t) fill[(size_t) a]++] = b;
      adj[(size_t) fill[(size_t) b]++] = a;
    }
  }
  auto deg = [&](int32_t v) { return (int32_t) (ptr[(size_t) v + 1] - ptr[(size_t) v]); };
  auto less = [&](int32_t a, int32_t b) { return deg(a) != deg(b) ? deg(a) < deg(b) : a < b; };
  std::vector<int32_t> by_deg((size_t) n);
  std::iota(by_deg.begin(), by_deg.end(), 0);
  std::stable_sort(by_deg.begin(), by_deg.end(), [&](int32_t a, int32_t b) { return deg(a) < deg(b); });
  std::vector<int32_t> order;
  order.reserve((size_t) n);
  std::vector<uint8_t> visited((size_t) n, 0);
  std::vector<int32_t> stamp((size_t) n, -1), queue, nb;
  int32_t epoch = 0;
  size_t seed_pos = 0;
  while ((int64_t) order.size() < n) {
    while (visited[(size_t) by_deg[seed_pos]]) ++seed_pos;
    const int32_t seed = by_deg[seed_pos];
    // one pseudo-peripheral refinement: last BFS level from the seed
    queue.clear();
    queue.push_back(seed);
    stamp[(size_t) seed] = epoch;
    size_t level_begin = 0, head = 0;
    while (head < queue.size()) {
      const size_t level_end = queue.size();
      level_begin = head;
      for (; head < level_end; ++head) {
        const int32_t v = queue[head];
        for (int64_t e = ptr[(size_t) v]; e < ptr[(size_t) v + 1]; ++e) {
          const int32_t w = adj[(size_t) e];
          if (!visited[(size_t) w] && stamp[(size_t) w] != epoch) stamp[(size_t) w] = epoch, queue.push_back(w);
        }
      }
    }
    ++epoch;
    int32_t root = queue[level_begin];
    for (size_t q = level_begin; q < queue.size(); ++q)
      if (less(queue[q], root)) root = queue[q];
    // Cuthill-McKee BFS from the root
    const size_t first = order.size();
    order.push_back(root);
    visited[(size_t) root] = 1;
    for (size_t h = first; h < order.size(); ++h) {
      const int32_t v = order[h];
      nb.clear();
      for (int64_t e = ptr[(size_t) v]; e < ptr[(size_t) v + 1]; ++e) {
        const int32_t w = adj[(size_t) e];
        if (!visited[(size_t) w]) visited[(size_t) w] = 1, nb.push_back(w);
      }
      std::sort(nb.begin(), nb.end(), less);
      order.insert(order.end(), nb.begin(), nb.end());
    }
  }
  perm.resize((size_t) n);
  for (int64_t k = 0; k < n; ++k) perm[(size_t) k] = order[(size_t) (n - 1 - k)];
}

} // namespace

extern "C" {

int sb_mesh_from_cells(int cell_kind, int64_t n_nodes, const double* h_xyz, int64_t n_cells,
                       const int32_t* h_cell_nodes, sb_mesh** out) {
  SBM_REQUIRE(out != nullptr, "out is null");
  *out = nullptr;
  SBM_REQUIRE(cell_kind == SB_CELL_TET || cell_kind == SB_CELL_HEX, "unknown cell kind");
  SBM_REQUIRE(n_nodes > 0 && n_cells > 0 && h_xyz != nullptr && h_cell_nodes != nullptr, "empty mesh");
  SBM_REQUIRE(n_cells < (int64_t) 250'000'000 && n_nodes < (int64_t) 0x7FFFFFFF, "mesh too large for int32 indices");
  std::unique_ptr<sb_mesh> m(new sb_mesh());
  m->kind = cell_kind;
  m->npc = cell_kind == SB_CELL_TET ? 4 : 8, m->nfc = cell_kind == SB_CELL_TET ? 4 : 6;
  m->n_nodes = n_nodes, m->n_cells = n_cells;
  m->xyz.assign(h_xyz, h_xyz + 3 * n_nodes);
  m->cells.assign(h_cell_nodes, h_cell_nodes + n_cells * m->npc);
  int rc = validate_cells(*m);
  if (rc != SB_OK) return rc;
  rc = derive(*m);
  if (rc != SB_OK) return rc;
  *out = m.release();
  return SB_OK;
}

int sb_mesh_from_faces(const sb_mesh_soa* h, const double* h_cell_ctr, const double* h_face_normal,
                       const double* h_bface_normal, sb_mesh** out) {
  SBM_REQUIRE(out != nullptr, "out is null");
  *out = nullptr;
  SBM_REQUIRE(h != nullptr && h->n_cells > 0 && h->cell_vol != nullptr, "empty mesh");
  SBM_REQUIRE(h->n_cells < (int64_t) 0x7FFFFFFF - 4096, "mesh too large for int32 indices");
  SBM_REQUIRE(h->n_faces >= 0 && h->n_bfaces >= 0, "negative face count");
  SBM_REQUIRE(h->n_faces == 0 || (h->face_cell && h->face_area && h->face_dist), "null face arrays");
  SBM_REQUIRE(h->n_bfaces == 0 || (h->bface_cell && h->bface_area && h->bface_dist), "null boundary-face arrays");
  const int64_t n = h->n_cells, F = h->n_faces, B = h->n_bfaces;
  const bool normals = h_face_normal != nullptr || h_bface_normal != nullptr;
  SBM_REQUIRE(!normals || ((F == 0 || h_face_normal != nullptr) && (B == 0 || h_bface_normal != nullptr)),
              "give both normal arrays or neither");
  std::unique_ptr<sb_mesh> m(new sb_mesh());
  m->kind = SB_CELL_FACELIST, m->npc = 0, m->nfc = 0, m->n_nodes = 0, m->n_cells = n;
  m->cell_vol.assign(h->cell_vol, h->cell_vol + n);
  if (h_cell_ctr != nullptr) m->cell_ctr.assign(h_cell_ctr, h_cell_ctr + 3 * n), m->has_centers = true;
  // local face index of a face in a cell = its ordinal among that cell's faces in the given order (interior faces
  // first, then boundary faces): a property of the cell, independent of any later renumbering
  std::vector<int32_t> count((size_t) n, 0);
  m->pairs.resize((size_t) (F + B));
  m->pair_area.resize((size_t) (F + B)), m->pair_dist.resize((size_t) (F + B));
  m->pair_first.assign((size_t) (F + B), 0);
  if (normals) m->pair_normal.resize(3 * (size_t) (F + B));
  for (int64_t f = 0; f < F; ++f) {
    const int32_t a = h->face_cell[2 * f], b = h->face_cell[2 * f + 1];
    SBM_REQUIRE(a >= 0 && a < n && b >= 0 && b < n && a != b, "face_cell index out of range");
    SBM_REQUIRE(count[(size_t) a] < 127 && count[(size_t) b] < 127, "more than 127 faces on a cell");
    FacePair& p = m->pairs[(size_t) f];
    p.cell[0] = a, p.cell[1] = b;
    p.lf[0] = (int8_t) count[(size_t) a]++, p.lf[1] = (int8_t) count[(size_t) b]++;
    m->pair_area[(size_t) f] = h->face_area[f], m->pair_dist[(size_t) f] = h->face_dist[f];
    if (normals) std::memcpy(&m->pair_normal[3 * (size_t) f], h_face_normal + 3 * f, 3 * sizeof(double));
  }
  for (int64_t q = 0; q < B; ++q) {
    const int32_t a = h->bface_cell[q];
    SBM_REQUIRE(a >= 0 && a < n, "bface_cell index out of range");
    SBM_REQUIRE(count[(size_t) a] < 127, "more than 127 faces on a cell");
    FacePair& p = m->pairs[(size_t) (F + q)];
    p.cell[0] = a, p.cell[1] = -1;
    p.lf[0] = (int8_t) count[(size_t) a]++, p.lf[1] = -1;
    m->pair_area[(size_t) (F + q)] = h->bface_area[q], m->pair_dist[(size_t) (F + q)] = h->bface_dist[q];
    if (normals) std::memcpy(&m->pair_normal[3 * (size_t) (F + q)], h_bface_normal + 3 * q, 3 * sizeof(double));
  }
  // the face list is taken verbatim (order and inner/outer as given: "the reference's face order"); only a later
  // renumbering re-derives the order by the creation rule
  m->face_cell.assign(h->face_cell, h->face_cell + 2 * F);
  m->face_area.assign(h->face_area, h->face_area + F), m->face_dist.assign(h->face_dist, h->face_dist + F);
  m->bface_cell.assign(h->bface_cell, h->bface_cell + B);
  m->bface_area.assign(h->bface_area, h->bface_area + B), m->bface_dist.assign(h->bface_dist, h->bface_dist + B);
  m->face_order.resize((size_t) (F + B));
  std::iota(m->face_order.begin(), m->face_order.end(), 0);
  *out = m.release();
  return SB_OK;
}

int sb_mesh_generate_box(int cell_kind, int nx, int ny, int nz, double jitter, uint64_t seed_jitter, int shuffle,
                         uint64_t seed_shuffle, sb_mesh** out) {
  SBM_REQUIRE(out != nullptr, "out is null");
  *out = nullptr;
  SBM_REQUIRE(cell_kind == SB_CELL_TET || cell_kind == SB_CELL_HEX, "unknown cell kind");
  SBM_REQUIRE(nx >= 1 && ny >= 1 && nz >= 1, "box dimensions must be >= 1");
  SBM_REQUIRE(jitter >= 0.0 && jitter < 0.5, "jitter must be in [0, 0.5)");
  const int64_t per_hex = cell_kind == SB_CELL_TET ? 6 : 1;
  const int64_t n_hex = (int64_t) nx * ny * nz, n_cells = n_hex * per_hex;
  const int64_t n_nodes = (int64_t) (nx + 1) * (ny + 1) * (nz + 1);
  SBM_REQUIRE(n_cells < (int64_t) 250'000'000, "mesh too large");
  std::vector<double> xyz(3 * (size_t) n_nodes);
  const double hx = 1.0 / nx, hy = 1.0 / ny, hz = 1.0 / nz;
  std::mt19937_64 eng(seed_jitter);
  std::uniform_real_distribution<double> ux(-jitter * hx, jitter * hx), uy(-jitter * hy, jitter * hy),
      uz(-jitter * hz, jitter * hz);
  auto nid = [&](int i, int j, int k) { return (int32_t) (((int64_t) k * (ny + 1) + j) * (nx + 1) + i); };
  for (int k = 0; k <= nz; ++k)
    for (int j = 0; j <= ny; ++j)
      for (int i = 0; i <= nx; ++i) {
        double x = i * hx, y = j * hy, z = k * hz;
        const bool interior = i > 0 && i < nx && j > 0 && j < ny && k > 0 && k < nz;
        if (interior && jitter > 0.0) x += ux(eng), y += uy(eng), z += uz(eng);
        const size_t q = 3 * (size_t) nid(i, j, k);
        xyz[q] = x, xyz[q + 1] = y, xyz[q + 2] = z;
      }
  const int npc = cell_kind == SB_CELL_TET ? 4 : 8;
  std::vector<int32_t> cells((size_t) n_cells * npc);
  // Kuhn triangulation: the six monotone corner paths n1 -> n7 (translation invariant, hence conforming)
  static const int kuhn[6][4] = {{0, 1, 2, 6}, {0, 1, 5, 6}, {0, 3, 2, 6}, {0, 3, 7, 6}, {0, 4, 5, 6}, {0, 4, 7, 6}};
  int64_t c = 0;
  for (int k = 0; k < nz; ++k)
    for (int j = 0; j < ny; ++j)
      for (int i = 0; i < nx; ++i) {
        // hexahedron node order of Shape.hpp:803-818: n1..n4 bottom ring, n5..n8 top ring
        const int32_t h[8] = {nid(i, j, k),     nid(i + 1, j, k),     nid(i + 1, j + 1, k),     nid(i, j + 1, k),
                              nid(i, j, k + 1), nid(i + 1, j, k + 1), nid(i + 1, j + 1, k + 1), nid(i, j + 1, k + 1)};
        if (cell_kind == SB_CELL_HEX) {
          std::memcpy(&cells[(size_t) c * 8], h, sizeof(h));
          ++c;
        } else {
          for (int t = 0; t < 6; ++t, ++c)
            for (int q = 0; q < 4; ++q) cells[(size_t) c * 4 + q] = h[kuhn[t][q]];
        }
      }
  if (shuffle) {
    // Fisher-Yates with explicit modulo draws (std::shuffle is implementation-defined)
    std::mt19937_64 se(seed_shuffle);
    std::vector<int32_t> tmp(npc);
    for (int64_t i = n_cells - 1; i >= 1; --i) {
      const int64_t j = (int64_t) (se() % (uint64_t) (i + 1));
      if (j != i) {
        std::memcpy(tmp.data(), &cells[(size_t) i * npc], sizeof(int32_t) * npc);
        std::memcpy(&cells[(size_t) i * npc], &cells[(size_t) j * npc], sizeof(int32_t) * npc);
        std::memcpy(&cells[(size_t) j * npc], tmp.data(), sizeof(int32_t) * npc);
      }
    }
  }
  return sb_mesh_from_cells(cell_kind, n_nodes, xyz.data(), n_cells, cells.data(), out);
}

namespace {
// Token reader with '#' comments (to end of line), like the reference's FilteringStreambuf<char,'#','\n'>.
struct TokenFile {
  std::vector<char> buf;
  size_t pos = 0;
  bool open(const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (f == nullptr) return false;
    std::fseek(f, 0, SEEK_END);
    const long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    buf.resize((size_t) (n > 0 ? n : 0) + 1);
    const size_t got = std::fread(buf.data(), 1, buf.size() - 1, f);
    std::fclose(f);
    buf[got] = 0, buf.resize(got + 1);
    return true;
  }
  void skip() {
    for (;;) {
      while (pos + 1 < buf.size() && std::isspace((unsigned char) buf[pos])) ++pos;
      if (pos + 1 < buf.size() && buf[pos] == '#') {
        while (pos + 1 < buf.size() && buf[pos] != '\n') ++pos;
      } else {
        return;
      }
    }
  }
  bool next_double(double& v) {
    skip();
    if (pos + 1 >= buf.size()) return false;
    char* end = nullptr;
    v = std::strtod(buf.data() + pos, &end);
    if (end == buf.data() + pos) return false;
    pos = (size_t) (end - buf.data());
    return true;
  }
  bool next_int(int64_t& v) {
    skip();
    if (pos + 1 >= buf.size()) return false;
    char* end = nullptr;
    v = std::strtoll(buf.data() + pos, &end, 10);
    if (end == buf.data() + pos) return false;
    pos = (size_t) (end - buf.data());
    return true;
  }
};
} // namespace

int sb_mesh_read_tetgen(const char* path_prefix, sb_mesh** out) {
  SBM_REQUIRE(path_prefix != nullptr && out != nullptr, "null argument");
  *out = nullptr;
  const std::string prefix{path_prefix};
  TokenFile nf, ef;
  if (!nf.open(prefix + ".node")) {
    sb::set_error("cannot open the node file '%s.node'", path_prefix);
    return SB_ERR_INVALID;
  }
  if (!ef.open(prefix + ".ele")) {
    sb::set_error("cannot open the cell file '%s.ele'", path_prefix);
    return SB_ERR_INVALID;
  }
  int64_t n_nodes = 0, dim = 0, n_attr = 0, has_labels = 0;
  SBM_REQUIRE(nf.next_int(n_nodes) && nf.next_int(dim) && nf.next_int(n_attr) && nf.next_int(has_labels),
              "cannot read the node file header");
  SBM_REQUIRE(dim == 3, "unexpected number of dimensions in the node file header (expected 3)");
  SBM_REQUIRE(n_nodes > 0 && n_nodes < (int64_t) INT32_MAX && n_attr >= 0, "bad node file header");
  std::vector<double> xyz(3 * (size_t) n_nodes);
  int64_t first_index = 0;
  for (int64_t k = 0; k < n_nodes; ++k) {
    int64_t idx = 0;
    double skipped = 0.0;
    bool ok = nf.next_int(idx) && nf.next_double(xyz[3 * (size_t) k]) && nf.next_double(xyz[3 * (size_t) k + 1]) &&
              nf.next_double(xyz[3 * (size_t) k + 2]);
    for (int64_t a = 0; ok && a < n_attr + (has_labels ? 1 : 0); ++a) ok = nf.next_double(skipped);
    if (!ok) {
      sb::set_error("cannot read node # %lld from '%s.node'", (long long) k, path_prefix);
      return SB_ERR_INVALID;
    }
    if (k == 0) first_index = idx;
    SBM_REQUIRE(idx == first_index + k, "node indices must be consecutive");
  }
  SBM_REQUIRE(first_index == 0 || first_index == 1, "node numbering must start at 0 or 1");
  int64_t n_cells = 0, npc = 0, has_attr = 0;
  SBM_REQUIRE(ef.next_int(n_cells) && ef.next_int(npc) && ef.next_int(has_attr), "cannot read the cell file header");
  SBM_REQUIRE(npc == 4, "unexpected number of nodes per cell in the cell file header (expected 4)");
  SBM_REQUIRE(n_cells > 0 && n_cells < (int64_t) 250'000'000, "bad cell count");
  std::vector<int32_t> cells(4 * (size_t) n_cells);
  for (int64_t c = 0; c < n_cells; ++c) {
    int64_t idx = 0, nd[4] = {0, 0, 0, 0};
    double skipped = 0.0;
    bool ok = ef.next_int(idx) && ef.next_int(nd[0]) && ef.next_int(nd[1]) && ef.next_int(nd[2]) && ef.next_int(nd[3]);
    for (int64_t a = 0; ok && a < has_attr; ++a) ok = ef.next_double(skipped);
    if (!ok) {
      sb::set_error("cannot read cell # %lld from '%s.ele'", (long long) c, path_prefix);
      return SB_ERR_INVALID;
    }
    for (int q = 0; q < 4; ++q) {
      const int64_t v = nd[q] - first_index;
      SBM_REQUIRE(v >= 0 && v < n_nodes, "cell node index out of range");
      cells[4 * (size_t) c + q] = (int32_t) v;
    }
  }
  return sb_mesh_from_cells(SB_CELL_TET, n_nodes, xyz.data(), n_cells, cells.data(), out);
}

// 2-D ingestion (Triangle's .node / .edge / .ele, the files the reference's own tests use): restates what
// read_mesh_from_tetgen (Mallard/IoTetgen.hpp:44-235) + UnstructuredMesh::insert (MeshUnstructured.hpp:350-425) +
// _update_face_orientation (:509-554) + assign_labels (:464-500) produce for a 2-D mesh:
//   * faces (edges) are numbered in .edge file order; an edge of a cell that the file does not list is appended
//     behind them with label 0 (find_or_insert);
//   * the first cell inserted next to a face is its inner cell, the second one its outer cell;
//   * assign_labels stable-sorts the faces by label: label 0 (interior) first in file order, then each boundary
//     label in file order;
//   * geometry, operation by operation (Shape.hpp:155-167,243-247,310-322; MatrixAlgorithms.hpp:262-270,303-305):
//     edge length = sqrt((0 + dx*dx) + dy*dy), centres = node sums / node count, triangle area =
//     0.5*|d0x*d1y - d0y*d1x|, centre distance as Playground.cpp:126, mirror-ghost distance 2*|x_f - x_i|.
// The result is a face-list handle whose SoA is bit-identical to the reference mesh classes' export
// (tests/golden/mesh_*.npz, made by oracle/_ref/ref_mesh_tool).
int sb_mesh_read_tetgen_2d(const char* path_prefix, sb_mesh** out) {
  SBM_REQUIRE(path_prefix != nullptr && out != nullptr, "null argument");
  *out = nullptr;
  const std::string prefix{path_prefix};
  TokenFile nf, gf, ef;
  if (!nf.open(prefix + ".node")) {
    sb::set_error("cannot open the node file '%s.node'", path_prefix);
    return SB_ERR_INVALID;
  }
  if (!gf.open(prefix + ".edge")) {
    sb::set_error("cannot open the edge file '%s.edge'", path_prefix);
    return SB_ERR_INVALID;
  }
  if (!ef.open(prefix + ".ele")) {
    sb::set_error("cannot open the cell file '%s.ele'", path_prefix);
    return SB_ERR_INVALID;
  }
  // nodes
  int64_t n_nodes = 0, dim = 0, n_attr = 0, has_labels = 0;
  SBM_REQUIRE(nf.next_int(n_nodes) && nf.next_int(dim) && nf.next_int(n_attr) && nf.next_int(has_labels),
              "cannot read the node file header");
  SBM_REQUIRE(dim == 2, "unexpected number of dimensions in the node file header (expected 2)");
  SBM_REQUIRE(n_nodes > 0 && n_nodes < (int64_t) INT32_MAX && n_attr >= 0, "bad node file header");
  std::vector<double> xy(2 * (size_t) n_nodes);
  int64_t first_index = 0;
  for (int64_t k = 0; k < n_nodes; ++k) {
    int64_t idx = 0;
    double skipped = 0.0;
    bool ok = nf.next_int(idx) && nf.next_double(xy[2 * (size_t) k]) && nf.next_double(xy[2 * (size_t) k + 1]);
    for (int64_t a = 0; ok && a < n_attr + (has_labels ? 1 : 0); ++a) ok = nf.next_double(skipped);
    if (!ok) {
      sb::set_error("cannot read node # %lld from '%s.node'", (long long) k, path_prefix);
      return SB_ERR_INVALID;
    }
    if (k == 0) first_index = idx;
    SBM_REQUIRE(idx == first_index + k, "node indices must be consecutive");
  }
  SBM_REQUIRE(first_index == 0 || first_index == 1, "node numbering must start at 0 or 1");
  // edges (= faces), in file order, with labels
  int64_t n_edges = 0, edges_have_labels = 0;
  SBM_REQUIRE(gf.next_int(n_edges) && gf.next_int(edges_have_labels), "cannot read the edge file header");
  SBM_REQUIRE(n_edges >= 0 && n_edges < (int64_t) INT32_MAX, "bad edge count");
  struct Face {
    int32_t n1, n2;      // node order as stored (reversed when the first cell sees the edge the other way round)
    int32_t cell[2];
    int32_t label;
    int64_t file_pos;    // index before the label sort
  };
  std::vector<Face> faces;
  faces.reserve((size_t) n_edges);
  std::unordered_map<uint64_t, int32_t> by_nodes;
  by_nodes.reserve((size_t) n_edges * 2);
  auto key = [](int32_t a, int32_t b) {
    const uint32_t lo = (uint32_t) std::min(a, b), hi = (uint32_t) std::max(a, b);
    return ((uint64_t) hi << 32) | lo;
  };
  for (int64_t e = 0; e < n_edges; ++e) {
    int64_t idx = 0, a = 0, b = 0, label = 0;
    bool ok = gf.next_int(idx) && gf.next_int(a) && gf.next_int(b);
    if (ok && edges_have_labels) ok = gf.next_int(label);
    if (!ok) {
      sb::set_error("cannot read edge # %lld from '%s.edge'", (long long) e, path_prefix);
      return SB_ERR_INVALID;
    }
    a -= first_index, b -= first_index;
    SBM_REQUIRE(a >= 0 && a < n_nodes && b >= 0 && b < n_nodes && a != b, "edge node index out of range");
    SBM_REQUIRE(label >= 0 && label < (1 << 20), "edge label out of range");
    const bool fresh = by_nodes.emplace(key((int32_t) a, (int32_t) b), (int32_t) faces.size()).second;
    SBM_REQUIRE(fresh, "the edge file lists an edge twice");
    faces.push_back(Face{(int32_t) a, (int32_t) b, {-1, -1}, (int32_t) label, e});
  }
  // cells, in file order: connect to the faces, fix inner / outer
  int64_t n_cells = 0, npc = 0, has_attr = 0;
  SBM_REQUIRE(ef.next_int(n_cells) && ef.next_int(npc) && ef.next_int(has_attr), "cannot read the cell file header");
  SBM_REQUIRE(npc == 3, "unexpected number of nodes per cell in the cell file header (expected 3)");
  SBM_REQUIRE(n_cells > 0 && n_cells < (int64_t) 250'000'000, "bad cell count");
  std::vector<int32_t> tri(3 * (size_t) n_cells);
  for (int64_t c = 0; c < n_cells; ++c) {
    int64_t idx = 0, nd[3] = {0, 0, 0};
    double skipped = 0.0;
    bool ok = ef.next_int(idx) && ef.next_int(nd[0]) && ef.next_int(nd[1]) && ef.next_int(nd[2]);
    for (int64_t a = 0; ok && a < has_attr; ++a) ok = ef.next_double(skipped);
    if (!ok) {
      sb::set_error("cannot read cell # %lld from '%s.ele'", (long long) c, path_prefix);
      return SB_ERR_INVALID;
    }
    for (int q = 0; q < 3; ++q) {
      const int64_t v = nd[q] - first_index;
      SBM_REQUIRE(v >= 0 && v < n_nodes, "cell node index out of range");
      tri[3 * (size_t) c + q] = (int32_t) v;
    }
    // Triangle::edges (Shape.hpp:303-305): (n1,n2), (n2,n3), (n3,n1)
    for (int q = 0; q < 3; ++q) {
      const int32_t a = tri[3 * (size_t) c + q], b = tri[3 * (size_t) c + (q + 1) % 3];
      SBM_REQUIRE(a != b, "degenerate triangle");
      auto it = by_nodes.find(key(a, b));
      if (it == by_nodes.end()) { // find_or_insert: an edge the file does not list, label 0
        it = by_nodes.emplace(key(a, b), (int32_t) faces.size()).first;
        faces.push_back(Face{a, b, {-1, -1}, 0, (int64_t) faces.size()});
      }
      Face& f = faces[(size_t) it->second];
      if (f.cell[0] < 0) {
        f.cell[0] = (int32_t) c;
        if (!(f.n1 == a && f.n2 == b)) std::swap(f.n1, f.n2); // the first cell becomes the inner one: flip the face
      } else {
        SBM_REQUIRE(f.cell[1] < 0, "an edge has more than two adjacent cells");
        SBM_REQUIRE(f.n1 == b && f.n2 == a, "inconsistent cell orientation: the second cell of an edge cannot be the outer one");
        f.cell[1] = (int32_t) c;
      }
    }
  }
  // assign_labels: stable sort by label
  std::vector<int32_t> order(faces.size());
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return faces[(size_t) x].label < faces[(size_t) y].label; });
  // geometry
  auto px = [&](int32_t i) { return xy[2 * (size_t) i]; };
  auto py = [&](int32_t i) { return xy[2 * (size_t) i + 1]; };
  auto len2d = [](double dx, double dy) { return std::sqrt((0.0 + dx * dx) + dy * dy); };
  std::vector<double> vol((size_t) n_cells), ctr(3 * (size_t) n_cells, 0.0);
  for (int64_t c = 0; c < n_cells; ++c) {
    const int32_t a = tri[3 * (size_t) c], b = tri[3 * (size_t) c + 1], d = tri[3 * (size_t) c + 2];
    const double d0x = px(b) - px(a), d0y = py(b) - py(a), d1x = px(d) - px(a), d1y = py(d) - py(a);
    vol[(size_t) c] = 0.5 * std::fabs(d0x * d1y - d0y * d1x);
    ctr[3 * (size_t) c] = ((px(a) + px(b)) + px(d)) / 3.0;
    ctr[3 * (size_t) c + 1] = ((py(a) + py(b)) + py(d)) / 3.0;
  }
  std::vector<int32_t> face_cell, bface_cell, labels;
  std::vector<double> face_area, face_dist, bface_area, bface_dist, fnorm, bnorm;
  for (const int32_t q : order) {
    const Face& f = faces[(size_t) q];
    SBM_REQUIRE(f.cell[0] >= 0, "an edge of the edge file belongs to no cell");
    const double dx = px(f.n2) - px(f.n1), dy = py(f.n2) - py(f.n1);
    const double area = len2d(dx, dy);
    // Seg normal (Shape.hpp:250-257): -(-d.y, d.x) with d = (v2 - v1)/|v2 - v1|, for the stored (possibly flipped) node order
    const double nx = dy / area, ny = -(dx / area);
    const int32_t ci = f.cell[0];
    if (f.label == 0) {
      SBM_REQUIRE(f.cell[1] >= 0, "an edge with label 0 (interior) has a single adjacent cell");
      const int32_t co = f.cell[1];
      face_cell.push_back(ci), face_cell.push_back(co);
      face_area.push_back(area);
      face_dist.push_back(len2d(ctr[3 * (size_t) co] - ctr[3 * (size_t) ci], ctr[3 * (size_t) co + 1] - ctr[3 * (size_t) ci + 1]));
      fnorm.push_back(nx), fnorm.push_back(ny), fnorm.push_back(0.0);
    } else {
      SBM_REQUIRE(f.cell[1] < 0, "an edge with a boundary label has two adjacent cells");
      const double cx = (px(f.n1) + px(f.n2)) / 2.0, cy = (py(f.n1) + py(f.n2)) / 2.0;
      bface_cell.push_back(ci);
      bface_area.push_back(area);
      bface_dist.push_back(2.0 * len2d(cx - ctr[3 * (size_t) ci], cy - ctr[3 * (size_t) ci + 1]));
      bnorm.push_back(nx), bnorm.push_back(ny), bnorm.push_back(0.0);
      labels.push_back(f.label);
    }
  }
  sb_mesh_soa soa{};
  soa.n_cells = n_cells, soa.n_faces = (int64_t) face_area.size(), soa.n_bfaces = (int64_t) bface_area.size();
  soa.face_cell = face_cell.data(), soa.face_area = face_area.data(), soa.face_dist = face_dist.data();
  soa.cell_vol = vol.data();
  soa.bface_cell = bface_cell.data(), soa.bface_area = bface_area.data(), soa.bface_dist = bface_dist.data();
  const int rc = sb_mesh_from_faces(&soa, ctr.data(), fnorm.data(), bnorm.data(), out);
  if (rc != SB_OK) return rc;
  (*out)->pair_label.assign((size_t) (soa.n_faces + soa.n_bfaces), 0);
  std::copy(labels.begin(), labels.end(), (*out)->pair_label.begin() + soa.n_faces);
  // keep the nodes and the cells' node lists (file order: what mesh.nodes() / cell.for_each_node visit) for
  // sb_mesh_write_vtk; cell renumbering permutes `cells` like the node lists of the 3-D meshes
  (*out)->npc = 3, (*out)->n_nodes = n_nodes;
  (*out)->xyz.assign(3 * (size_t) n_nodes, 0.0);
  for (int64_t k = 0; k < n_nodes; ++k) (*out)->xyz[3 * (size_t) k] = xy[2 * (size_t) k], (*out)->xyz[3 * (size_t) k + 1] = xy[2 * (size_t) k + 1];
  (*out)->cells = tri;
  return SB_OK;
}

int sb_mesh_bface_labels(const sb_mesh* m, int32_t* h_labels) {
  SBM_REQUIRE(m != nullptr && h_labels != nullptr, "null argument");
  const int64_t n_int = (int64_t) m->face_area.size(), n_b = (int64_t) m->bface_area.size();
  for (int64_t b = 0; b < n_b; ++b)
    h_labels[b] = m->pair_label.empty() ? 1 : m->pair_label[(size_t) m->face_order[(size_t) (n_int + b)]];
  return SB_OK;
}

int sb_mesh_destroy(sb_mesh* mesh) {
  delete mesh;
  return SB_OK;
}

int sb_mesh_renumber_rcm(sb_mesh* mesh, int32_t* h_perm) {
  SBM_REQUIRE(mesh != nullptr, "mesh is null");
  std::vector<int32_t> perm;
  rcm_order(*mesh, perm);
  if (h_perm != nullptr) std::memcpy(h_perm, perm.data(), sizeof(int32_t) * perm.size());
  return apply_permutation(*mesh, perm.data());
}

int sb_mesh_permute_cells(sb_mesh* mesh, const int32_t* h_perm) {
  SBM_REQUIRE(mesh != nullptr && h_perm != nullptr, "null argument");
  return apply_permutation(*mesh, h_perm);
}

int sb_mesh_get_soa(const sb_mesh* m, sb_mesh_soa* soa) {
  SBM_REQUIRE(m != nullptr && soa != nullptr, "null argument");
  soa->n_cells = m->n_cells;
  soa->n_faces = (int64_t) m->face_area.size();
  soa->face_cell = m->face_cell.data(), soa->face_area = m->face_area.data(), soa->face_dist = m->face_dist.data();
  soa->cell_vol = m->cell_vol.data();
  soa->n_bfaces = (int64_t) m->bface_area.size();
  soa->bface_cell = m->bface_cell.data(), soa->bface_area = m->bface_area.data(), soa->bface_dist = m->bface_dist.data();
  return SB_OK;
}

int sb_mesh_cell_centers(const sb_mesh* m, double* h_xyz) {
  SBM_REQUIRE(m != nullptr && h_xyz != nullptr, "null argument");
  SBM_REQUIRE(!m->cell_ctr.empty(), "this face-list mesh was created without cell centres");
  std::memcpy(h_xyz, m->cell_ctr.data(), sizeof(double) * m->cell_ctr.size());
  return SB_OK;
}

int sb_mesh_face_normals(const sb_mesh* m, double* h_fn, double* h_bn) {
  SBM_REQUIRE(m != nullptr, "mesh is null");
  SBM_REQUIRE(m->kind != SB_CELL_FACELIST || !m->pair_normal.empty(), "this face-list mesh was created without normals");
  face_normals(*m, h_fn, h_bn);
  return SB_OK;
}

// Legacy-VTK dump in the file grammar of the playground's save_vtk (Playground.cpp:65-109): same header lines, 16
// significant digits (digits10 + 1), "POINTS n double", count-prefixed node lists under CELLS, one type per cell under
// CELL_TYPES, one "SCALARS <name> double 1 / LOOKUP_TABLE default" block per field under CELL_DATA. The reference
// writes 2-D triangles (z = 0, type 5): a mesh read by sb_mesh_read_tetgen_2d produces that file byte for byte
// (tests/golden/vtk_square_nb_head.txt); for the 3-D meshes the types are VTK_TETRA (10) / VTK_HEXAHEDRON (12), whose
// node orders are those of Shape.hpp:559-606 / 803-818.
int sb_mesh_write_vtk(const sb_mesh* m, const char* path, int n_fields, const char* const* names,
                      const double* const* h_fields) {
  SBM_REQUIRE(m != nullptr && path != nullptr && n_fields >= 0, "null argument");
  SBM_REQUIRE(m->kind == SB_CELL_TET || m->kind == SB_CELL_HEX || (m->npc == 3 && !m->cells.empty()),
              "a face-list mesh has no nodes to write");
  SBM_REQUIRE(n_fields == 0 || (names != nullptr && h_fields != nullptr), "null field arrays");
  for (int f = 0; f < n_fields; ++f) SBM_REQUIRE(names[f] != nullptr && h_fields[f] != nullptr, "null field");
  std::ofstream file(path);
  if (!file) {
    sb::set_error("cannot open '%s' for writing", path);
    return SB_ERR_INVALID;
  }
  file << std::setprecision(std::numeric_limits<double>::digits10 + 1);
  file << "# vtk DataFile Version 2.0" << '\n';
  file << "# Generated by Feathers/StormRuler/Mesh2VTK" << '\n';
  file << "ASCII" << '\n';
  file << "DATASET UNSTRUCTURED_GRID" << '\n';
  file << "POINTS " << m->n_nodes << " double" << '\n';
  for (int64_t i = 0; i < m->n_nodes; ++i)
    file << m->xyz[3 * (size_t) i] << " " << m->xyz[3 * (size_t) i + 1] << " " << m->xyz[3 * (size_t) i + 2] << '\n';
  file << '\n';
  file << "CELLS " << m->n_cells << " " << m->n_cells * (m->npc + 1) << '\n';
  for (int64_t c = 0; c < m->n_cells; ++c) {
    file << m->npc << " ";
    for (int k = 0; k < m->npc; ++k) file << m->cells[(size_t) c * m->npc + k] << " ";
    file << '\n';
  }
  file << '\n';
  file << "CELL_TYPES " << m->n_cells << '\n';
  // VTK_TRIANGLE for the 2-D meshes of sb_mesh_read_tetgen_2d: the playground's own output (Playground.cpp:97-99)
  const char* type = m->kind == SB_CELL_TET ? "10" : (m->kind == SB_CELL_HEX ? "12" : "5");
  for (int64_t c = 0; c < m->n_cells; ++c) file << type << '\n';
  file << '\n';
  file << "CELL_DATA " << m->n_cells << '\n';
  for (int f = 0; f < n_fields; ++f) {
    file << "SCALARS " << names[f] << " double 1" << '\n';
    file << "LOOKUP_TABLE default" << '\n';
    for (int64_t c = 0; c < m->n_cells; ++c) file << h_fields[f][c] << '\n';
  }
  file << '\n';
  file.close();
  if (!file) {
    sb::set_error("write to '%s' failed", path);
    return SB_ERR_INVALID;
  }
  return SB_OK;
}

int64_t sb_mesh_bandwidth(const sb_mesh* m) {
  if (m == nullptr) return -1;
  int64_t bw = 0;
  for (size_t f = 0; f < m->face_area.size(); ++f)
    bw = std::max<int64_t>(bw, std::llabs((long long) m->face_cell[2 * f] - (long long) m->face_cell[2 * f + 1]));
  return bw;
}

} // extern "C"
