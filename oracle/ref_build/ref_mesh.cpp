// TEST INFRASTRUCTURE (oracle/_ref build) -- not product code.
//
// Drives the reference's OWN mesh reader, mesh, views, CellField, Operator and CgSolver,
// unmodified, from /root/reference/source (nothing is copied into this repo):
//   Storm/Mallard/IoTetgen.hpp:44-235   read_mesh_from_tetgen
//   Storm/Mallard/MeshUnstructured.hpp  UnstructuredMesh<2,2,VovTable> -> <2,2,CsrTable> (:291-312)
//   Storm/Mallard/Mesh.hpp:241-323      FaceView / CellView getters, :453-455 interior_faces
//   Storm/Feathers/Field.hpp:60-114     CellField (the solver's Vector)
//   Storm/Solvers/SolverCg.hpp          CgSolver, through Solver.hpp:116-147
// mirroring source_apps/playground/Playground.cpp:248-255 (load) and :153-167 (operator lambda).
//
// Commands:
//   ref_mesh_tool export <prefix> <out.bin>
//       SoA dump of the hot-path connectivity and geometry, in the reference's entity order:
//       interior faces (label 0) with inner/outer cell, area, centre distance as evaluated by
//       Playground.cpp:125-126; boundary faces with inner cell, area, label and the mirror-ghost
//       distance 2*|face centre - cell centre|; cell volumes and centres.
//   ref_mesh_tool cg <prefix> <dt> <num_iterations> <rel_tol> <out.bin>
//       Runs the reference CgSolver on CellField with y = x - dt * div grad x written as the
//       playground writes it, b[k] = sin(0.37 k), x0 = 0 (SURVEY.md 8d "Config 1");
//       dumps b, x, the residual history and the solver's final public fields.
//
// The only code here that is not the reference's is the restated face loop `div_grad` below
// (the original is a file-local function of the playground app and cannot be included).

#include <Storm/Solvers/Operator.hpp>
#include <Storm/Solvers/SolverCg.hpp>

#include <Storm/Feathers/Field.hpp>
#include <Storm/Mallard/IoTetgen.hpp>
#include <Storm/Mallard/MeshUnstructured.hpp>
#include <Storm/Mallard/Shape.hpp>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

using namespace Storm;
using namespace Storm::Feathers;

using RefMesh = UnstructuredMesh<2, 2, CsrTable>;
using RefField = CellField<RefMesh, real_t>;

namespace {

// u += dt * div grad c over interior faces; statement order of Playground.cpp:119-130.
void div_grad(const RefMesh& mesh, RefField& u, real_t dt, const RefField& c) {
  std::ranges::for_each(mesh.interior_faces(), [&](FaceView<RefMesh> face) {
    const CellView<RefMesh> cell_inner = face.inner_cell();
    const CellView<RefMesh> cell_outer = face.outer_cell();
    const auto flux = dt * (c[cell_outer] - c[cell_inner]) /
                      length(cell_outer.center() - cell_inner.center());
    u[cell_inner] += (face.area() / cell_inner.volume()) * flux;
    u[cell_outer] -= (face.area() / cell_outer.volume()) * flux;
  });
}

std::shared_ptr<RefMesh> load(const std::string& prefix) {
  UnstructuredMesh<2, 2, VovTable> mesh1{};
  read_mesh_from_tetgen(mesh1, prefix);
  auto mesh = std::make_shared<RefMesh>();
  mesh->assign(std::move(mesh1));
  return mesh;
}

struct Writer {
  FILE* f;
  explicit Writer(const char* path) : f{std::fopen(path, "wb")} {
    if (f == nullptr) {
      std::perror(path);
      std::exit(2);
    }
  }
  ~Writer() { std::fclose(f); }
  void i64(int64_t v) { std::fwrite(&v, sizeof(v), 1, f); }
  void f64(double v) { std::fwrite(&v, sizeof(v), 1, f); }
  template<class T>
  void arr(const std::vector<T>& v) {
    i64((int64_t) v.size());
    if (!v.empty()) std::fwrite(v.data(), sizeof(T), v.size(), f);
  }
};

int cmd_export(const std::string& prefix, const char* out) {
  const auto mesh = load(prefix);
  const size_t n_cells = mesh->num_cells();
  std::vector<int32_t> face_cell, bface_cell, bface_label;
  std::vector<double> face_area, face_dist, bface_area, bface_dist, cell_vol, cell_cx, cell_cy;
  for (auto face : mesh->interior_faces()) {
    const auto ci = face.inner_cell(), co = face.outer_cell();
    face_cell.push_back((int32_t) ci.index_sz());
    face_cell.push_back((int32_t) co.index_sz());
    face_area.push_back(face.area());
    face_dist.push_back(length(co.center() - ci.center()));
  }
  for (auto face : mesh->boundary_faces()) {
    const auto ci = face.inner_cell();
    bface_cell.push_back((int32_t) ci.index_sz());
    bface_area.push_back(face.area());
    bface_dist.push_back(2.0 * length(face.center() - ci.center()));
    bface_label.push_back((int32_t) (size_t) face.label());
  }
  for (auto cell : mesh->cells()) {
    cell_vol.push_back(cell.volume());
    const auto c = cell.center();
    cell_cx.push_back(c(0));
    cell_cy.push_back(c(1));
  }
  Writer w{out};
  w.i64((int64_t) n_cells);
  w.i64((int64_t) mesh->num_nodes());
  w.i64((int64_t) mesh->num_faces());
  w.i64((int64_t) mesh->num_face_labels());
  w.arr(face_cell), w.arr(face_area), w.arr(face_dist);
  w.arr(bface_cell), w.arr(bface_area), w.arr(bface_dist), w.arr(bface_label);
  w.arr(cell_vol), w.arr(cell_cx), w.arr(cell_cy);
  std::printf("export: cells=%zu nodes=%zu faces=%zu interior=%zu boundary=%zu labels=%zu\n", n_cells,
              mesh->num_nodes(), mesh->num_faces(), face_area.size(), bface_area.size(),
              mesh->num_face_labels());
  return 0;
}

struct SamplingOperator final : Operator<RefField> {
  const RefMesh* mesh;
  real_t dt;
  const CgSolver<RefField>* solver;
  std::vector<double>* hist;
  void mul(RefField& y, const RefField& x) const override {
    const size_t it = solver->iteration;
    if (hist->size() <= it) hist->resize(it + 1);
    (*hist)[it] = solver->absolute_error;
    y <<= x;                     // Playground.cpp:162  `c_hat <<= c_in`
    div_grad(*mesh, y, -dt, x);  // Playground.cpp:165  `stormDivGrad(mesh, c_hat, -tau, w_hat)`
  }
};

int cmd_cg(const std::string& prefix, double dt, size_t num_iterations, double rel_tol,
           const char* out) {
  const auto mesh = load(prefix);
  const size_t n = mesh->num_cells();
  RefField x{*mesh}, b{*mesh};
  for (size_t k = 0; k < n; ++k) b(k) = std::sin(0.37 * (double) k), x(k) = 0.0;
  CgSolver<RefField> solver{};
  solver.num_iterations = num_iterations;
  solver.absolute_error_tolerance = 0.0;
  solver.relative_error_tolerance = rel_tol;
  std::vector<double> hist;
  SamplingOperator op;
  op.mesh = mesh.get(), op.dt = dt, op.solver = &solver, op.hist = &hist;
  const bool converged = solver.solve(x, b, op);
  if (hist.size() <= solver.iteration) hist.resize(solver.iteration + 1);
  hist[solver.iteration] = solver.absolute_error;
  std::vector<double> xv(n), bv(n);
  for (size_t k = 0; k < n; ++k) xv[k] = x(k), bv[k] = b(k);
  Writer w{out};
  w.i64((int64_t) n);
  w.i64(converged ? 1 : 0);
  w.i64((int64_t) solver.iteration);
  w.f64(solver.absolute_error);
  w.f64(solver.relative_error);
  w.arr(bv), w.arr(xv), w.arr(hist);
  std::printf("cg: n=%zu converged=%d iterations=%zu abs=%.17g rel=%.17g\n", n, (int) converged,
              solver.iteration, solver.absolute_error, solver.relative_error);
  return 0;
}

} // namespace

int main(int argc, char** argv) {
  if (argc >= 4 && std::strcmp(argv[1], "export") == 0) return cmd_export(argv[2], argv[3]);
  if (argc >= 7 && std::strcmp(argv[1], "cg") == 0)
    return cmd_cg(argv[2], std::atof(argv[3]), (size_t) std::atoll(argv[4]), std::atof(argv[5]),
                  argv[6]);
  std::fprintf(stderr,
               "usage: ref_mesh_tool export <prefix> <out.bin>\n"
               "       ref_mesh_tool cg <prefix> <dt> <num_iterations> <rel_tol> <out.bin>\n");
  return 1;
}
