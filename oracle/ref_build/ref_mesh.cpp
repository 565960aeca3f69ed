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
//   ref_mesh_tool vtk <prefix> <out.vtk>
//       The playground's save_vtk (Playground.cpp:65-109) on the reference's mesh with one cell field
//       c[k] = sin(0.37 k): the file a drop-in has to reproduce (node order, per-cell node lists, number format).
//   ref_mesh_tool ch <prefix> <num_steps> <out.bin> [uniformed]
//       (uniformed: the same step through the reference's solve_non_uniform instead of solve<CgSolver>)
//       The playground's own caller of the path, statement for statement (Playground.cpp:133-175 and the
//       initial condition / swap of :176-210): c[cell] = rand()/RAND_MAX (glibc, default seed), then per
//       time step  f <<= map(dF_dc, c);  c_hat <<= c;  solve<CgSolver>(c_hat, c, make_operator(lambda))
//       with the two chained stormDivGrad calls of :153-167 and the constants of :113. Dumps the initial c
//       and, per step, c_hat, the CG iteration count, final errors and the residual history.
//
// The only code here that is not the reference's is the restated face loop `div_grad` below
// (the original is a file-local function of the playground app and cannot be included).

#include <Storm/Bittern/Matrix.hpp>

namespace Storm {
// a name the legacy solver headers expect but the tree no longer defines (Solver.hpp:281)
template<class M>
constexpr auto& fill_with(M& m, double s) {
  return fill(m, s);
}
} // namespace Storm

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
#include <fstream>
#include <iomanip>
#include <limits>
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

// The file the playground's save_vtk (Playground.cpp:65-109) writes for one cell field, produced from the reference's
// mesh classes: what matters for a drop-in is the order in which those classes enumerate nodes, cells and a cell's
// nodes, and the stream formatting (setprecision(digits10 + 1), default float format); the writer itself is restated.
int cmd_vtk(const std::string& prefix, const char* out) {
  const auto mesh_ptr = load(prefix);
  const RefMesh& mesh = *mesh_ptr;
  const size_t n_cells = mesh.num_cells();
  RefField field{mesh};
  for (size_t k = 0; k < n_cells; ++k) field(k) = std::sin(0.37 * (double) k);
  std::ofstream os(out);
  os << std::setprecision(std::numeric_limits<real_t>::digits10 + 1);
  for (const char* header : {"# vtk DataFile Version 2.0", "# Generated by Feathers/StormRuler/Mesh2VTK", "ASCII",
                             "DATASET UNSTRUCTURED_GRID"}) {
    os << header << '\n';
  }
  os << "POINTS " << mesh.num_nodes() << " double\n";
  for (auto node : mesh.nodes()) {
    const auto& x = node.position();
    os << x(0) << ' ' << x(1) << ' ' << 0.0 << '\n';
  }
  size_t list_size = 0;
  for (auto cell : mesh.interior_cells()) list_size += cell.nodes().size() + 1;
  os << "\nCELLS " << n_cells << ' ' << list_size << '\n';
  for (auto cell : mesh.interior_cells()) {
    os << cell.nodes().size() << ' ';
    cell.for_each_node([&os](NodeIndex node) { os << node << ' '; });
    os << '\n';
  }
  os << "\nCELL_TYPES " << n_cells << '\n';
  for (size_t k = 0; k < n_cells; ++k) os << "5\n"; // VTK_TRIANGLE
  os << "\nCELL_DATA " << n_cells << "\nSCALARS c double 1\nLOOKUP_TABLE default\n";
  for (auto cell : mesh.interior_cells()) os << field[cell] << '\n';
  os << '\n';
  return os ? 0 : 2;
}

// One Cahn-Hilliard time step, Playground.cpp:133-175 (the CgSolver is constructed here instead of inside
// solve<CgSolver>, Solver.hpp:261-265, so that its public progress fields can be sampled; same defaults).
constexpr double ch_tau = 1.0e-3, ch_Gamma = 1.0e-4, ch_sigma = 2.0; // Playground.cpp:113

struct ChStepReport {
  bool converged;
  size_t iterations;
  double abs_err, rel_err;
  std::vector<double> hist;
};

ChStepReport cahn_hilliard_step(const RefMesh& mesh, const RefField& c, RefField& c_hat, RefField& w_hat,
                                bool uniformed = false) {
  constexpr auto dF_dc = [](real_t c) noexcept { return 2.0 * c * (c - 1.0) * (2.0 * c - 1.0); };
  RefField f{mesh};
  f <<= map(dF_dc, c);
  c_hat <<= c;
  CgSolver<RefField> solver{};
  ChStepReport rep{};
  const auto op = make_operator<RefField>([&](RefField& c_out, const RefField& c_in) {
    const size_t it = solver.iteration;
    if (rep.hist.size() <= it) rep.hist.resize(it + 1);
    rep.hist[it] = solver.absolute_error;
    w_hat <<= f + ch_sigma * (c_in - c);
    div_grad(mesh, w_hat, -ch_Gamma, c_in);
    c_out <<= c_in;
    div_grad(mesh, c_out, -ch_tau, w_hat);
  });
  // uniformed: the operator is affine (A(0) = -tau lap(f - sigma c) != 0), which is what the reference's own
  // solve_non_uniform (Solver.hpp:271-292) exists for; the playground calls plain solve<CgSolver> (:149)
  rep.converged = uniformed ? solve_non_uniform(solver, c_hat, c, *op) : solver.solve(c_hat, c, *op);
  rep.iterations = solver.iteration;
  rep.abs_err = solver.absolute_error, rep.rel_err = solver.relative_error;
  if (rep.hist.size() <= solver.iteration) rep.hist.resize(solver.iteration + 1);
  rep.hist[solver.iteration] = solver.absolute_error;
  return rep;
}

int cmd_ch(const std::string& prefix, size_t num_steps, const char* out, bool uniformed) {
  const auto mesh = load(prefix);
  const size_t n = mesh->num_cells();
  RefField c{*mesh}, c_hat{*mesh}, w_hat{*mesh};
  std::ranges::for_each(mesh->interior_cells(), [&](CellView<RefMesh> cell) {
    c[cell] = (1.0 * rand()) / RAND_MAX; // Playground.cpp:182-184
  });
  Writer w{out};
  w.i64((int64_t) n);
  w.i64((int64_t) num_steps);
  std::vector<double> v(n);
  for (size_t k = 0; k < n; ++k) v[k] = c(k);
  w.arr(v);
  for (size_t step = 0; step < num_steps; ++step) {
    const ChStepReport rep = cahn_hilliard_step(*mesh, c, c_hat, w_hat, uniformed);
    std::swap(c, c_hat); // Playground.cpp:203
    for (size_t k = 0; k < n; ++k) v[k] = c(k);
    w.i64(rep.converged ? 1 : 0);
    w.i64((int64_t) rep.iterations);
    w.f64(rep.abs_err);
    w.f64(rep.rel_err);
    w.arr(v), w.arr(rep.hist);
    std::printf("ch: step=%zu converged=%d iterations=%zu abs=%.17g rel=%.17g\n", step, (int) rep.converged,
                rep.iterations, rep.abs_err, rep.rel_err);
  }
  return 0;
}

} // namespace

int main(int argc, char** argv) {
  if (argc >= 4 && std::strcmp(argv[1], "export") == 0) return cmd_export(argv[2], argv[3]);
  if (argc >= 7 && std::strcmp(argv[1], "cg") == 0)
    return cmd_cg(argv[2], std::atof(argv[3]), (size_t) std::atoll(argv[4]), std::atof(argv[5]),
                  argv[6]);
  if (argc >= 4 && std::strcmp(argv[1], "vtk") == 0) return cmd_vtk(argv[2], argv[3]);
  if (argc >= 5 && std::strcmp(argv[1], "ch") == 0)
    return cmd_ch(argv[2], (size_t) std::atoll(argv[3]), argv[4], argc >= 6 && std::strcmp(argv[5], "uniformed") == 0);
  std::fprintf(stderr,
               "usage: ref_mesh_tool export <prefix> <out.bin>\n"
               "       ref_mesh_tool cg <prefix> <dt> <num_iterations> <rel_tol> <out.bin>\n"
               "       ref_mesh_tool vtk <prefix> <out.vtk>\n"
               "       ref_mesh_tool ch <prefix> <num_steps> <out.bin> [uniformed]\n");
  return 1;
}
