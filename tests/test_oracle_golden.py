"""CPU tests: the oracle restatement against the golden vectors produced by the reference itself
(tests/golden/make_golden.py), and -- when oracle/_ref is built -- against the reference live."""
import numpy as np
import pytest

from conftest import load_golden, rhs
from oracle import orc

DT, ITERS, RTOL = 0.05, 500, 1e-10


def helmholtz(mesh):
    return orc.FaceOp(mesh, prefill=1, dt=-DT)


def poisson(mesh):
    return orc.FaceOp(mesh, prefill=0, dt=-1.0, dirichlet=True)


# ---- BLAS-1 known answers of the reference's unit tests -----------------------------------------
def test_blas1_known_answers():
    g = load_golden("blas1_kat.npz")
    m1, m2, m3 = g["mat1"], g["mat2"], g["mat3"]
    assert orc.dot(m1, m2) == 70.0 == float(g["dot_mat1_mat2"])          # BitternReductions.cpp:109
    assert abs(orc.norm2(m1) - 5.47723) < 1e-5 * 5.47723                 # BitternReductions.cpp:72
    assert orc.norm2(m1) == float(g["norm2_mat1"])
    assert np.array_equal(m1 + 10.0 * (m2 - m3), g["expr1"])             # BitternMath.cpp:143-148
    assert orc.dot(g["big_a"], g["big_b"]) == float(g["dot_big"])        # sequential order, bit-exact
    assert orc.norm2(g["big_a"]) == float(g["norm2_big"])


def test_safe_divide():
    L = orc.lib()
    assert L.orc_safe_divide(1.0, 0.0) == 0.0 and L.orc_safe_divide(1.0, -0.0) == 0.0
    assert L.orc_safe_divide(3.0, 2.0) == 1.5


def test_tree_reduction_close_to_sequential():
    rng = np.random.default_rng(3)
    for n in (0, 1, 2, 63, 64, 65, 2047, 2048, 2049, 10_000, 600_001):
        a, b = rng.standard_normal(n), rng.standard_normal(n)
        s, t = orc.dot(a, b), orc.dot(a, b, orc.RED_TREE)
        assert abs(s - t) <= 1e-12 * max(1.0, np.abs(a * b).sum())
    assert orc.dot(np.zeros(0), np.zeros(0), orc.RED_TREE) == 0.0


# ---- operator -----------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["square_nb", "rectangle"])
def test_mesh_fixture_shape(name, request):
    mesh = request.getfixturevalue(name)
    g = load_golden(f"mesh_{name}.npz")
    # 6 252 / 12 776 triangles, 9 256 / 18 828 interior faces, 3 face labels
    expect = {"square_nb": (6252, 9256), "rectangle": (12776, 18828)}[name]
    assert (mesh.n_cells, mesh.n_faces) == expect and int(g["n_face_labels"]) == 3
    assert mesh.n_faces + mesh.n_bfaces == int(g["n_faces_total"])
    assert (mesh.face_cell >= 0).all() and (mesh.face_cell < mesh.n_cells).all()
    assert (mesh.face_cell[:, 0] != mesh.face_cell[:, 1]).all()
    assert (mesh.face_area > 0).all() and (mesh.face_dist > 0).all() and (mesh.cell_vol > 0).all()


@pytest.mark.parametrize("name", ["square_nb", "rectangle"])
def test_face_loop_cg_matches_reference_native(name, request):
    """Reference mesh classes + CellField + CgSolver (golden) == oracle face loop + restated CG."""
    mesh = request.getfixturevalue(name)
    g = load_golden(f"cg_native_{name}.npz")
    r = orc.solve("cg", helmholtz(mesh), rhs(mesh.n_cells), num_iterations=ITERS, abs_tol=0.0, rel_tol=RTOL)
    assert r.iterations == int(g["iterations"]) and r.converged == bool(g["converged"])
    assert r.abs_err == float(g["abs_err"]) and r.rel_err == float(g["rel_err"])
    assert np.array_equal(r.hist, g["hist"])
    assert np.array_equal(r.x, g["x"])


def test_face_loop_cg_matches_reference_native_on_step(step):
    """The largest reference mesh (79 672 cells): the reference's own run does not converge within 500 iterations; the
    restated face loop + CG reproduces its residual history and its iterate bit for bit (the fixture keeps every 8th
    entry of x and its norm)."""
    g = load_golden("cg_native_step.npz")
    r = orc.solve("cg", helmholtz(step), rhs(step.n_cells), num_iterations=ITERS, abs_tol=0.0, rel_tol=RTOL)
    assert (r.converged, r.iterations) == (False, 500) == (bool(g["converged"]), int(g["iterations"]))
    assert r.abs_err == float(g["abs_err"]) == 0.23301109656816443
    assert np.array_equal(r.hist, g["hist"])
    assert np.array_equal(r.x[::8], g["x_every_8th"]) and np.linalg.norm(r.x) == float(g["x_norm"])


def test_survey_probe_numbers(square_nb, rectangle):
    """The known-answer results recorded in BASELINE.md / SURVEY.md Appendix A-3."""
    r = orc.solve("cg", helmholtz(square_nb), rhs(square_nb.n_cells), num_iterations=ITERS, abs_tol=0.0, rel_tol=RTOL)
    assert (r.converged, r.iterations, r.abs_err) == (True, 366, 5.3201229897199972e-09)
    r = orc.solve("cg", helmholtz(rectangle), rhs(rectangle.n_cells), num_iterations=ITERS, abs_tol=0.0, rel_tol=RTOL)
    assert (r.converged, r.iterations, r.abs_err) == (False, 500, 1.6613183551327937e-07)


@pytest.mark.parametrize("make", [helmholtz, poisson])
def test_row_forms(square_nb, make):
    op = make(square_nb)
    rng = np.random.default_rng(0)
    for _ in range(3):
        x = rng.standard_normal(op.n)
        y = op.apply(x)
        assert np.array_equal(op.apply_rows_faithful(x), y)   # bit-identical to the face loop
        yc = op.apply_rows_coef(x)
        assert np.abs(yc - y).max() <= 1e-13 * np.abs(y).max()  # same algebra, different rounding
    w, ld, col, face = op.rows(ld=2048 * 4)
    assert ld == 8192 and col.shape == (w, ld)
    deg = (col != orc.COL_PAD).sum(0)
    assert (deg[op.n:] == 0).all() and deg[:op.n].max() == w
    # entries of a row are in ascending face index
    f = np.where(face >= 0, face, np.iinfo(np.int64).max)
    assert (np.diff(f, axis=0) >= 0).all()


def test_linearity_and_constant_nullspace(square_nb):
    lap = orc.FaceOp(square_nb, prefill=0, dt=1.0)  # pure Neumann Laplacian
    assert np.abs(lap.apply(np.ones(lap.n))).max() == 0.0
    rng = np.random.default_rng(1)
    x, y = rng.standard_normal(lap.n), rng.standard_normal(lap.n)
    lhs, rhs_ = lap.apply(x + 2.0 * y), lap.apply(x) + 2.0 * lap.apply(y)
    assert np.abs(lhs - rhs_).max() <= 1e-12 * np.abs(lhs).max()


# ---- solvers ------------------------------------------------------------------------------------
@pytest.mark.parametrize("solver", ["cg", "bicgstab"])
@pytest.mark.parametrize("prefix,make", [("", helmholtz), ("poisson_", poisson)])
def test_restated_solvers_match_reference_headers(square_nb, solver, prefix, make):
    g = load_golden("solvers_square_nb.npz")
    r = orc.solve(solver, make(square_nb), rhs(square_nb.n_cells), num_iterations=ITERS, abs_tol=0.0, rel_tol=RTOL)
    conv, it, abs_err, rel_err, n_apply, n_trace = g[f"{prefix}{solver}_stats"]
    assert (r.converged, r.iterations, r.abs_err, r.rel_err) == (bool(conv), int(it), abs_err, rel_err)
    assert np.array_equal(r.hist, g[f"{prefix}{solver}_hist"])
    assert np.array_equal(r.x, g[f"{prefix}{solver}_x"])
    assert len(r.trace) == int(n_trace)
    if not prefix:
        head = g[f"{solver}_trace_head"]
        assert np.array_equal(r.trace[:len(head)], head)


def test_early_exit_and_zero_iterations(square_nb):
    op = helmholtz(square_nb)
    b = rhs(op.n)
    r = orc.solve("cg", op, b, num_iterations=0, abs_tol=0.0, rel_tol=RTOL)
    assert r.iterations == 0 and not r.converged and len(r.hist) == 1 and np.array_equal(r.x, np.zeros(op.n))
    r = orc.solve("bicgstab", op, b, num_iterations=10, abs_tol=1e9, rel_tol=RTOL)  # Solver.hpp:124-128
    assert r.iterations == 0 and r.converged and len(r.hist) == 1


def test_reduction_order_sensitivity_is_documented(square_nb):
    """CG is insensitive to the reduction order; BiCGStab's erratic convergence amplifies it
    (DESIGN.md 'parity'): sequential-sum and tree-sum histories agree to 1e-10 only early on."""
    op, b = helmholtz(square_nb), rhs(square_nb.n_cells)
    s = orc.solve("cg", op, b, num_iterations=ITERS, abs_tol=0.0, rel_tol=RTOL)
    t = orc.solve("cg", op, b, num_iterations=ITERS, abs_tol=0.0, rel_tol=RTOL, mode=orc.RED_TREE)
    assert s.iterations == t.iterations
    assert np.max(np.abs(s.hist - t.hist) / s.hist) < 1e-10
    assert np.linalg.norm(s.x - t.x) <= 1e-8 * np.linalg.norm(s.x)
    s = orc.solve("bicgstab", op, b, num_iterations=ITERS, abs_tol=0.0, rel_tol=RTOL)
    t = orc.solve("bicgstab", op, b, num_iterations=ITERS, abs_tol=0.0, rel_tol=RTOL, mode=orc.RED_TREE)
    k = min(len(s.hist), len(t.hist))
    rel = np.abs(s.hist[:k] - t.hist[:k]) / s.hist[:k]
    assert rel[:20].max() < 1e-10
    assert np.linalg.norm(s.x - t.x) <= 1e-8 * np.linalg.norm(s.x)


# ---- live reference (only where oracle/_ref exists: the build container) --------------------------
needs_ref = pytest.mark.skipif(not orc.have_ref(), reason="oracle/_ref not built (needs /root/reference)")


@needs_ref
@pytest.mark.parametrize("solver", orc.REF_SOLVERS)
def test_reference_headers_reproduce_golden(square_nb, solver):
    g = load_golden("solvers_square_nb.npz")
    r = orc.ref_solve(solver, helmholtz(square_nb), rhs(square_nb.n_cells), num_iterations=ITERS, abs_tol=0.0,
                      rel_tol=RTOL)
    assert np.array_equal(r.hist, g[f"{solver}_hist"])
    conv, it, abs_err, rel_err, n_apply, n_trace = g[f"{solver}_stats"]
    assert (r.converged, r.iterations, r.abs_err, r.n_apply, len(r.trace)) == (bool(conv), int(it), abs_err,
                                                                             int(n_apply), int(n_trace))


@needs_ref
def test_fill_randomly_stream_matches_reference_template():
    """Our resettable engine == the reference's fill_randomly template (MatrixAlgorithms.hpp:140-153)."""
    g = load_golden("blas1_kat.npz")
    import numpy.random  # noqa: F401
    # libstdc++ mt19937_64 default seed 5489, generate_canonical<double,53>: x / 2^64
    head = g["fill_randomly_head"]
    assert ((head >= 0) & (head < 1)).all()
    assert abs(head[0] - 14514284786278117030 / 2.0**64) < 1e-16  # first mt19937_64 output, default seed


# ---- convection-diffusion: face loop in the pattern of ConvectionScheme.hpp:83-106 ----
def convdiff_case(kind="tet", dims=(6, 5, 4), nu=0.02, beta=(1.0, 0.5, 0.25)):
    from stormruler_b200.mesh import CELL_HEX, CELL_TET, Mesh
    mesh = Mesh.box(CELL_TET if kind == "tet" else CELL_HEX, *dims)
    mesh.renumber_rcm()
    fu, bu = mesh.face_flux(beta)
    fm = orc.FaceMesh(mesh.n_cells, mesh.face_cell, mesh.face_area, mesh.face_dist, mesh.cell_vol, mesh.bface_cell,
                      mesh.bface_area, mesh.bface_dist)
    return mesh, orc.ConvDiffOp(fm, nu, fu, bu)


@pytest.mark.parametrize("kind", ["tet", "hex"])
def test_convdiff_face_loop_properties_and_row_form(kind):
    mesh, op = convdiff_case(kind)
    rng = np.random.default_rng(3)
    x, z = rng.standard_normal(op.n), rng.standard_normal(op.n)
    # linear; with nu = 0 and beta = 0 it vanishes; with beta = 0 it is nu times the Poisson operator
    assert np.allclose(op.apply(2.0 * x + z), 2.0 * op.apply(x) + op.apply(z), rtol=1e-12, atol=1e-9)
    zero = orc.ConvDiffOp(op.mesh, 0.0, np.zeros(op.mesh.n_faces), np.zeros(op.mesh.n_bfaces))
    assert np.array_equal(zero.apply(x), np.zeros(op.n))
    diff = orc.ConvDiffOp(op.mesh, 0.7, np.zeros(op.mesh.n_faces), np.zeros(op.mesh.n_bfaces))
    lap = orc.FaceOp(op.mesh, prefill=0, dt=-0.7, dirichlet=True)
    assert np.allclose(diff.apply(x), lap.apply(x), rtol=1e-12, atol=1e-9)
    # the operator is genuinely non-symmetric (two different coefficients per face)
    assert abs(np.dot(z, op.apply(x)) - np.dot(x, op.apply(z))) > 1e-3
    # row form = face loop within rounding; upwinding makes it an M-matrix: positive diagonal, a <= 0
    rows = op.rows_coef()
    y_rows, y_faces = rows.apply(x), op.apply(x)
    assert np.abs(y_rows - y_faces).max() <= 1e-13 * np.abs(y_faces).max()
    w, ld, col, a, diag = rows.rows
    assert (diag[:op.n] > 0).all() and (a[col != orc.COL_PAD] <= 0).all()


def test_convdiff_reference_gmres_and_bicgstab_converge():
    """The reference's own GMRES / FGMRES / BiCGStab headers on the restated operator."""
    _, op = convdiff_case("tet", (7, 6, 5))
    b = np.sin(0.37 * np.arange(op.n))
    for s in ("gmres", "fgmres", "bicgstab", "idrs"):
        r = orc.ref_solve(s, op, b, num_iterations=300, abs_tol=0.0, rel_tol=1e-10)
        res = np.linalg.norm(b - op.apply(r.x)) / np.linalg.norm(b)
        assert r.converged and res < 1e-8, (s, r.iterations, res)


@pytest.mark.parametrize("seed", range(6))
def test_restated_solvers_match_the_compiled_reference_on_random_3d_problems(seed):
    """Pin of the C restatement beyond the golden meshes: random jittered tetrahedral / hexahedral / polyhedral
    problems, Helmholtz and Dirichlet-Poisson, sequential and tree reductions -- iteration count, residual history,
    reduction trace and solution bit for bit against the reference's own headers (oracle/_ref)."""
    if not orc.have_ref():
        pytest.skip("oracle/_ref not built (needs the StormRuler sources at build time)")
    from stormruler_b200.mesh import CELL_HEX, CELL_TET, Mesh, PolyMesh
    rng = np.random.default_rng(seed)
    if seed % 3 == 0:
        m = Mesh.box(CELL_TET, *[int(v) for v in rng.integers(2, 6, 3)], jitter=0.2, seed_jitter=seed)
    elif seed % 3 == 1:
        m = Mesh.box(CELL_HEX, *[int(v) for v in rng.integers(2, 7, 3)], jitter=0.2, seed_jitter=seed)
    else:
        m = PolyMesh.bcc(int(rng.integers(2, 5)), (1.0, float(rng.uniform(.7, 1.4)), float(rng.uniform(.7, 1.4))))
    fm = orc.FaceMesh(m.n_cells, m.face_cell, m.face_area, m.face_dist, m.cell_vol, m.bface_cell, m.bface_area, m.bface_dist)
    op = orc.FaceOp(fm, prefill=1, dt=-0.05) if seed % 2 else orc.FaceOp(fm, prefill=0, dt=-1.0, dirichlet=True)
    b = rng.standard_normal(m.n_cells)
    for solver in ("cg", "bicgstab"):
        for mode in (orc.RED_SEQ, orc.RED_TREE):
            a = orc.solve(solver, op, b, num_iterations=60, abs_tol=0.0, rel_tol=1e-11, mode=mode)
            r = orc.ref_solve(solver, op, b, num_iterations=60, abs_tol=0.0, rel_tol=1e-11, mode=mode)
            assert (a.converged, a.iterations) == (r.converged, r.iterations)
            assert np.array_equal(a.hist, r.hist) and np.array_equal(a.x, r.x)
            k = min(len(a.trace), len(r.trace))
            assert k > 0 and np.array_equal(a.trace[:k], r.trace[:k])
