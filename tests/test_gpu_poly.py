"""Rows wider than 8 entries: polyhedral cells (dual-polyhedra leg; 14 faces per cell) and
the other wide instantiations of the apply kernels (10, 12, 14, 16; odd degrees are padded to the next even
width). Same bars as the tetrahedral / hexahedral cases: row layout, faithful apply and coefficient apply bit for
bit against the oracle, fused solvers bit for bit against the tree oracle."""
import numpy as np
import pytest

import stormruler_b200 as sb
from oracle import orc
from stormruler_b200.mesh import PolyMesh

pytestmark = pytest.mark.gpu


def face_mesh(m):
    return orc.FaceMesh(m.n_cells, m.face_cell, m.face_area, m.face_dist, m.cell_vol, m.bface_cell, m.bface_area,
                        m.bface_dist)


def band_mesh(n, k, odd, n_ghost, seed):
    """A synthetic face list with prescribed row degrees: cell i shares a face with i+1 .. i+k (interior degree 2k
    away from the ends), `odd` adds one more face to every second cell, `n_ghost` boundary faces go to random cells.
    Geometry is random but positive. Faces are ordered by creating cell, like the reference's insertion order."""
    rng = np.random.default_rng(seed)
    inner, outer = [], []
    for i in range(n):
        for o in range(1, k + 1):
            if i + o < n:
                inner.append(i), outer.append(i + o)
        if odd and i % 4 == 0 and i + 102 < n:
            inner.append(i), outer.append(i + 102)
    fc = np.stack([np.array(inner, np.int32), np.array(outer, np.int32)], axis=1)
    F = fc.shape[0]
    # boundary faces: random cells, plus two on one full-degree cell so the faithful width is exactly degree + 2
    bc = np.zeros(0, np.int32)
    if n_ghost:
        others = np.setdiff1d(np.arange(n), [n // 2])
        bc = np.sort(np.concatenate([rng.choice(others, size=n_ghost, replace=False), [n // 2, n // 2]])).astype(np.int32)
    return orc.FaceMesh(n, fc, rng.uniform(0.5, 1.5, F), rng.uniform(0.5, 1.5, F), rng.uniform(0.5, 1.5, n), bc,
                        rng.uniform(0.5, 1.5, len(bc)), rng.uniform(0.5, 1.5, len(bc)))


def check_apply(ctx, fm, coef_width, faithful_width, dirichlet=True):
    n = fm.n_cells
    cpu = orc.FaceOp(fm, prefill=1, dt=-0.05, dirichlet=dirichlet)
    rng = np.random.default_rng(5)
    xs = [rng.standard_normal(n), np.sin(0.37 * np.arange(n))]
    # coefficient form (the streaming TMA kernel): rows and apply against the row oracle
    gpu = sb.FvmOperator(ctx, fm, prefill=1, dt=-0.05, form=sb.FORM_COEF, dirichlet=dirichlet)
    assert gpu.info.width == coef_width
    col, a, _, diag = gpu.rows()
    w, ld, ocol, oa, odiag = cpu.rows_coef(ld=gpu.info.ld)
    wi = min(w, coef_width)
    assert (ocol[wi:] == orc.COL_PAD).all() and (col[wi:] == orc.COL_PAD).all()   # ghost rows / even-width padding
    assert np.array_equal(col[:wi], ocol[:wi]) and np.array_equal(a[:wi], oa[:wi]) and np.array_equal(diag, odiag)
    assert gpu.info.algorithmic_bytes_per_apply == 24 * n + 12 * int((ocol != orc.COL_PAD).sum())
    rows = cpu.rows_coef()
    y = ctx.zeros(n)
    for x in xs:
        gpu.mul(y, ctx.vector(x))
        got = y.numpy()
        assert np.array_equal(got, cpu.apply_rows_coef(x, rows))
        ref = cpu.apply(x)
        assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()
    # faithful form: bit-identical to the reference's face loop
    gpu_f = sb.FvmOperator(ctx, fm, prefill=1, dt=-0.05, form=sb.FORM_FAITHFUL, dirichlet=dirichlet)
    assert gpu_f.info.width == faithful_width
    for x in xs:
        gpu_f.mul(y, ctx.vector(x))
        assert np.array_equal(y.numpy(), cpu.apply(x))
    return cpu, gpu, gpu_f


@pytest.mark.parametrize("k,odd,n_ghost,coef_w,faith_w", [
    (4, True, 0, 10, 10),      # degree 9  -> padded to 10
    (5, False, 0, 10, 10),
    (5, True, 0, 12, 12),      # degree 11 -> 12
    (6, False, 300, 12, 14),   # ghosts widen the faithful rows only (they fold into the diagonal of the coef form)
    (6, True, 0, 14, 14),      # degree 13 -> 14
    (7, False, 0, 14, 14),
    (7, True, 0, 16, 16),      # degree 15 -> 16
    (8, False, 0, 16, 16),
])
def test_wide_rows_apply_bit_exact(ctx, k, odd, n_ghost, coef_w, faith_w):
    fm = band_mesh(7000, k, odd, n_ghost, seed=100 + k)
    check_apply(ctx, fm, coef_w, faith_w)


def test_more_than_16_faces_is_rejected(ctx):
    fm = band_mesh(3000, 8, True, 0, seed=1)   # degree 17
    with pytest.raises(sb.StormB200Error, match="16"):
        sb.FvmOperator(ctx, fm, prefill=1, dt=-0.05, form=sb.FORM_COEF)


@pytest.mark.parametrize("stretch", [(1.0, 1.0, 1.0), (1.0, 1.3, 0.7)])
def test_polyhedral_mesh_apply_and_solvers(ctx, stretch):
    m = PolyMesh.bcc(20, stretch)       # 16 000 truncated octahedra, 8 tiles
    fm = face_mesh(m)
    cpu, gpu, gpu_f = check_apply(ctx, fm, 14, 14)
    n = m.n_cells
    bh = np.sin(0.37 * np.arange(n))
    b = ctx.vector(bh)
    for name, Solver in (("cg", sb.CgSolver), ("bicgstab", sb.BiCgStabSolver)):
        want = orc.solve(name, cpu, bh, num_iterations=60, abs_tol=0.0, rel_tol=1e-10, mode=orc.RED_TREE)
        for use_graph in (False, True):
            s = Solver(num_iterations=60, absolute_error_tolerance=0.0, relative_error_tolerance=1e-10, use_graph=use_graph)
            x = ctx.zeros(n)
            conv = s.solve(x, b, gpu_f)
            assert conv == want.converged and s.iteration == want.iterations
            assert np.array_equal(s.history, want.hist) and np.array_equal(x.numpy(), want.x)
        rows_op = orc.RowsOp(n, *cpu.rows_coef())
        want = orc.solve(name, rows_op, bh, num_iterations=60, abs_tol=0.0, rel_tol=1e-10, mode=orc.RED_TREE)
        s = Solver(num_iterations=60, absolute_error_tolerance=0.0, relative_error_tolerance=1e-10, use_graph=True)
        x = ctx.zeros(n)
        s.solve(x, b, gpu)
        assert s.iteration == want.iterations
        assert np.array_equal(s.history, want.hist) and np.array_equal(x.numpy(), want.x)


def test_polyhedral_mesh_poisson_converges_to_the_exact_solution(ctx):
    """Dirichlet Poisson on the truncated-octahedra mesh: the discrete solution of A x = A x* is x*."""
    m = PolyMesh.bcc(16)
    n = m.n_cells
    c = m.cell_centers()
    x_star = np.sin(np.pi * c[:, 0]) * np.sin(np.pi * c[:, 1]) * np.sin(np.pi * c[:, 2])
    gpu = sb.FvmOperator(ctx, m, prefill=0, dt=-1.0, form=sb.FORM_COEF, dirichlet=True)
    b = ctx.zeros(n)
    gpu.mul(b, ctx.vector(x_star))
    s = sb.CgSolver(num_iterations=2000, absolute_error_tolerance=0.0, relative_error_tolerance=1e-12, use_graph=True)
    x = ctx.zeros(n)
    assert s.solve(x, b, gpu)
    assert np.linalg.norm(x.numpy() - x_star) / np.linalg.norm(x_star) < 1e-8


def test_polyhedral_mesh_at_scale(ctx):
    """1.46 M truncated octahedra (257 MB of operator: beyond L2, many waves of the 1-CTA/SM wide kernel): apply
    against the row oracle over repeated launches, BiCGStab against the tree oracle, run-to-run determinism."""
    m = PolyMesh.bcc(90)
    fm = face_mesh(m)
    n = m.n_cells
    gpu = sb.FvmOperator(ctx, m, prefill=0, dt=-1.0, form=sb.FORM_COEF, dirichlet=True)
    assert gpu.info.width == 14
    cpu = orc.FaceOp(fm, prefill=0, dt=-1.0, dirichlet=True)
    rows_op = orc.RowsOp(n, *cpu.rows_coef())
    bh = np.sin(0.37 * np.arange(n))
    b, y = ctx.vector(bh), ctx.zeros(n)
    want = rows_op.apply(bh)
    for _ in range(6):
        gpu.mul(y, b)
        assert np.array_equal(y.numpy(), want), "wide apply differs from the row oracle at 1.46 M cells"
    w = orc.solve("bicgstab", rows_op, bh, num_iterations=20, abs_tol=0.0, rel_tol=0.0, mode=orc.RED_TREE)
    for use_graph in (False, True, True):
        s = sb.BiCgStabSolver(num_iterations=20, absolute_error_tolerance=0.0, relative_error_tolerance=0.0,
                              use_graph=use_graph)
        x = ctx.zeros(n)
        s.solve(x, b, gpu)
        assert np.array_equal(s.history, w.hist) and np.array_equal(x.numpy(), w.x)
