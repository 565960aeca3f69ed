"""Parity and determinism at sizes where the grid spans many waves of CTAs (the small reference meshes are 3
tiles): 1.2 M tetrahedra against the oracle bit for bit, 6 M tetrahedra run-to-run. Guards the properties the
small cases cannot see -- inter-CTA ordering, the final-reduce stage over thousands of partials, graph replay, the
persistent kernel's grid barriers with several tiles per CTA."""
import os

import numpy as np
import pytest

import stormruler_b200 as sb
from oracle import orc
from stormruler_b200.mesh import CELL_HEX, CELL_TET, Mesh

pytestmark = pytest.mark.gpu


def box(n, kind=CELL_TET):
    mesh = Mesh.box(kind, n, jitter=0.2, seed_jitter=42, shuffle=True, seed_shuffle=43)
    mesh.renumber_rcm()
    fm = orc.FaceMesh(mesh.n_cells, mesh.face_cell, mesh.face_area, mesh.face_dist, mesh.cell_vol, mesh.bface_cell,
                      mesh.bface_area, mesh.bface_dist)
    return mesh, fm


@pytest.fixture(scope="module")
def mesh_1m():
    return box(58)   # 1 170 672 cells = 572 tiles


@pytest.fixture(scope="module")
def any_ctx():
    c = sb.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("solver", ["cg", "bicgstab"])
def test_million_cell_solvers_bit_identical_to_oracle(any_ctx, mesh_1m, solver):
    ctx = any_ctx
    mesh, fm = mesh_1m
    n = mesh.n_cells
    c = mesh.cell_centers()
    x_star = np.sin(np.pi * c[:, 0]) * np.sin(np.pi * c[:, 1]) * np.sin(np.pi * c[:, 2])
    cpu = orc.FaceOp(fm, prefill=0, dt=-1.0, dirichlet=True)
    rows = cpu.rows_coef()
    rows_op = orc.RowsOp(n, *rows)
    Solver = sb.CgSolver if solver == "cg" else sb.BiCgStabSolver
    iters = 40
    for form, oracle_op in ((sb.FORM_COEF, rows_op), (sb.FORM_FAITHFUL, cpu)):
        gpu = sb.FvmOperator(ctx, mesh, prefill=0, dt=-1.0, form=form, dirichlet=True)
        b_dev, xs = ctx.zeros(n), ctx.vector(x_star)
        gpu.mul(b_dev, xs)
        b = b_dev.numpy()
        assert np.array_equal(b, oracle_op.apply(x_star)), "apply differs from the oracle at 1.2 M cells"
        want = orc.solve(solver, oracle_op, b, num_iterations=iters, abs_tol=0.0, rel_tol=0.0, mode=orc.RED_TREE)
        # stepwise without / with graph replay, then (coefficient form) the persistent whole-solve kernel, twice
        for use_graph, schedule in ((False, 1), (True, 1), (True, 0), (False, 0)):
            s = Solver(num_iterations=iters, absolute_error_tolerance=0.0, relative_error_tolerance=0.0,
                       use_graph=use_graph, schedule=schedule)
            x = ctx.zeros(n)
            s.solve(x, b_dev, gpu)
            assert s.iteration == iters
            assert np.array_equal(s.history, want.hist), f"{solver} form {form} graph={use_graph}: residual history differs"
            assert np.array_equal(s.trace, want.trace[:len(s.trace)])
            assert np.array_equal(x.numpy(), want.x), f"{solver} form {form} graph={use_graph}: solution differs"


@pytest.mark.parametrize("solver", ["cg", "bicgstab"])
def test_six_million_cells_run_to_run_deterministic(any_ctx, solver):
    ctx = any_ctx
    mesh, fm = box(100)   # 6 000 000 cells, 2930 tiles, ~6.6 waves of the apply kernel: DRAM-bound, unlike 1 M cells
    n = mesh.n_cells
    gpu = sb.FvmOperator(ctx, mesh, prefill=0, dt=-1.0, form=sb.FORM_COEF, dirichlet=True)
    bh = np.sin(0.37 * np.arange(n))
    b = ctx.vector(bh)
    # the apply itself, against the row oracle, several launches (a stage that reads the wrong bytes shows here)
    cpu = orc.FaceOp(fm, prefill=0, dt=-1.0, dirichlet=True)
    want = orc.RowsOp(n, *cpu.rows_coef()).apply(bh)
    y = ctx.zeros(n)
    for _ in range(8):
        gpu.mul(y, b)
        assert np.array_equal(y.numpy(), want), "apply differs from the row oracle at 6 M cells"
    Solver = sb.CgSolver if solver == "cg" else sb.BiCgStabSolver
    runs = []
    for use_graph, schedule in ((True, 1), (True, 1), (False, 1), (True, 0), (True, 0)):   # 0: persistent kernel
        s = Solver(num_iterations=50, absolute_error_tolerance=0.0, relative_error_tolerance=0.0, use_graph=use_graph,
                   schedule=schedule)
        x = ctx.zeros(n)
        s.solve(x, b, gpu)
        runs.append((s.history.copy(), x.numpy()))
    for hist, x in runs[1:]:
        assert np.array_equal(hist, runs[0][0]), "residual history is not run-to-run deterministic"
        assert np.array_equal(x, runs[0][1]), "solution is not run-to-run deterministic"


def test_hexahedra_width_six_rows_at_scale(ctx):
    """The 6-wide instantiation of the TMA kernel (hexahedra; 5.6 KB stages, 2 CTAs/SM) on 3.4 M cells: apply against
    the row oracle over repeated launches, BiCGStab and GMRES against the oracle / the reference headers."""
    mesh, fm = box(150, CELL_HEX)
    n = mesh.n_cells
    gpu = sb.FvmOperator(ctx, mesh, prefill=0, dt=-1.0, form=sb.FORM_COEF, dirichlet=True)
    assert gpu.info.width == 6
    cpu = orc.FaceOp(fm, prefill=0, dt=-1.0, dirichlet=True)
    rows_op = orc.RowsOp(n, *cpu.rows_coef())
    bh = np.sin(0.37 * np.arange(n))
    b, y = ctx.vector(bh), ctx.zeros(n)
    want = rows_op.apply(bh)
    for _ in range(6):
        gpu.mul(y, b)
        assert np.array_equal(y.numpy(), want)
    w = orc.solve("bicgstab", rows_op, bh, num_iterations=25, abs_tol=0.0, rel_tol=0.0, mode=orc.RED_TREE)
    for use_graph, schedule in ((False, 1), (True, 1), (False, 2)):   # 2: the persistent kernel, two CTAs per SM at this width
        s = sb.BiCgStabSolver(num_iterations=25, absolute_error_tolerance=0.0, relative_error_tolerance=0.0, use_graph=use_graph,
                              schedule=schedule)
        x = ctx.zeros(n)
        s.solve(x, b, gpu)
        assert np.array_equal(s.history, w.hist) and np.array_equal(x.numpy(), w.x)
    g = sb.GmresSolver(num_iterations=24, absolute_error_tolerance=0.0, relative_error_tolerance=0.0, num_inner_iterations=10)
    x = ctx.zeros(n)
    g.solve(x, b, gpu)
    wg = orc.ref_solve("gmres", rows_op, bh, num_iterations=24, abs_tol=0.0, rel_tol=0.0, num_inner=10, mode=orc.RED_TREE)
    assert np.array_equal(g.trace, wg.trace) and np.array_equal(g.history, wg.hist) and np.array_equal(x.numpy(), wg.x)
