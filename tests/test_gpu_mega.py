"""The persistent whole-solve kernel (csrc/sb_mega.cuh, SB_SCHEDULE_PERSISTENT) against the oracle and against the
one-kernel-per-step schedule: same bodies, same reduction tree, so every residual, every reduction scalar and the
solution must be bit-identical -- on grids of 3 CTAs (the reference's 2-D meshes) and on grids where every CTA owns
several tiles, for every row width the kernel is instantiated for that the meshes here produce (3, 4, 6, 14)."""
import numpy as np
import pytest

import stormruler_b200 as sb
from conftest import rhs
from oracle import orc
from stormruler_b200 import capi
from stormruler_b200.mesh import CELL_HEX, CELL_TET, Mesh, PolyMesh

pytestmark = pytest.mark.gpu

SOLVERS = {"cg": sb.CgSolver, "bicgstab": sb.BiCgStabSolver}
STEP, PERS = capi.SCHEDULE_STEPWISE, capi.SCHEDULE_PERSISTENT


def face_mesh(m):
    return orc.FaceMesh(m.n_cells, m.face_cell, m.face_area, m.face_dist, m.cell_vol, m.bface_cell, m.bface_area,
                        m.bface_dist)


def problem(ctx, name):
    """(n, oracle rows operator, device operator) of a 3-D Poisson problem in the coefficient form."""
    if name == "tet":
        mesh = Mesh.box(CELL_TET, 30, 28, 26, jitter=0.2, seed_jitter=42, shuffle=True, seed_shuffle=43)   # 131 040 cells
    elif name == "tet_large":
        mesh = Mesh.box(CELL_TET, 80, jitter=0.2, seed_jitter=42, shuffle=True, seed_shuffle=43)           # 3.07 M cells, 1500 tiles
    elif name == "hex":
        mesh = Mesh.box(CELL_HEX, 70, 64, 60, jitter=0.2, seed_jitter=42, shuffle=True, seed_shuffle=43)    # 268 800 cells
    else:
        mesh = PolyMesh.bcc(40, stretch=(1.0, 1.3, 0.7)).to_mesh()                                            # 14-face cells
        mesh.permute_cells(np.random.default_rng(43).permutation(mesh.n_cells).astype(np.int32))
    mesh.renumber_rcm()
    cpu = orc.FaceOp(face_mesh(mesh), prefill=0, dt=-1.0, dirichlet=True)
    gpu = sb.FvmOperator(ctx, mesh, prefill=0, dt=-1.0, form=sb.FORM_COEF, dirichlet=True)
    return mesh.n_cells, orc.RowsOp(mesh.n_cells, *cpu.rows_coef()), gpu


def run(ctx, gpu, solver, b_dev, schedule, iters, rel_tol=0.0, abs_tol=0.0, **kw):
    s = SOLVERS[solver](num_iterations=iters, absolute_error_tolerance=abs_tol, relative_error_tolerance=rel_tol,
                        schedule=schedule, **kw)
    x = ctx.zeros(gpu.n)
    conv = s.solve(x, b_dev, gpu)
    return s, conv, x.numpy()


def same(a, b):
    return (a[1] == b[1] and a[0].iteration == b[0].iteration and np.array_equal(a[0].history, b[0].history)
            and np.array_equal(a[0].trace, b[0].trace) and np.array_equal(a[2], b[2])
            and a[0].absolute_error == b[0].absolute_error and a[0].relative_error == b[0].relative_error)


@pytest.mark.parametrize("solver", ["cg", "bicgstab"])
@pytest.mark.parametrize("name", ["tet", "hex", "poly", "tet_large"])
def test_persistent_schedule_bit_identical_to_oracle_and_to_stepwise(ctx, name, solver):
    n, rows_op, gpu = problem(ctx, name)
    bh = rhs(n)
    b = ctx.vector(bh)
    iters = 30 if name == "tet_large" else 60
    want = orc.solve(solver, rows_op, bh, num_iterations=iters, abs_tol=0.0, rel_tol=0.0, mode=orc.RED_TREE)
    pers = run(ctx, gpu, solver, b, PERS, iters)
    assert pers[0].schedule_used == PERS and pers[0].iteration == iters
    assert np.array_equal(pers[0].history, want.hist), "residual history differs from the oracle"
    k = min(len(pers[0].trace), len(want.trace))   # the fused BiCGStab computes the next rho one step early
    assert k >= len(want.trace) - 1 and np.array_equal(pers[0].trace[:k], want.trace[:k])
    assert np.array_equal(pers[2], want.x), "solution differs from the oracle"
    step = run(ctx, gpu, solver, b, STEP, iters, use_graph=True)
    assert step[0].schedule_used == STEP and same(pers, step)
    # one launch for the whole loop: initialisation (apply + final stage + copy) + 1; the stepwise schedule: 3 (CG) or
    # 5 (BiCGStab) kernels per iteration + a one-CTA final stage per reduction (2 / 3); folded: nothing in between
    assert pers[0].launches <= 4 < step[0].launches <= (5 if solver == "cg" else 8) * (iters + 1) + 6
    fold = run(ctx, gpu, solver, b, capi.SCHEDULE_FOLDED, iters, use_graph=True)
    assert fold[0].schedule_used == capi.SCHEDULE_FOLDED and same(pers, fold)
    assert fold[0].launches <= (3 if solver == "cg" else 5) * (iters + 1) + 6
    # the tuning bits of the stepwise schedule change how kernels are launched and what the L2 keeps, never the bits
    for tuning in (capi.TUNE_OFF, capi.TUNE_STREAM_OPERATOR, capi.TUNE_PDL_FINAL | capi.TUNE_PDL_AFTER_FINAL,
                   capi.TUNE_STREAM_OPERATOR | capi.TUNE_PDL_FINAL | capi.TUNE_PDL_AFTER_FINAL | capi.TUNE_PDL_APPLY,
                   capi.TUNE_IN_KERNEL_REDUCER,
                   capi.TUNE_IN_KERNEL_REDUCER | capi.TUNE_STREAM_OPERATOR | capi.TUNE_PDL_AFTER_FINAL | capi.TUNE_PDL_APPLY):
        for graph in (True, False):
            got = run(ctx, gpu, solver, b, STEP, iters, use_graph=graph, tuning=tuning)
            assert same(pers, got), (tuning, graph)
            if tuning & capi.TUNE_IN_KERNEL_REDUCER:   # one-CTA kernels only behind the applies (1 / 2 per iteration)
                assert got[0].launches <= (4 if solver == "cg" else 7) * (iters + 1) + 6
    # stops in the middle of the loop, on the relative tolerance: same iterate as the stepwise schedule, and again
    # when the two schedules alternate on one context (the all-reduce mailbox parity carries over)
    rel = want.hist / want.hist[0]
    tol = float(rel[1:].min()) * (1.0 + 1e-9)            # reached for the first time somewhere inside the loop
    stop_at = int(np.flatnonzero(rel < tol)[0])
    runs = [run(ctx, gpu, solver, b, sch, iters, rel_tol=tol, use_graph=True) for sch in (PERS, STEP, PERS, PERS, STEP)]
    assert runs[0][1] and runs[0][0].iteration == stop_at and 0 < stop_at <= iters
    assert all(same(runs[0], r) for r in runs[1:])


@pytest.mark.parametrize("solver", ["cg", "bicgstab"])
def test_reference_2d_meshes_three_ctas(ctx, square_nb, solver):
    """6 252 cells = 4 tiles: fewer CTAs than SMs, width-3 rows, the config-1 operator y = x - dt div grad x."""
    cpu = orc.FaceOp(square_nb, prefill=1, dt=-0.05)
    gpu = sb.FvmOperator(ctx, square_nb, prefill=1, dt=-0.05, form=sb.FORM_COEF)
    bh = rhs(cpu.n)
    rows_op = orc.RowsOp(cpu.n, *cpu.rows_coef())
    want = orc.solve(solver, rows_op, bh, num_iterations=500, abs_tol=0.0, rel_tol=1e-10, mode=orc.RED_TREE)
    s, conv, x = run(ctx, gpu, solver, ctx.vector(bh), PERS, 500, rel_tol=1e-10)
    assert (conv, s.iteration) == (want.converged, want.iterations)
    assert np.array_equal(s.history, want.hist) and np.array_equal(x, want.x)
    # the automatic choice is the stepwise schedule (reductions folded into their consumers), with and without graph
    # replay and with per-kernel timing: same bits
    for kw in ({}, {"use_graph": True}, {"profile": True}, {"use_graph": True, "check_every": 7}):
        s2, conv2, x2 = run(ctx, gpu, solver, ctx.vector(bh), capi.SCHEDULE_AUTO, 500, rel_tol=1e-10, **kw)
        assert s2.schedule_used == STEP and conv2 == conv and s2.iteration == s.iteration
        assert np.array_equal(s2.history, s.history) and np.array_equal(s2.trace, s.trace) and np.array_equal(x2, x)
        if kw.get("profile"):
            assert len(s2.kernel_ms) in (3, 5) and min(s2.kernel_ms) > 0 and min(s2.wait_ms) >= 0.0
            # the one-CTA final stage behind every reducing kernel has an event of its own: part of its slot, never all
            reducing = (0, 1) if solver == "cg" else (1, 3, 4)
            for k in range(len(s2.kernel_ms)):
                if k in reducing:
                    assert 0.0 < s2.final_ms[k] < s2.kernel_ms[k], (k, s2.final_ms, s2.kernel_ms)
                else:
                    assert s2.final_ms[k] == 0.0


def test_stopping_rules_persistent(ctx, square_nb):
    cpu = orc.FaceOp(square_nb, prefill=1, dt=-0.05)
    gpu = sb.FvmOperator(ctx, square_nb, prefill=1, dt=-0.05, form=sb.FORM_COEF)
    rows_op = orc.RowsOp(cpu.n, *cpu.rows_coef())
    bh = rhs(cpu.n)
    b = ctx.vector(bh)
    s, conv, x = run(ctx, gpu, "cg", b, PERS, 0)                                  # no iterations
    assert (conv, s.iteration, len(s.history)) == (False, 0, 1) and not x.any()
    s, conv, x = run(ctx, gpu, "bicgstab", b, PERS, 10, abs_tol=1e9)              # early exit, Solver.hpp:124-128
    assert (conv, s.iteration, len(s.history)) == (True, 0, 1) and not x.any()
    for solver, iters in (("cg", 17), ("bicgstab", 9), ("bicgstab", 1)):          # iteration cap (odd and even counts)
        s, conv, x = run(ctx, gpu, solver, b, PERS, iters)
        want = orc.solve(solver, rows_op, bh, num_iterations=iters, abs_tol=0.0, rel_tol=0.0, mode=orc.RED_TREE)
        assert (conv, s.iteration) == (False, iters) and np.array_equal(x, want.x) and np.array_equal(s.history, want.hist)
    s, conv, x = run(ctx, gpu, "cg", b, PERS, 500, abs_tol=1e-3)                  # absolute tolerance
    want = orc.solve("cg", rows_op, bh, num_iterations=500, abs_tol=1e-3, rel_tol=0.0, mode=orc.RED_TREE)
    assert conv and s.iteration == want.iterations and np.array_equal(x, want.x)
    s, conv, x = run(ctx, gpu, "cg", ctx.zeros(cpu.n), PERS, 3)                   # b = 0: safe_divide masks the breakdown
    assert not np.isnan(x).any() and not x.any()


def test_persistent_needs_the_coefficient_form(ctx, square_nb):
    gpu = sb.FvmOperator(ctx, square_nb, prefill=1, dt=-0.05, form=sb.FORM_FAITHFUL)
    with pytest.raises(sb.StormB200Error):
        run(ctx, gpu, "cg", ctx.vector(rhs(gpu.n)), PERS, 5)
    s, _, _ = run(ctx, gpu, "cg", ctx.vector(rhs(gpu.n)), capi.SCHEDULE_AUTO, 5)   # automatic: stepwise
    assert s.schedule_used == STEP


def test_solver_rejects_vectors_shorter_than_the_operator(ctx, square_nb):
    gpu = sb.FvmOperator(ctx, square_nb, prefill=1, dt=-0.05, form=sb.FORM_COEF)
    short = ctx.zeros(gpu.n - 3000)   # a smaller padded capacity than the operator's rows
    with pytest.raises(sb.StormB200Error):
        sb.CgSolver(num_iterations=3).solve(short, ctx.vector(rhs(gpu.n)), gpu)
    with pytest.raises(sb.StormB200Error):
        sb.BiCgStabSolver(num_iterations=3).solve(ctx.zeros(gpu.n), short, gpu)


def test_timeline_is_monotone_and_covers_the_loop(ctx):
    n, rows_op, gpu = problem(ctx, "tet")
    b = ctx.vector(rhs(n))
    for solver, barriers in (("bicgstab", 5), ("cg", 3)):
        s, _, _ = run(ctx, gpu, solver, b, PERS, 12, timeline_iters=8)
        tl = s.timeline.astype(np.int64)
        assert tl.shape == (8, capi.TIMELINE_WORDS)
        stamps = tl[:, :barriers + 1]
        assert (np.diff(stamps, axis=1) > 0).all(), "barrier stamps of an iteration are not increasing"
        assert (stamps[1:, 0] >= stamps[:-1, barriers]).all(), "iterations overlap"
        span_ms = (stamps[-1, barriers] - stamps[0, 0]) * 1e-6
        assert 0 < span_ms <= s.iter_ms * 1.05 + 0.05
        assert (tl[:, 6:6 + barriers] > 0).all() and (tl[:, 6:6 + barriers] < 50_000_000).all()
