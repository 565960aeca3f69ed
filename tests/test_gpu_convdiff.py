"""Convection-diffusion on the GPU (BASELINE.json configs[2]): the non-symmetric
upwind operator in the face-loop pattern of UpwindConvectionScheme (Feathers/ConvectionScheme.hpp:83-106),
stored as coefficient rows and applied by the same kernels; GMRES(m) / FGMRES / BiCGStab / IDR(s) are the
reference's own templates on Storm::DeviceVector.

Oracle: oracle/sb_oracle.c (orc_apply_convdiff_faces: the face loop; orc_rows_convdiff: the row form whose
operation order is the layout contract). Rows and apply are bit-exact against the row oracle and within
rounding of the face loop; the solvers are bit-exact against the same reference headers on a host vector
with the reduction tree matched, and within the north_star bars of the reference's sequential sums.
"""
import numpy as np
import pytest

import stormruler_b200 as sb
from oracle import orc
from stormruler_b200 import dropin
from stormruler_b200.mesh import CELL_HEX, CELL_TET, Mesh

pytestmark = pytest.mark.gpu

NU, BETA = 0.02, (1.0, 0.5, 0.25)
X_TOL = 1e-8


def make_case(ctx, kind, dims):
    mesh = Mesh.box(CELL_TET if kind == "tet" else CELL_HEX, *dims)
    mesh.renumber_rcm()
    fu, bu = mesh.face_flux(BETA)
    fm = orc.FaceMesh(mesh.n_cells, mesh.face_cell, mesh.face_area, mesh.face_dist, mesh.cell_vol, mesh.bface_cell,
                      mesh.bface_area, mesh.bface_dist)
    cpu = orc.ConvDiffOp(fm, NU, fu, bu)
    gpu = sb.ConvDiffOperator(ctx, mesh, NU, fu, bu)
    return mesh, cpu, gpu


@pytest.fixture(scope="module")
def tet_case(ctx):
    return make_case(ctx, "tet", (12, 10, 9))


@pytest.mark.parametrize("kind,dims", [("tet", (7, 6, 5)), ("hex", (11, 9, 8))])
def test_rows_and_apply_bit_exact(ctx, kind, dims):
    mesh, cpu, gpu = make_case(ctx, kind, dims)
    rows = cpu.rows_coef(ld=gpu.info.ld)
    w, ld, ocol, oa, odiag = rows.rows
    col, a, _, diag = gpu.rows()
    assert gpu.info.width == w and gpu.info.form == sb.FORM_COEF
    assert np.array_equal(col, ocol) and np.array_equal(a, oa) and np.array_equal(diag, odiag)
    assert gpu.info.algorithmic_bytes_per_apply == 24 * cpu.n + 12 * 2 * mesh.n_faces
    rng = np.random.default_rng(5)
    y = ctx.zeros(cpu.n)
    tight = cpu.rows_coef()
    for x in (rng.standard_normal(cpu.n), np.ones(cpu.n), np.zeros(cpu.n)):
        gpu.mul(y, ctx.vector(x))
        got = y.numpy()
        assert np.array_equal(got, tight.apply(x))
        ref = cpu.apply(x)
        assert np.abs(got - ref).max() <= 1e-13 * max(np.abs(ref).max(), 1e-300)


def test_create_rejects_bad_input(ctx):
    mesh = Mesh.box(CELL_TET, 2, 2, 2)
    fu, bu = mesh.face_flux(BETA)
    with pytest.raises(sb.StormB200Error):
        sb.ConvDiffOperator(ctx, mesh, -1.0, fu, bu)          # negative diffusion coefficient
    with pytest.raises(AssertionError):
        sb.ConvDiffOperator(ctx, mesh, NU, fu[:-1], bu)       # wrong face count


@pytest.mark.parametrize("solver,num_inner", [("gmres", 0), ("gmres", 30), ("fgmres", 0), ("bicgstab", 0), ("bicgstabl", 0),
                                              ("idrs", 0), ("tfqmr", 0), ("cgs", 0)])
def test_reference_templates_on_convdiff_bit_identical(ctx, tet_case, solver, num_inner):
    """Same reference headers, host vector + row oracle vs device vector + CUDA rows, reduction tree matched."""
    assert dropin.available() and orc.have_ref()
    _, cpu, gpu = tet_case
    rows = cpu.rows_coef()
    b = np.sin(0.37 * np.arange(cpu.n))
    want = orc.ref_solve(solver, rows, b, num_iterations=400, abs_tol=0.0, rel_tol=1e-10, num_inner=num_inner,
                         mode=orc.RED_TREE)
    x = ctx.zeros(cpu.n)
    got = dropin.solve(solver, gpu, x, ctx.vector(b), num_iterations=400, abs_tol=0.0, rel_tol=1e-10, num_inner=num_inner)
    assert (got.converged, got.iterations, got.n_apply) == (want.converged, want.iterations, want.n_apply)
    assert np.array_equal(got.trace, want.trace) and np.array_equal(got.hist, want.hist)
    assert np.array_equal(x.numpy(), want.x)
    assert want.converged
    # against the reference's own sequential sums on the FACE LOOP (row rounding + reduction order differ)
    seq = orc.ref_solve(solver, cpu, b, num_iterations=400, abs_tol=0.0, rel_tol=1e-10, num_inner=num_inner)
    assert seq.converged
    assert np.linalg.norm(x.numpy() - seq.x) <= X_TOL * np.linalg.norm(seq.x)
    k = min(10, len(seq.hist), len(got.hist))
    assert (np.abs(got.hist[:k] - seq.hist[:k]) <= 1e-10 * seq.hist[:k]).all()


@pytest.mark.parametrize("use_graph", [False, True])
def test_fused_bicgstab_on_convdiff_bit_identical(ctx, tet_case, use_graph):
    _, cpu, gpu = tet_case
    rows = cpu.rows_coef()
    b = np.sin(0.37 * np.arange(cpu.n))
    want = orc.solve("bicgstab", rows, b, num_iterations=300, abs_tol=0.0, rel_tol=1e-10, mode=orc.RED_TREE)
    s = sb.BiCgStabSolver(num_iterations=300, absolute_error_tolerance=0.0, relative_error_tolerance=1e-10,
                          use_graph=use_graph)
    x = ctx.zeros(cpu.n)
    conv = s.solve(x, ctx.vector(b), gpu)
    assert conv and conv == want.converged and s.iteration == want.iterations
    assert np.array_equal(s.history, want.hist) and np.array_equal(x.numpy(), want.x)
    res = np.linalg.norm(b - cpu.apply(x.numpy())) / np.linalg.norm(b)
    assert res < 1e-9


# ---- the preconditioner slot: Storm::JacobiPreconditioner in IterativeSolver::pre_op ----
def test_jacobi_apply_bit_exact(ctx, tet_case):
    _, cpu, gpu = tet_case
    w, ld, col, a, diag = cpu.rows_coef().rows
    rng = np.random.default_rng(9)
    x = rng.standard_normal(cpu.n)
    xd, yd = ctx.vector(x), ctx.zeros(cpu.n)
    gpu.jacobi(yd, xd)
    assert np.array_equal(yd.numpy(), x / diag[:cpu.n])
    gpu.jacobi(xd, xd)                                   # in place
    assert np.array_equal(xd.numpy(), x / diag[:cpu.n])
    lap = sb.FvmOperator(ctx, tet_case[0], prefill=0, dt=-1.0, form=sb.FORM_FAITHFUL, dirichlet=True)
    with pytest.raises(sb.StormB200Error):
        lap.jacobi(yd, xd)                               # the faithful form keeps no diagonal


@pytest.mark.parametrize("solver", ["bicgstab", "gmres", "fgmres", "idrs", "tfqmr", "bicgstabl", "cgs"])
@pytest.mark.parametrize("side", ["right", "left"])
def test_reference_templates_with_jacobi_preconditioner_bit_identical(ctx, tet_case, solver, side):
    """Left/right/flexible preconditioned branches of the reference solvers (e.g. SolverBiCgStab.hpp:135-137,
    SolverGmres.hpp:149-156,233-248), device vs host, reduction tree matched: every reduction scalar, the
    residual history and the solution agree bit for bit; Jacobi never needs more iterations than no preconditioner."""
    _, cpu, gpu = tet_case
    rows = cpu.rows_coef()
    diag = rows.rows[4][:cpu.n]
    b = np.sin(0.37 * np.arange(cpu.n))
    want = orc.ref_solve(solver, rows, b, num_iterations=400, abs_tol=0.0, rel_tol=1e-10, mode=orc.RED_TREE,
                         pre=orc.JacobiOp(diag), pre_side=side)
    x = ctx.zeros(cpu.n)
    got = dropin.solve(solver, gpu, x, ctx.vector(b), num_iterations=400, abs_tol=0.0, rel_tol=1e-10,
                       precond="jacobi", pre_side=side)
    assert want.converged
    assert (got.converged, got.iterations, got.n_apply) == (want.converged, want.iterations, want.n_apply)
    assert np.array_equal(got.trace, want.trace) and np.array_equal(got.hist, want.hist)
    assert np.array_equal(x.numpy(), want.x)
    plain = orc.ref_solve(solver, rows, b, num_iterations=400, abs_tol=0.0, rel_tol=1e-10, mode=orc.RED_TREE)
    assert want.iterations <= plain.iterations
    assert np.linalg.norm(b - cpu.apply(x.numpy())) <= 1e-8 * np.linalg.norm(b)


# ---- fused GMRES(m): device-resident Arnoldi process, host-side Givens bookkeeping (sb_gmres_solve) ----
@pytest.mark.parametrize("m,lookahead", [(50, 0), (7, 0), (7, 1), (13, 6), (1, 0)])
def test_fused_gmres_bit_identical_to_reference_headers(ctx, tet_case, m, lookahead):
    _, cpu, gpu = tet_case
    rows = cpu.rows_coef()
    b = np.sin(0.37 * np.arange(cpu.n))
    iters = 400 if m > 1 else 60
    want = orc.ref_solve("gmres", rows, b, num_iterations=iters, abs_tol=0.0, rel_tol=1e-10, num_inner=m, mode=orc.RED_TREE)
    s = sb.GmresSolver(num_iterations=iters, absolute_error_tolerance=0.0, relative_error_tolerance=1e-10,
                       num_inner_iterations=m, lookahead=lookahead)
    x = ctx.zeros(cpu.n)
    conv = s.solve(x, ctx.vector(b), gpu)
    assert (conv, s.iteration) == (want.converged, want.iterations)
    assert len(s.trace) == len(want.trace) and np.array_equal(s.trace, want.trace)
    assert np.array_equal(s.history, want.hist)
    assert s.absolute_error == want.abs_err and s.relative_error == want.rel_err
    assert np.array_equal(x.numpy(), want.x)
    if want.converged:
        assert np.linalg.norm(b - cpu.apply(x.numpy())) <= 1e-8 * np.linalg.norm(b)
    # FGMRES without a preconditioner is the same algorithm
    f = orc.ref_solve("fgmres", rows, b, num_iterations=iters, abs_tol=0.0, rel_tol=1e-10, num_inner=m, mode=orc.RED_TREE)
    assert np.array_equal(f.x, want.x)


def test_fused_gmres_fixed_count_stops_mid_cycle_and_nonzero_guess(ctx, tet_case):
    _, cpu, gpu = tet_case
    rows = cpu.rows_coef()
    b = np.sin(0.37 * np.arange(cpu.n))
    x0 = np.cos(0.11 * np.arange(cpu.n))
    want = orc.ref_solve("gmres", rows, b, x0=x0, num_iterations=61, abs_tol=0.0, rel_tol=0.0, num_inner=9, mode=orc.RED_TREE)
    s = sb.GmresSolver(num_iterations=61, absolute_error_tolerance=0.0, relative_error_tolerance=0.0, num_inner_iterations=9)
    x = ctx.vector(x0)
    assert s.solve(x, ctx.vector(b), gpu) is False and s.iteration == 61
    assert np.array_equal(s.trace, want.trace) and np.array_equal(s.history, want.hist) and np.array_equal(x.numpy(), want.x)
    # through the reference's abstract Solver interface (Storm::B200::GmresSolver in the C++ drop-in)
    x = ctx.vector(x0)
    got = dropin.solve("fused_gmres", gpu, x, ctx.vector(b), num_iterations=61, abs_tol=0.0, rel_tol=0.0, num_inner=9)
    assert got.iterations == 61 and np.array_equal(got.hist, want.hist) and np.array_equal(x.numpy(), want.x)


def test_fused_gmres_early_exit_leaves_x_untouched(ctx, tet_case):
    """Solver.hpp:124-128 + SolverGmres.hpp:207-212: the reference back-substitutes through an all-zero H here and
    returns NaN -- shown on the host vector; the fused path keeps x (documented deviation)."""
    _, cpu, gpu = tet_case
    rows = cpu.rows_coef()
    b = np.sin(0.37 * np.arange(cpu.n))
    ref = orc.ref_solve("gmres", rows, b, num_iterations=10, abs_tol=1e30, rel_tol=0.0)
    assert ref.converged and ref.iterations == 0 and not np.isfinite(ref.x).all()
    s = sb.GmresSolver(num_iterations=10, absolute_error_tolerance=1e30, relative_error_tolerance=0.0)
    x = ctx.zeros(cpu.n)
    assert s.solve(x, ctx.vector(b), gpu) is True and s.iteration == 0
    assert np.array_equal(x.numpy(), np.zeros(cpu.n))
    with pytest.raises(sb.StormB200Error):
        sb.GmresSolver(num_inner_iterations=500).solve(x, ctx.vector(b), gpu)   # restart length out of range
