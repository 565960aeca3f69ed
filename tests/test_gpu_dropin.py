"""The C++23 drop-in on the GPU: StormRuler's own solver templates (all ten) instantiated on
Storm::DeviceVector (stormruler_b200/host), every vector statement a CUDA kernel behind the C ABI.

Oracle: the SAME reference headers on a host vector (oracle/_ref, built where /root/reference
exists; the built library travels to the GPU box). With the reduction order matched
(ORC_RED_TREE) the two runs must agree BIT FOR BIT in iteration count, every reduction scalar
(the full dot/norm trace, i.e. every alpha/beta/omega input), the residual history and the
solution; against the reference's sequential sums (golden file) the north_star bars apply.
"""
import numpy as np
import pytest

import stormruler_b200 as sb
from conftest import load_golden, rhs
from oracle import orc
from stormruler_b200 import dropin

pytestmark = pytest.mark.gpu

DT, ITERS, RTOL = 0.05, 500, 1e-10
HIST_TOL, X_TOL = 1e-10, 1e-8


@pytest.fixture(scope="module")
def ops(ctx, square_nb):
    assert dropin.available(), "libstorm_dropin.so must be built before the GPU run (make -C stormruler_b200/host)"
    assert orc.have_ref(), "oracle/_ref must be built before the GPU run (make -C oracle)"
    cpu = orc.FaceOp(square_nb, prefill=1, dt=-DT)
    gpu = sb.FvmOperator(ctx, square_nb, prefill=1, dt=-DT, form=sb.FORM_FAITHFUL)
    return cpu, gpu


def run_both(ctx, ops, solver, iters=ITERS, rel_tol=RTOL, abs_tol=0.0, num_inner=0, b=None):
    cpu, gpu = ops
    b = rhs(cpu.n) if b is None else b
    relax = 0.0   # Richardson keeps the reference's default relaxation factor (1e-4)
    want = orc.ref_solve(solver, cpu, b, num_iterations=iters, abs_tol=abs_tol, rel_tol=rel_tol, num_inner=num_inner,
                         mode=orc.RED_TREE, relaxation_factor=relax)
    x = ctx.zeros(cpu.n)
    got = dropin.solve(solver, gpu, x, ctx.vector(b), num_iterations=iters, abs_tol=abs_tol, rel_tol=rel_tol,
                       num_inner=num_inner, relaxation_factor=relax)
    return want, got, x.numpy()


@pytest.mark.parametrize("solver", dropin.GENERIC_SOLVERS)
def test_reference_templates_on_device_vector_bit_identical(ctx, ops, solver):
    iters = 120 if solver == "richardson" else ITERS
    want, got, x = run_both(ctx, ops, solver, iters=iters)
    assert (got.converged, got.iterations) == (want.converged, want.iterations)
    assert got.n_apply == want.n_apply
    assert len(got.trace) == len(want.trace) and np.array_equal(got.trace, want.trace)
    assert np.array_equal(got.hist, want.hist)
    assert got.abs_err == want.abs_err and got.rel_err == want.rel_err
    assert np.array_equal(x, want.x)


@pytest.mark.parametrize("solver,num_inner", [("gmres", 7), ("fgmres", 30), ("bicgstabl", 4), ("idrs", 2), ("idrs", 8)])
def test_inner_outer_variants(ctx, ops, solver, num_inner):
    """Restart length / l / s other than the defaults, incl. a stop in the middle of a cycle
    (InnerOuterIterativeSolver::finalize, Solver.hpp:250-257)."""
    want, got, x = run_both(ctx, ops, solver, iters=61, rel_tol=0.0, num_inner=num_inner)
    assert got.iterations == want.iterations == 61
    assert np.array_equal(got.trace, want.trace) and np.array_equal(got.hist, want.hist)
    assert np.array_equal(x, want.x)


@pytest.mark.parametrize("solver", ["cg", "cgs", "tfqmr", "tfqmr1", "gmres", "idrs", "bicgstabl"])
def test_against_reference_sequential_golden(ctx, ops, solver):
    """Reference headers with their own sequential sums (golden file made in the build container):
    only the reduction order differs. Converged solution within 1e-8; CG's whole residual history
    within 1e-10 (the Lanczos-type methods amplify rounding noise, as the reference's own -O2 vs -Ofast
    builds do, SURVEY.md F8: leading iterations only)."""
    g = load_golden("solvers_square_nb.npz")
    cpu, gpu = ops
    x = ctx.zeros(cpu.n)
    got = dropin.solve(solver, gpu, x, ctx.vector(rhs(cpu.n)), num_iterations=ITERS, abs_tol=0.0, rel_tol=RTOL)
    ref_hist, ref_x = g[f"{solver}_hist"], g[f"{solver}_x"]
    k = min(len(ref_hist), len(got.hist))
    rel = np.abs(got.hist[:k] - ref_hist[:k]) / ref_hist[:k]
    if solver == "cg":
        assert got.iterations == int(g[f"{solver}_stats"][1]) and rel.max() < HIST_TOL
    else:
        assert rel[:10].max() < HIST_TOL
    assert got.converged == bool(g[f"{solver}_stats"][0])
    assert np.linalg.norm(x.numpy() - ref_x) <= X_TOL * np.linalg.norm(ref_x)


def test_jfnk_template_on_device_vector_bit_identical(ctx, ops):
    """the reference's JfnkSolver (SolverNewton.hpp:101-173) instantiated on the device
    vector -- Jacobian-free products x + delta*y, delta^-1 (z - w), norms, and the reference's own BiCgStabSolver
    as the inner solve. The operator is linear here, so two Newton steps reach the tolerance; every reduction
    scalar, the outer residual history and the solution match the same header on a host vector bit for bit."""
    want, got, x = run_both(ctx, ops, "jfnk", iters=10, rel_tol=1e-9)
    assert want.converged and want.iterations >= 2
    assert (got.converged, got.iterations, got.n_apply) == (want.converged, want.iterations, want.n_apply)
    assert len(got.trace) == len(want.trace) and np.array_equal(got.trace, want.trace)
    assert np.array_equal(got.hist, want.hist) and np.array_equal(x, want.x)
    cpu, _ = ops
    assert np.linalg.norm(rhs(cpu.n) - cpu.apply(x)) <= 1e-8 * np.linalg.norm(rhs(cpu.n))


@pytest.mark.parametrize("solver", ["fused_cg", "fused_bicgstab"])
def test_fused_solvers_behind_the_reference_solver_interface(ctx, ops, solver):
    """Storm::B200::CgSolver / BiCgStabSolver called through Solver<DeviceVector>::solve give the
    iterates of the reference templates bit for bit (same statements, fused schedule)."""
    cpu, gpu = ops
    b = rhs(cpu.n)
    want = orc.ref_solve(solver.removeprefix("fused_"), cpu, b, num_iterations=ITERS, abs_tol=0.0, rel_tol=RTOL,
                         mode=orc.RED_TREE)
    x = ctx.zeros(cpu.n)
    got = dropin.solve(solver, gpu, x, ctx.vector(b), num_iterations=ITERS, abs_tol=0.0, rel_tol=RTOL)
    assert (got.converged, got.iterations) == (want.converged, want.iterations)
    assert np.array_equal(got.hist, want.hist)
    assert np.array_equal(x.numpy(), want.x)


def test_idrs_second_solve_sees_the_advanced_random_stream(ctx, ops):
    """fill_randomly's engine is static in the reference: without a reset, a second
    IDR(s) solve draws different shadow vectors. The drop-in keeps that behaviour."""
    cpu, gpu = ops
    b = ctx.vector(rhs(cpu.n))
    x1, x2, x3 = ctx.zeros(cpu.n), ctx.zeros(cpu.n), ctx.zeros(cpu.n)
    r1 = dropin.solve("idrs", gpu, x1, b, num_iterations=40, abs_tol=0.0, rel_tol=0.0, reset_rng=True)
    r2 = dropin.solve("idrs", gpu, x2, b, num_iterations=40, abs_tol=0.0, rel_tol=0.0, reset_rng=False)
    r3 = dropin.solve("idrs", gpu, x3, b, num_iterations=40, abs_tol=0.0, rel_tol=0.0, reset_rng=True)
    assert np.array_equal(r1.hist, r3.hist) and not np.array_equal(r1.hist, r2.hist)


def test_early_exit_and_misuse(ctx, ops):
    cpu, gpu = ops
    b = rhs(cpu.n)
    for solver in ("cg", "bicgstab", "tfqmr", "idrs"):
        want, got, x = run_both(ctx, ops, solver, iters=10, abs_tol=1e9)   # Solver.hpp:124-128
        assert (got.converged, got.iterations, want.converged, want.iterations) == (True, 0, True, 0)
        assert np.array_equal(x, want.x)
    assert dropin.load().dropin_selftest_errors(ctx.handle) == 3
    with pytest.raises(sb.StormB200Error):
        dropin.solve("no_such_solver", gpu, ctx.zeros(cpu.n), ctx.vector(b))


@pytest.mark.parametrize("grouping", [0, 2])
@pytest.mark.parametrize("side", ["left", "right"])
def test_chebyshev_preconditioner_on_the_device_bit_identical(ctx, square_nb, side, grouping):
    """Storm::ChebyshevPreconditioner in the reference's pre_op slot (Preconditioner.hpp:63-77), on the device vector,
    against THE SAME template compiled on the reference's host vector inside oracle/_ref with the reference's solver
    headers around it: iteration count, every reduction value (build()'s power iterations included), residual history,
    number of operator applies and the solution, bit for bit -- as written and with statement grouping (the default)."""
    cpu = orc.FaceOp(square_nb, prefill=0, dt=-1.0, dirichlet=True)
    b = rhs(cpu.n)
    dropin.set_statement_grouping(grouping)
    try:
        for form in (sb.FORM_FAITHFUL, sb.FORM_COEF):
            gpu = sb.FvmOperator(ctx, square_nb, prefill=0, dt=-1.0, form=form, dirichlet=True)
            oracle_op = cpu if form == sb.FORM_FAITHFUL else orc.RowsOp(cpu.n, *cpu.rows_coef())
            for solver in ("cg", "bicgstab", "fgmres"):
                kw = dict(num_iterations=400, abs_tol=0.0, rel_tol=1e-9, pre_side=side, cheb_degree=5, cheb_eig_ratio=20.0,
                          cheb_power_iterations=8)
                want = orc.ref_solve(solver, oracle_op, b, pre="chebyshev", mode=orc.RED_TREE, **kw)
                x = ctx.zeros(cpu.n)
                got = dropin.solve(solver, gpu, x, ctx.vector(b), precond="chebyshev", **kw)
                assert want.converged and (got.converged, got.iterations, got.n_apply) == (True, want.iterations, want.n_apply)
                assert np.array_equal(got.trace, want.trace) and np.array_equal(got.hist, want.hist)
                assert np.array_equal(x.numpy(), want.x), (solver, form)
    finally:
        dropin.set_statement_grouping(2)   # the default
