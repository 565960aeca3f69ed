"""GPU parity of the playground's own caller of the path: `stormDivGrad(mesh, u, dt, c)` as it
is called there -- accumulating onto a pre-filled field (sb_apply_accumulate) -- the traced `map(dF_dc, c)`, and one
whole Cahn-Hilliard time step (Playground.cpp:133-175) through the C++ drop-in on the device.

Checker: the oracle restatement, pinned on the CPU against the reference's own run of the same step
(tests/test_dropin_emulated.py, tests/golden/cahn_hilliard_square_nb.npz).
"""
import numpy as np
import pytest

import stormruler_b200 as sb
from conftest import load_golden
from oracle import orc
from stormruler_b200 import dropin
from stormruler_b200 import mesh as sbmesh

pytestmark = pytest.mark.gpu


# Order: the tests that only involve kernels already proven on the device come first (pytest -x).
def _random_state(n, seed):
    rng = np.random.default_rng(seed)
    return rng.standard_normal(n), rng.standard_normal(n)


def _chain_host(vals, y, base, terms):
    a = None if base is None else vals[base].copy()
    for t, (c, x, sub) in enumerate(terms):
        p = c * vals[x]
        a = p if (a is None and t == 0) else (a - p if sub else a + p)
    vals[y] = a


def test_traced_map_program_of_dF_dc(ctx):
    """The postfix program B200::map records for the playground's dF_dc (2.0*c*(c-1.0)*(2.0*c-1.0), constants
    deduplicated): S0 V0 MUL V0 S1 SUB MUL S0 V0 MUL S1 SUB MUL -- run through sb_eval's run-time interpreter."""
    rng = np.random.default_rng(8)
    c = rng.random(5000)
    S0, S1, V0 = sb.capi.OP_SCAL0, sb.capi.OP_SCAL0 + 1, sb.capi.OP_VEC0
    MUL, SUB = sb.capi.OP_MUL, sb.capi.OP_SUB
    cv, f = ctx.vector(c), ctx.zeros(c.shape[0])
    ctx.eval(f, sb.ASSIGN, [S0, V0, MUL, V0, S1, SUB, MUL, S0, V0, MUL, S1, SUB, MUL], [cv], [2.0, 1.0])
    assert np.array_equal(f.numpy(), orc.ch_dF_dc(c))


@pytest.mark.parametrize("solver", ["cg", "idrs"])
def test_solve_non_uniform_on_the_device(ctx, square_nb, solver):
    """The reference's solve_non_uniform template (Solver.hpp:271-292) on DeviceVector, affine operator A(x) = L x + shift."""
    from conftest import rhs
    cpu = orc.FaceOp(square_nb, prefill=1, dt=-0.05)
    gpu = sb.FvmOperator(ctx, square_nb, prefill=1, dt=-0.05, form=sb.FORM_FAITHFUL)
    b, shift = rhs(cpu.n), np.cos(0.11 * np.arange(cpu.n))
    orc.ref().ref_reset_rng()
    want = orc.ref_solve_non_uniform(solver, cpu, b, shift, num_iterations=1500, abs_tol=0.0, rel_tol=1e-10, mode=orc.RED_TREE)
    x = ctx.zeros(cpu.n)
    got = dropin.solve_non_uniform(solver, gpu, x, ctx.vector(b), ctx.vector(shift), num_iterations=1500, abs_tol=0.0, rel_tol=1e-10)
    assert got.converged and (got.iterations, got.n_apply) == (want.iterations, want.n_apply)
    assert np.array_equal(got.trace, want.trace) and np.array_equal(x.numpy(), want.x)


def test_apply_with_a_riding_dot(ctx, square_nb):
    """sb_apply_dot: y = A x and <x, y> / <u, y> from one kernel -- the same y as sb_apply, the same value as sb_dot."""
    rng = np.random.default_rng(21)
    n = square_nb.n_cells
    xh, uh = rng.standard_normal(n), rng.standard_normal(n)
    for form in (sb.FORM_COEF, sb.FORM_FAITHFUL):
        op = sb.FvmOperator(ctx, square_nb, prefill=1, dt=-0.05, form=form)
        x, u, y, y2 = ctx.vector(xh), ctx.vector(uh), ctx.zeros(n), ctx.zeros(n)
        op.mul(y, x)
        d_xy = op.mul_dot(y2, x)
        assert np.array_equal(y2.numpy(), y.numpy()) and d_xy == ctx.dot(x, y)
        y2.fill(0.0)
        d_uy = op.mul_dot(y2, x, u)
        assert np.array_equal(y2.numpy(), y.numpy()) and d_uy == ctx.dot(u, y)
        assert d_xy == orc.dot(xh, y.numpy(), orc.RED_TREE)
        with pytest.raises(sb.StormB200Error, match="alias"):
            op.mul_dot(y2, x, y2)
        y2.fill(0.0)
        yy, yx = op.mul_dot_yy_yx(y2, x)
        assert np.array_equal(y2.numpy(), y.numpy()) and yy == ctx.dot(y, y) and yx == ctx.dot(y, x)


@pytest.mark.parametrize("dirichlet", [False, True])
@pytest.mark.parametrize("name", ["square_nb", "rectangle"])
def test_div_grad_accumulates_like_the_face_loop(ctx, name, dirichlet, request):
    m = request.getfixturevalue(name)
    cpu = orc.FaceOp(m if dirichlet else m.without_boundary(), prefill=0, dt=0.0, dirichlet=dirichlet)
    # the operator's own prefill / dt must not matter
    gpu = sb.FvmOperator(ctx, m, prefill=1, dt=123.0, form=sb.FORM_FAITHFUL, dirichlet=dirichlet)
    c, u0 = _random_state(m.n_cells, 3)
    for dt in (-1.0e-4, -1.0e-3, 0.37):
        want = cpu.divgrad_accumulate(dt, c, u0.copy())
        u = ctx.vector(u0)
        gpu.div_grad(u, dt, ctx.vector(c))
        assert np.array_equal(u.numpy(), want), (name, dirichlet, dt)
    # twice in a row accumulates twice (the second call starts from the first one's result)
    u = ctx.vector(u0)
    cv = ctx.vector(c)
    gpu.div_grad(u, -0.5, cv)
    gpu.div_grad(u, -0.5, cv)
    want = cpu.divgrad_accumulate(-0.5, c, cpu.divgrad_accumulate(-0.5, c, u0.copy()))
    assert np.array_equal(u.numpy(), want)


def test_div_grad_on_a_3d_mesh_and_plain_apply_unchanged(ctx):
    """Tetrahedra (rows up to 4 wide, Dirichlet ghosts), sizes off the 2048-row tile; sb_apply on the same operator
    still starts from x (prefill 1) -- the per-call override must not leak into the stored operator."""
    m = sbmesh.Mesh.box(sbmesh.CELL_TET, 9, 8, 7, jitter=0.2)   # 3024 Kuhn tetrahedra, shuffled
    m.renumber_rcm()
    cpu_acc = orc.FaceOp(m, prefill=0, dt=0.0, dirichlet=True)
    cpu_mul = orc.FaceOp(m, prefill=1, dt=-0.05, dirichlet=True)
    gpu = sb.FvmOperator(ctx, m, prefill=1, dt=-0.05, form=sb.FORM_FAITHFUL, dirichlet=True)
    c, u0 = _random_state(m.n_cells, 4)
    cv, u, y = ctx.vector(c), ctx.vector(u0), ctx.zeros(m.n_cells)
    gpu.mul(y, cv)
    assert np.array_equal(y.numpy(), cpu_mul.apply(c))
    gpu.div_grad(u, -2.5e-3, cv)
    assert np.array_equal(u.numpy(), cpu_acc.divgrad_accumulate(-2.5e-3, c, u0.copy()))
    gpu.mul(y, cv)
    assert np.array_equal(y.numpy(), cpu_mul.apply(c))


def test_div_grad_needs_the_faithful_form_and_distinct_vectors(ctx, square_nb):
    coef = sb.FvmOperator(ctx, square_nb, prefill=1, dt=-0.05, form=sb.FORM_COEF)
    u, c = ctx.zeros(square_nb.n_cells), ctx.zeros(square_nb.n_cells)
    with pytest.raises(sb.StormB200Error, match="faithful"):
        coef.div_grad(u, -1.0, c)
    faithful = sb.FvmOperator(ctx, square_nb, prefill=0, dt=0.0, form=sb.FORM_FAITHFUL)
    with pytest.raises(sb.StormB200Error, match="alias"):
        faithful.div_grad(u, -1.0, u)


def test_playground_cahn_hilliard_step_on_the_device(ctx, square_nb):
    """One whole time step (2000 CG iterations of the reference's CgSolver template on DeviceVector, each with two
    element-wise statements and two accumulating stormDivGrad applies): bit-identical to the oracle with the GPU
    reduction tree, and within the stated tolerances of the reference's own sequential run (golden fixture)."""
    if not dropin.available():
        pytest.fail("libstorm_dropin.so is missing on the GPU box (it is built where the reference tree is mounted)")
    g = load_golden("cahn_hilliard_square_nb.npz")
    n = square_nb.n_cells
    faces = sb.FvmOperator(ctx, square_nb, prefill=0, dt=0.0, form=sb.FORM_FAITHFUL)
    c, c_hat, w_hat = ctx.vector(g["c0"]), ctx.zeros(n), ctx.zeros(n)
    got = dropin.cahn_hilliard_step(faces, c, c_hat, w_hat)
    want = orc.cahn_hilliard_step(square_nb, g["c0"], mode=orc.RED_TREE)
    assert got.iterations == want.iterations == 2000 and got.converged == want.converged
    assert np.array_equal(got.trace, want.trace), "reduction trace differs from the oracle (GPU tree)"
    assert np.array_equal(got.hist, want.hist)
    assert np.array_equal(c_hat.numpy(), want.x)
    assert got.n_apply == 2001
    # against the reference's own run (sequential reductions): tolerances of the north star; the history of this
    # non-converging 2000-iteration CG drifts apart slowly with the reduction order (1e-10 for the first 80 iterations, 1.3e-8 at worst)
    gold_c, gold_h = g["step0_c"], g["step0_hist"]
    assert np.linalg.norm(c_hat.numpy() - gold_c) <= 1e-8 * np.linalg.norm(gold_c)
    rel = np.abs(got.hist - gold_h) / gold_h
    assert rel[:80].max() < 1e-10 and rel.max() < 1e-7
    assert np.isfinite(w_hat.numpy()).all()   # the chemical potential of the last operator evaluation


def test_cahn_hilliard_step_through_solve_non_uniform_on_the_device(ctx, square_nb):
    """The same step through the reference's solve_non_uniform (the operator is affine): CG converges in 49 iterations;
    bit-identical to the oracle with the GPU tree, within tolerance of the reference's own sequential run."""
    g = load_golden("cahn_hilliard_uniformed_square_nb.npz")
    n = square_nb.n_cells
    faces = sb.FvmOperator(ctx, square_nb, prefill=0, dt=0.0, form=sb.FORM_FAITHFUL)
    c, c_hat, w_hat = ctx.vector(g["c0"]), ctx.zeros(n), ctx.zeros(n)
    got = dropin.cahn_hilliard_step(faces, c, c_hat, w_hat, uniformed=True)
    want = orc.cahn_hilliard_step(square_nb, g["c0"], mode=orc.RED_TREE, uniformed=True)
    assert got.converged and got.iterations == want.iterations
    assert np.array_equal(got.trace, want.trace) and np.array_equal(c_hat.numpy(), want.x)
    assert abs(got.iterations - int(g["step0_stats"][1])) <= 1
    assert np.linalg.norm(c_hat.numpy() - g["step0_c"]) <= 1e-8 * np.linalg.norm(g["step0_c"])


def test_playground_driver_end_to_end(tmp_path):
    """scripts/playground_cahn_hilliard.py: mesh files -> reader -> upload -> two time steps -> VTK output, as a
    separate process (the application a reference user would run)."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "scripts", "playground_cahn_hilliard.py"), "--generate", "24", "16",
                          "--steps", "2", "--max-iterations", "60", "--out", str(tmp_path)],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["cells"] == 2 * 24 * 16 and [r["cg_iterations"] for r in line["per_step"]] == [60, 60]
    # the scheme conserves the mean of c on this uniform mesh whatever the (here: capped, non-converged) CG does
    import importlib.util
    spec = importlib.util.spec_from_file_location("playground_driver", os.path.join(root, "scripts", "playground_cahn_hilliard.py"))
    drv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(drv)
    assert abs(line["c_mean"] - drv.initial_condition(line["cells"]).mean()) < 1e-9
    assert np.isfinite([line["c_min"], line["c_max"]]).all() and line["c_min"] < line["c_mean"] < line["c_max"]
    assert out.stdout.count("time = ") == 3
    for k in range(3):
        text = (tmp_path / f"fields-{k:05d}.vtk").read_text()
        assert text.startswith("# vtk DataFile Version 2.0\n") and "SCALARS c double 1" in text


@pytest.mark.parametrize("n", [1, 2047, 2048, 2049, 100_003])
def test_statement_group_kernel_bit_exact(ctx, n):
    """sb_eval_group against numpy, statement by statement (every product and sum rounded separately), with the
    aliasing the solvers use: targets that are their own base or term, later statements reading earlier targets, and
    up to 8 dots over the final values (3 ride on the statement kernel, the rest are stand-alone dot kernels)."""
    rng = np.random.default_rng(n)
    names = "abcdefgh"
    host = {k: rng.standard_normal(n) for k in names}
    dev = {k: ctx.vector(v) for k, v in host.items()}
    stmts = [("a", "b", [(0.3, "c", 1), (-1.7, "d", 0), (2.5, "a", 1)]),            # a = ((b - .3c) + (-1.7)d) - 2.5a
             ("e", None, [(1.25, "a", 0), (0.5, "e", 0), (3.0, "f", 0)]),           # e = (1.25a + .5e) + 3f   (reads new a)
             ("g", "g", [(0.75, "e", 1)]),                                           # g -= .75e                (reads new e)
             ("h", "a", [(c, x, s) for c, x, s in zip(rng.standard_normal(8), "bcdefgab", [0, 1] * 4)])]   # 8 terms
    dots = [("a", "e"), ("g", "g"), ("h", "b"), ("a", "a"), ("c", "h"), ("e", "g"), ("b", "b"), ("h", "h")]
    for y, base, terms in stmts:
        _chain_host(host, y, base, terms)
    got = ctx.eval_group([(dev[y], dev[b] if b else None, [(c, dev[x], s) for c, x, s in terms]) for y, b, terms in stmts],
                         [(dev[p], dev[q]) for p, q in dots])
    for k in names:
        assert np.array_equal(dev[k].numpy(), host[k]), k
    want = np.array([orc.dot(host[p], host[q], orc.RED_TREE) for p, q in dots])
    assert np.array_equal(got, want)
    # statements only / dots only
    ctx.eval_group([(dev["c"], dev["c"], [(2.0, dev["d"], 0)])])
    assert np.array_equal(dev["c"].numpy(), host["c"] + 2.0 * host["d"])
    assert ctx.eval_group(dots=[(dev["d"], dev["d"])])[0] == orc.dot(host["d"], host["d"], orc.RED_TREE)
    with pytest.raises(sb.StormB200Error):
        ctx.eval_group()


@pytest.mark.parametrize("name,inner", [("grouped_idrs", 0), ("grouped_idrs", 7), ("grouped_idrs", 11),
                                        ("grouped_bicgstabl", 0), ("grouped_bicgstabl", 3)])
def test_grouped_solvers_bit_identical_to_the_reference_templates(ctx, square_nb, name, inner):
    """Storm::B200::IdrsSolver / BiCgStabLSolver on the device against the reference's own templates on a host vector
    (GPU reduction tree): iteration count, every reduction value, residual history, solution."""
    ref = dropin.GROUPED_SOLVERS[name]
    from conftest import rhs
    rng = np.random.default_rng(12)
    un, bun = rng.standard_normal(square_nb.n_faces), rng.standard_normal(square_nb.n_bfaces)
    # (oracle operator, device operator) pairs whose applies are bit-identical: face loop vs faithful rows
    # (tests/test_gpu_dropin.py), coefficient rows of the convection-diffusion operator (tests/test_gpu_convdiff.py)
    cases = ((orc.FaceOp(square_nb, prefill=1, dt=-0.05), sb.FvmOperator(ctx, square_nb, prefill=1, dt=-0.05, form=sb.FORM_FAITHFUL)),
             (orc.ConvDiffOp(square_nb, 0.02, un, bun).rows_coef(), sb.ConvDiffOperator(ctx, square_nb, 0.02, un, bun)))
    b = rhs(square_nb.n_cells)
    for cpu_op, gpu in cases:
        want = orc.ref_solve(ref, cpu_op, b, num_iterations=200, abs_tol=0.0, rel_tol=1e-10, num_inner=inner, mode=orc.RED_TREE)
        x = ctx.zeros(cpu_op.n)
        got = dropin.solve(name, gpu, x, ctx.vector(b), num_iterations=200, abs_tol=0.0, rel_tol=1e-10, num_inner=inner)
        assert (got.converged, got.iterations, got.n_apply) == (want.converged, want.iterations, want.n_apply)
        assert np.array_equal(got.trace, want.trace) and np.array_equal(got.hist, want.hist)
        assert np.array_equal(x.numpy(), want.x)


@pytest.mark.parametrize("level", [1, 2], ids=["grouping", "grouping+scheduling"])
@pytest.mark.parametrize("solver", ["cg", "bicgstab", "tfqmr", "idrs", "gmres"])
def test_automatic_statement_grouping_on_the_device(ctx, square_nb, solver, level):
    """Storm::B200::set_statement_grouping(true): the reference templates' statements queued and launched as
    sb_eval_group -- same bits as without, i.e. as the reference on a host vector with the GPU reduction tree."""
    from conftest import rhs
    cpu = orc.FaceOp(square_nb, prefill=1, dt=-0.05)
    gpu = sb.FvmOperator(ctx, square_nb, prefill=1, dt=-0.05, form=sb.FORM_FAITHFUL)
    b = rhs(square_nb.n_cells)
    want = orc.ref_solve(solver, cpu, b, num_iterations=150, abs_tol=0.0, rel_tol=1e-10, mode=orc.RED_TREE)
    dropin.set_statement_grouping(level)
    try:
        launches0 = ctx.launch_count
        x = ctx.zeros(cpu.n)
        got = dropin.solve(solver, gpu, x, ctx.vector(b), num_iterations=150, abs_tol=0.0, rel_tol=1e-10)
        grouped_launches = ctx.launch_count - launches0
    finally:
        dropin.set_statement_grouping(False)
    assert (got.converged, got.iterations, got.n_apply) == (want.converged, want.iterations, want.n_apply)
    assert np.array_equal(got.trace, want.trace) and np.array_equal(got.hist, want.hist)
    assert np.array_equal(x.numpy(), want.x)
    launches0 = ctx.launch_count
    dropin.solve(solver, gpu, ctx.zeros(cpu.n), ctx.vector(b), num_iterations=150, abs_tol=0.0, rel_tol=1e-10)
    assert grouped_launches < ctx.launch_count - launches0


@pytest.mark.parametrize("seed", range(6))
def test_random_programs_on_the_device_agree_across_grouping_modes_and_with_the_emulator(ctx, square_nb, seed):
    """The property test of tests/test_dropin_emulated.py on the device: a random program of statements, reductions,
    applies (plain and accumulating), fills, swaps, re-allocations and host reads gives the same bits with grouping
    off, on, and on with dependency-aware scheduling -- and the same bits as the host emulator of the C ABI with the
    GPU reduction tree, i.e. every kernel involved honours the contract the emulator restates."""
    from oracle import emu
    face_cpu = orc.FaceOp(square_nb, prefill=1, dt=-1.0e-4, dirichlet=True)
    gpu = sb.FvmOperator(ctx, square_nb, prefill=1, dt=-1.0e-4, form=sb.FORM_FAITHFUL, dirichlet=True)
    init = np.random.default_rng(seed).standard_normal((3 + seed % 5, square_nb.n_cells))
    res = []
    try:
        for level in (0, 1, 2):
            dropin.set_statement_grouping(level)
            final, rec = dropin.random_program(gpu, init, seed, 250, with_accumulate=True)
            res.append((final.view(np.uint64).copy(), rec.view(np.uint64).copy()))
    finally:
        dropin.set_statement_grouping(0)
    for level in (1, 2):
        assert np.array_equal(res[0][0], res[level][0]) and np.array_equal(res[0][1], res[level][1]), level
    if emu.available():
        want_final, want_rec = emu.random_program(emu.EmuOp(face_cpu), init, seed, 250, mode=orc.RED_TREE, with_accumulate=True)
        assert np.array_equal(res[0][1], want_rec.view(np.uint64)) and np.array_equal(res[0][0], want_final.view(np.uint64))
