"""CPU tests of the C++23 host layer (Storm::DeviceVector, DevExpr, traced B200::map, B200::div_grad, the drop-in TU):
the product's drop-in translation unit, linked against a host-executing stand-in of the C ABI (oracle/emu), must issue
exactly the reference's arithmetic. With sequential reductions its results equal the reference's own run bit for bit:

  * all ten solver templates + JFNK on Storm::DeviceVector  ==  the same headers on a host vector (oracle/_ref);
  * the playground's Cahn-Hilliard time step (Playground.cpp:133-175) on the device types  ==  the committed golden
    fixture, produced by the reference's own mesh reader / CellField / map / CgSolver (tests/golden/make_golden.py).

The GPU tests (tests/test_gpu_playground.py, test_gpu_dropin.py) show that libstormb200.so honours the same contract.
"""
import numpy as np
import pytest

from conftest import load_golden, rhs
from oracle import emu, orc

pytestmark = pytest.mark.skipif(not (emu.available() and orc.have_ref()),
                                reason="oracle/_ref (emulated drop-in + reference build) needs the StormRuler sources at build time")

DT, ITERS, RTOL = 0.05, 300, 1e-10


def same(a, b):
    return (a.converged == b.converged and a.iterations == b.iterations and np.array_equal(a.x, b.x)
            and np.array_equal(a.hist, b.hist) and np.array_equal(a.trace, b.trace) and a.n_apply == b.n_apply)


@pytest.mark.parametrize("solver", orc.REF_SOLVERS + orc.REF_NONLINEAR)
def test_reference_templates_on_device_vector_issue_the_reference_arithmetic(square_nb, solver):
    op = orc.FaceOp(square_nb, prefill=1, dt=-DT)
    b = rhs(square_nb.n_cells)
    kw = dict(num_iterations=60 if solver == "richardson" else ITERS, abs_tol=0.0, rel_tol=RTOL)
    want = orc.ref_solve(solver, op, b, **kw)
    got = emu.solve(solver, emu.EmuOp(op), b, **kw)
    assert got.iterations > 10 and same(got, want)


@pytest.mark.parametrize("mode", [orc.RED_SEQ, orc.RED_TREE])
def test_nonsymmetric_operator_and_both_reduction_modes(square_nb, mode):
    rng = np.random.default_rng(5)
    un, bun = rng.standard_normal(square_nb.n_faces), rng.standard_normal(square_nb.n_bfaces)
    op = orc.ConvDiffOp(square_nb, 0.02, un, bun)
    b = rhs(square_nb.n_cells)
    for solver in ("bicgstab", "gmres", "idrs", "tfqmr"):
        kw = dict(num_iterations=80, abs_tol=0.0, rel_tol=1e-12, mode=mode, num_inner=20 if solver == "gmres" else 0)
        assert same(emu.solve(solver, emu.EmuOp(op), b, **kw), orc.ref_solve(solver, op, b, **kw)), solver


@pytest.mark.parametrize("side", ["left", "right"])
def test_preconditioner_slot(square_nb, side):
    """Storm::JacobiPreconditioner in the reference's pre_op slot == the same headers with x / diag as a callback."""
    op = orc.FaceOp(square_nb, prefill=0, dt=-1.0, dirichlet=True)
    _, _, _, _, diag = op.rows_coef()
    diag = diag[:square_nb.n_cells].copy()
    b = rhs(square_nb.n_cells)
    for solver in ("cg", "bicgstab", "fgmres"):
        kw = dict(num_iterations=120, abs_tol=0.0, rel_tol=1e-10, pre_side=side)
        want = orc.ref_solve(solver, op, b, pre=orc.JacobiOp(diag), **kw)
        got = emu.solve(solver, emu.EmuOp(op, diag=diag), b, precond="jacobi", **kw)
        assert same(got, want), solver
        assert emu.counts()["jacobi"] > 0


@pytest.mark.parametrize("name,ref,inner", [("grouped_idrs", "idrs", 0), ("grouped_idrs", "idrs", 1), ("grouped_idrs", "idrs", 6),
                                            ("grouped_idrs", "idrs", 11), ("grouped_bicgstabl", "bicgstabl", 0),
                                            ("grouped_bicgstabl", "bicgstabl", 1), ("grouped_bicgstabl", "bicgstabl", 4)])
def test_grouped_solvers_issue_the_reference_arithmetic(square_nb, name, ref, inner):
    """Storm::B200::IdrsSolver / BiCgStabLSolver (statements issued as sb_eval_group launches) against the reference
    templates: iterate, residual history, the whole sequence of reduction values and the number of operator applies,
    bit for bit; symmetric and non-symmetric operator, sequential and GPU-tree reductions; s = 1 / 4 (default) / 6 / 11
    (11: chains longer than the ABI's 8 terms and batches of more than 8 dots are split), l = 1 / 2 (default) / 4; and
    far fewer launches than statements."""
    rng = np.random.default_rng(5)
    b = rhs(square_nb.n_cells)
    ops = (orc.FaceOp(square_nb, prefill=1, dt=-DT),
           orc.ConvDiffOp(square_nb, 0.02, rng.standard_normal(square_nb.n_faces), rng.standard_normal(square_nb.n_bfaces)))
    for op in ops:
        for mode in (orc.RED_SEQ, orc.RED_TREE):
            kw = dict(num_iterations=150, abs_tol=0.0, rel_tol=RTOL, num_inner=inner, mode=mode)
            want = orc.ref_solve(ref, op, b, **kw)
            got = emu.solve(name, emu.EmuOp(op), b, **kw)
            assert got.iterations > 20 and same(got, want), (name, inner, mode)
            groups = emu.group_count()
            plain = emu.solve(ref, emu.EmuOp(op), b, **kw)
            statements = emu.counts()["eval"] + emu.counts()["dot"] + emu.counts()["norm"]
            assert same(plain, want) and 0 < groups < 0.62 * statements


def test_grouped_solvers_stop_mid_cycle_and_reject_a_preconditioner(square_nb):
    op = orc.FaceOp(square_nb, prefill=0, dt=-1.0, dirichlet=True)
    b = rhs(square_nb.n_cells)
    for name, ref in (("grouped_idrs", "idrs"), ("grouped_bicgstabl", "bicgstabl")):
        for iters in (1, 2, 3, 5, 7):          # num_iterations cuts the inner cycle at every position
            kw = dict(num_iterations=iters, abs_tol=0.0, rel_tol=0.0)
            assert same(emu.solve(name, emu.EmuOp(op), b, **kw), orc.ref_solve(ref, op, b, **kw)), (name, iters)
        kw = dict(num_iterations=400, abs_tol=1e-7, rel_tol=0.0)          # stop on the absolute tolerance
        got, want = emu.solve(name, emu.EmuOp(op), b, **kw), orc.ref_solve(ref, op, b, **kw)
        assert got.converged and same(got, want)
        _, _, _, _, diag = op.rows_coef()
        with pytest.raises(RuntimeError, match="preconditioner"):
            emu.solve(name, emu.EmuOp(op, diag=diag[:square_nb.n_cells].copy()), b, precond="jacobi", num_iterations=5)


@pytest.fixture(params=[1, 2], ids=["grouping", "grouping+scheduling"])
def grouping(request):
    """1: chain-shaped statements queued, everything launched at the next consumer; 2: consumers launch only what they
    depend on (statements may wait across applies and reductions they share no vector with)."""
    emu.set_statement_grouping(request.param)
    yield request.param
    emu.set_statement_grouping(False)


@pytest.mark.parametrize("solver", orc.REF_SOLVERS + orc.REF_NONLINEAR + ("grouped_idrs", "grouped_bicgstabl"))
def test_automatic_statement_grouping_changes_no_bit(square_nb, grouping, solver):
    """Storm::B200::set_statement_grouping(true): the reference templates' chain-shaped statements are queued and launched
    as one sb_eval_group with the reduction that follows. Same iterate, history, reduction values and apply count as the
    reference, in both reduction modes -- with fewer launches than statements."""
    ref = {"grouped_idrs": "idrs", "grouped_bicgstabl": "bicgstabl"}.get(solver, solver)
    op = orc.FaceOp(square_nb, prefill=1, dt=-DT)
    b = rhs(square_nb.n_cells)
    for mode in (orc.RED_SEQ, orc.RED_TREE):
        kw = dict(num_iterations=50 if solver == "richardson" else 120, abs_tol=0.0, rel_tol=RTOL, mode=mode)
        got = emu.solve(solver, emu.EmuOp(op), b, **kw)
        c = emu.counts()
        assert same(got, orc.ref_solve(ref, op, b, **kw)), (solver, mode)
        assert 0 < emu.group_count() < c["eval"] + c["dot"] + c["norm"]
        if solver in ("cg", "bicgstab", "cgs", "idrs", "gmres"):     # `lin_op.mul(z, p); dot_product(p, z)`: one sb_apply_dot
            assert emu.apply_dot_count() >= got.iterations - 1


def test_automatic_grouping_with_preconditioner_and_cahn_hilliard(square_nb, grouping):
    op = orc.FaceOp(square_nb, prefill=0, dt=-1.0, dirichlet=True)
    _, _, _, _, diag = op.rows_coef()
    diag = diag[:square_nb.n_cells].copy()
    b = rhs(square_nb.n_cells)
    for solver in ("cg", "bicgstab", "tfqmr", "idrs"):
        kw = dict(num_iterations=100, abs_tol=0.0, rel_tol=1e-10, pre_side="right")
        want = orc.ref_solve(solver, op, b, pre=orc.JacobiOp(diag), **kw)
        assert same(emu.solve(solver, emu.EmuOp(op, diag=diag), b, precond="jacobi", **kw), want), solver
    g = load_golden("cahn_hilliard_square_nb.npz")
    faces = emu.EmuOp(orc.FaceOp(square_nb.without_boundary(), prefill=0, dt=0.0))
    res, _ = emu.cahn_hilliard_step(faces, g["c0"])
    assert np.array_equal(res.x, g["step0_c"]) and np.array_equal(res.hist, g["step0_hist"])
    assert emu.group_count() == 3 * 2000 + 3 + (grouping == 2)   # 10006 statements + 4001 dots (+ 4002 applies)
    assert emu.selftest_errors() == 3
    # an affine operator through solve_non_uniform, a non-symmetric one, and stops at every position of an inner cycle
    shift = np.cos(0.11 * np.arange(square_nb.n_cells))
    hop = orc.FaceOp(square_nb, prefill=1, dt=-DT)
    for solver in ("cg", "bicgstab", "idrs"):
        orc.ref().ref_reset_rng()
        emu._load()[1].dropin_reset_rng()
        kw = dict(num_iterations=300, abs_tol=0.0, rel_tol=1e-10)
        want, got = orc.ref_solve_non_uniform(solver, hop, b, shift, **kw), emu.solve_non_uniform(solver, emu.EmuOp(hop), b, shift, **kw)
        assert (got.iterations, got.n_apply) == (want.iterations, want.n_apply) and np.array_equal(got.trace, want.trace)
        assert np.array_equal(got.x, want.x)
    rng = np.random.default_rng(5)
    cd = orc.ConvDiffOp(square_nb, 0.02, rng.standard_normal(square_nb.n_faces), rng.standard_normal(square_nb.n_bfaces))
    for solver in ("bicgstab", "bicgstabl", "tfqmr", "idrs", "gmres", "cgs"):
        for iters in (1, 2, 3, 5, 60):
            kw = dict(num_iterations=iters, abs_tol=0.0, rel_tol=0.0, num_inner=7 if solver == "gmres" else 0)
            assert same(emu.solve(solver, emu.EmuOp(cd), b, **kw), orc.ref_solve(solver, cd, b, **kw)), (solver, iters)


@pytest.mark.parametrize("solver", ["cg", "bicgstab", "gmres", "idrs"])
def test_solve_non_uniform_on_device_vector(square_nb, solver):
    """solve_non_uniform (Solver.hpp:271-292: A(x) = b with A(0) != 0) instantiated on Storm::DeviceVector for the affine
    operator A(x) = L x + shift against the same function template on a host vector -- and the solution really solves
    the affine equation."""
    op = orc.FaceOp(square_nb, prefill=1, dt=-DT)
    b = rhs(square_nb.n_cells)
    shift = np.cos(0.11 * np.arange(square_nb.n_cells))
    for mode in (orc.RED_SEQ, orc.RED_TREE):
        kw = dict(num_iterations=1500, abs_tol=0.0, rel_tol=1e-10, mode=mode)
        orc.ref().ref_reset_rng()                 # IDR(s) draws its shadow vectors from fill_randomly
        emu._load()[1].dropin_reset_rng()
        want = orc.ref_solve_non_uniform(solver, op, b, shift, **kw)
        got = emu.solve_non_uniform(solver, emu.EmuOp(op), b, shift, **kw)
        assert (got.converged, got.iterations, got.n_apply) == (want.converged, want.iterations, want.n_apply)
        assert got.converged and np.array_equal(got.trace, want.trace) and np.array_equal(got.x, want.x)
        assert np.linalg.norm(op.apply(got.x) + shift - b) <= 1e-8 * np.linalg.norm(b - shift)


@pytest.mark.parametrize("name,ref", [("fused_cg", "cg"), ("fused_bicgstab", "bicgstab")])
def test_fused_solver_classes_map_options_and_reports(square_nb, name, ref):
    """Storm::B200::CgSolver / BiCgStabSolver (Storm/B200/FusedSolvers.hpp) called through the reference's abstract
    Solver interface: option mapping, public progress fields, residual history and reduction trace -- with the emulator
    answering sb_cg_solve / sb_bicgstab_solve from the oracle's restatements of the same contract. (The CUDA schedules
    behind those entry points are checked on the device, tests/test_gpu_parity.py and test_gpu_dropin.py.)"""
    op = orc.FaceOp(square_nb, prefill=1, dt=-DT)
    b = rhs(square_nb.n_cells)
    for mode in (orc.RED_SEQ, orc.RED_TREE):
        for iters, rtol in ((ITERS, RTOL), (7, 0.0)):
            kw = dict(num_iterations=iters, abs_tol=0.0, rel_tol=rtol, mode=mode)
            want = orc.ref_solve(ref, op, b, **kw)
            got = emu.solve(name, emu.EmuOp(op), b, **kw)
            assert (got.converged, got.iterations) == (want.converged, want.iterations)
            assert (got.abs_err, got.rel_err) == (want.abs_err, want.rel_err)
            assert np.array_equal(got.x, want.x) and np.array_equal(got.hist, want.hist)
            assert np.array_equal(got.trace, want.trace)


def test_reference_identity_preconditioner_on_device_vector(square_nb, capfd):
    """IdentityPreconditioner<DeviceVector> (Preconditioner.hpp:84-97, the one preconditioner the reference ships) in the
    pre_op slot, left and right: the same run as the same headers with y = x as a callback."""
    op = orc.FaceOp(square_nb, prefill=1, dt=-DT)
    b = rhs(square_nb.n_cells)
    ident = orc.CallbackOp(lambda x: x.copy(), square_nb.n_cells)
    for solver in ("cg", "bicgstab", "gmres", "tfqmr", "idrs", "cgs"):
        for side in ("left", "right"):
            kw = dict(num_iterations=40, abs_tol=0.0, rel_tol=1e-10, pre_side=side)
            want = orc.ref_solve(solver, op, b, pre=ident, **kw)
            got = emu.solve(solver, emu.EmuOp(op), b, precond="identity", **kw)
            assert same(got, want), (solver, side)
    capfd.readouterr()   # the reference's class logs every call to std::clog


def test_host_layer_rejects_misuse():
    assert emu.selftest_errors() == 3


def test_playground_cahn_hilliard_step_equals_the_reference_run(square_nb):
    """Two time steps of the playground's Cahn-Hilliard solver: 2 x 2000 CG iterations on the affine 4th-order
    operator (the reference's CG does not converge on it within its 2000-iteration cap -- that IS the reference's
    behaviour, and the iterate is reproduced bit for bit)."""
    g = load_golden("cahn_hilliard_square_nb.npz")
    faces = emu.EmuOp(orc.FaceOp(square_nb.without_boundary(), prefill=0, dt=0.0))
    c = g["c0"]
    for k in range(int(g["num_steps"])):
        res, w_hat = emu.cahn_hilliard_step(faces, c)
        conv, its, abs_err, rel_err = g[f"step{k}_stats"]
        assert res.converged == bool(conv) and res.iterations == int(its) == 2000
        assert res.abs_err == abs_err and res.rel_err == rel_err
        assert np.array_equal(res.hist, g[f"step{k}_hist"])
        assert np.array_equal(res.x, g[f"step{k}_c"])
        # statement stream of one step: per operator evaluation 2 element-wise statements + 2 stormDivGrad; per CG
        # iteration 3 updates + 2 dots; init: f, c_hat, residual, <r,r>, p
        n_op = res.n_apply
        assert n_op == its + 1
        assert emu.counts() == dict(eval=2 * n_op + 3 * its + 4, fill=0, copy=0, dot=2 * its + 1, norm=0, apply=0,
                                    accumulate=2 * n_op, jacobi=0)
        c = res.x


@pytest.mark.parametrize("level", [0, 1, 2])
def test_cahn_hilliard_step_through_solve_non_uniform_equals_the_reference_run(square_nb, level):
    """The playground hands an AFFINE operator to plain solve<CgSolver>, which is why its CG never converges; the
    reference's own solve_non_uniform (Solver.hpp:271-292) is made for exactly that. The same step through it
    (dropin_ch_params::uniformed), on the device types, against the reference's run of it (its mesh, CellField, map,
    CgSolver, solve_non_uniform: golden fixture): three time steps of 49 / 53 / 54 CG iterations instead of 3 x 2000, bit
    for bit, with statement grouping off, on, and on with scheduling."""
    g = load_golden("cahn_hilliard_uniformed_square_nb.npz")
    faces = emu.EmuOp(orc.FaceOp(square_nb.without_boundary(), prefill=0, dt=0.0))
    emu.set_statement_grouping(level)
    try:
        c = g["c0"]
        for k in range(int(g["num_steps"])):
            res, _ = emu.cahn_hilliard_step(faces, c, uniformed=True)
            conv, its, abs_err, rel_err = g[f"step{k}_stats"]
            assert (res.converged, res.iterations) == (bool(conv), int(its)) and int(its) < 60
            assert res.abs_err == abs_err and res.rel_err == rel_err
            assert np.array_equal(res.x, g[f"step{k}_c"])
            c = res.x
    finally:
        emu.set_statement_grouping(0)
    # and the oracle restatement used by the GPU test
    res = orc.cahn_hilliard_step(square_nb, g["c0"], uniformed=True)
    assert res.iterations == 49 and np.array_equal(res.x, g["step0_c"])


def test_oracle_restatement_of_the_cahn_hilliard_step_is_pinned(square_nb):
    """oracle.orc.cahn_hilliard_step (numpy statements + the C face loop, the checker of the GPU test) against the
    same golden fixture, and the traced dF/dc against its numpy restatement."""
    g = load_golden("cahn_hilliard_square_nb.npz")
    res = orc.cahn_hilliard_step(square_nb, g["c0"])
    assert np.array_equal(res.x, g["step0_c"]) and np.array_equal(res.hist, g["step0_hist"])
    op = orc.CahnHilliardOp(square_nb, g["c0"])
    c_restated = orc.solve("cg", op, op.c, x0=op.c)   # plain-C CG restatement instead of the reference template
    assert np.array_equal(c_restated.x, g["step0_c"])
    # reduction order: the GPU tree changes the last bits of every dot; over 2000 non-converging CG iterations the
    # iterate stays within 1e-11 of the sequential run, the residual history within 1e-10 for the first 80 iterations (1.3e-8 at worst over all 2000)
    tree = orc.cahn_hilliard_step(square_nb, g["c0"], mode=orc.RED_TREE)
    assert np.linalg.norm(tree.x - g["step0_c"]) <= 1e-10 * np.linalg.norm(g["step0_c"])
    rel = np.abs(tree.hist - g["step0_hist"]) / g["step0_hist"]
    assert rel[:80].max() < 1e-10 and rel.max() < 1e-7


@pytest.mark.parametrize("n", [1, 2047, 2048, 2049, 100_003, 1_200_000])
def test_group_kernel_body_compiled_for_the_host(n):
    """The product's own device code of sb_eval_group (csrc/sb_group_body.cuh, compiled for the host by
    oracle/emu/group_body_host.cpp and driven thread by thread like ew_kernel + the reduction tree) against numpy,
    statement by statement, and the oracle's restatement of the tree: chains with and without a base, + and - terms,
    targets that are their own base or term, later statements reading earlier targets, 8-term chains, 3 fused dots.
    Same statements as tests/test_gpu_playground.py::test_statement_group_kernel_bit_exact runs on the device."""
    import ctypes as C
    import os
    from stormruler_b200 import capi
    path = os.path.join(os.path.dirname(emu.EMU), "libgroup_body_host.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/emu/libgroup_body_host.so not built")
    lib = C.CDLL(path)
    rng = np.random.default_rng(n)
    cap = -(-n // 2048) * 2048
    names = "abcdefgh"
    host = {k: rng.standard_normal(n) for k in names}
    dev = {}
    for k, v in host.items():
        dev[k] = np.zeros(cap)
        dev[k][:n] = v
    ptr = lambda k: dev[k].ctypes.data_as(C.c_void_p)  # noqa: E731
    stmts = [("a", "b", [(0.3, "c", 1), (-1.7, "d", 0), (2.5, "a", 1)]),
             ("e", None, [(1.25, "a", 0), (0.5, "e", 0), (3.0, "f", 0)]),
             ("g", "g", [(0.75, "e", 1)]),
             ("h", "a", [(c, x, s) for c, x, s in zip(rng.standard_normal(8), "bcdefgab", [0, 1] * 4)])]
    dots = [("a", "e"), ("g", "g"), ("h", "b")]
    S = (capi.Chain * len(stmts))()
    for k, (y, base, terms) in enumerate(stmts):
        S[k].y, S[k].base, S[k].n_terms = ptr(y), (ptr(base) if base else None), len(terms)
        for t, (c, x, sub) in enumerate(terms):
            S[k].c[t], S[k].x[t], S[k].sub[t] = float(c), ptr(x), sub
        a = None if base is None else host[base].copy()
        for t, (c, x, sub) in enumerate(terms):
            p = c * host[x]
            a = p if (a is None and t == 0) else (a - p if sub else a + p)
        host[y] = a
    A = (C.c_void_p * 3)(*[ptr(p) for p, _ in dots])
    B = (C.c_void_p * 3)(*[ptr(q) for _, q in dots])
    out = np.zeros(3)
    lib.group_body_host_run.argtypes = [C.c_size_t, C.c_int, C.POINTER(capi.Chain), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    assert lib.group_body_host_run(n, len(stmts), S, 3, A, B, out.ctypes.data_as(C.c_void_p)) == 0
    for k in names:
        assert np.array_equal(dev[k][:n], host[k]), k
    assert np.array_equal(out, [orc.dot(host[p], host[q], orc.RED_TREE) for p, q in dots])


def _apply_rows_host():
    import ctypes as C
    import os
    path = os.path.join(os.path.dirname(emu.EMU), "libapply_rows_host.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/emu/libapply_rows_host.so not built")
    lib = C.CDLL(path)
    lib.apply_rows_host.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_int, C.c_double, C.c_void_p, C.c_void_p]

    def run(form, w, n, ld, col, v0, v1, diag, prefill, dt, x, y):
        p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)  # noqa: E731
        assert lib.apply_rows_host(form, w, n, ld, p(col), p(v0), p(v1), p(diag), prefill, dt, p(x), p(y)) == 0
    return run


@pytest.mark.parametrize("dirichlet", [False, True])
@pytest.mark.parametrize("mesh_name", ["square_nb", "rectangle", "tets", "poly"])
def test_apply_row_arithmetic_compiled_for_the_host(request, mesh_name, dirichlet):
    """The product's own per-row device code (csrc/sb_apply_rows.cuh, compiled for the host) on the oracle's rows in the
    product's layout contract: the faithful form must reproduce the reference's FACE LOOP bit for bit -- from 0, from x
    (`c_hat <<= c_in` first) and from the old y (sb_apply_accumulate: stormDivGrad as the playground calls it) -- and
    the coefficient form its row oracle; triangles, tetrahedra (Dirichlet ghosts) and 14-face polyhedra (rows chunked by 8)."""
    from stormruler_b200.mesh import CELL_TET, Mesh, PolyMesh
    run = _apply_rows_host()
    if mesh_name == "tets":
        m = Mesh.box(CELL_TET, 6, 5, 4, jitter=0.2)
        m.renumber_rcm()
    elif mesh_name == "poly":
        m = PolyMesh.bcc(4, stretch=(1.0, 1.3, 0.7))
    else:
        m = request.getfixturevalue(mesh_name)
    fm = orc.FaceMesh(m.n_cells, m.face_cell, m.face_area, m.face_dist, m.cell_vol, m.bface_cell, m.bface_area, m.bface_dist)
    n = fm.n_cells
    ld = -(-n // 2048) * 2048
    rng = np.random.default_rng(3)
    x, y0 = np.zeros(ld), np.zeros(ld)
    x[:n], y0[:n] = rng.standard_normal(n), rng.standard_normal(n)
    for prefill, dt in ((0, -1.0), (1, -0.05)):
        op = orc.FaceOp(fm, prefill=prefill, dt=dt, dirichlet=dirichlet)
        w, _, col, g, d = op.rows_faithful(ld)
        if w not in (1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 14, 16):
            pytest.skip(f"row width {w} has no kernel instantiation")
        y = np.full(ld, np.nan)
        run(0, w, n, ld, col, g, d, None, prefill, dt, x, y)
        assert np.array_equal(y[:n], op.apply(x[:n])), ("faithful", prefill)
        # accumulate mode: the operator's own prefill / dt are overridden per call
        for dt2 in (-1.0e-4, 0.37):
            y = y0.copy()
            run(0, w, n, ld, col, g, d, None, 2, dt2, x, y)
            assert np.array_equal(y[:n], op.divgrad_accumulate(dt2, x[:n], y0[:n].copy())), ("accumulate", dt2)
        wc, _, colc, a, diag = op.rows_coef(ld)
        y = np.full(ld, np.nan)
        run(1, max(wc, 1), n, ld, colc, a, None, diag, prefill, dt, x, y)
        assert np.array_equal(y[:n], op.apply_rows_coef(x[:n], (wc, ld, colc, a, diag))), ("coef", prefill)


@pytest.mark.parametrize("block", range(4))
def test_random_programs_agree_across_grouping_modes(square_nb, block):
    """Property test of the statement queue: random programs over a pool of device vectors -- chain-shaped and nested
    statements, in-place updates, reductions, operator applies, fills, scalings, pointer swaps, re-allocations, host
    reads (stormruler_b200/host/dropin.cpp: dropin_random_program) -- must leave every vector and every recorded value
    bit-identical whether statements are launched one by one, queued and grouped, or queued with dependency-aware
    scheduling. (This test found the one scheduling hazard the solver runs do not exercise: a dot riding on a deferred
    apply whose third operand still had a queued update.)"""
    face_op = orc.FaceOp(square_nb, prefill=1, dt=-1.0e-4, dirichlet=True)   # norm ~ 2: values stay finite over 300 steps
    _, _, _, _, diag = face_op.rows_coef()
    op = emu.EmuOp(face_op, diag=diag[:square_nb.n_cells].copy())      # accumulate + Jacobi available on the emulator
    n = square_nb.n_cells
    try:
        for seed in range(25 * block, 25 * (block + 1)):
            init = np.random.default_rng(seed).standard_normal((3 + seed % 6, n))
            res = []
            for level in (0, 1, 2):
                emu.set_statement_grouping(level)
                final, rec = emu.random_program(op, init, seed, 300, mode=seed % 2, with_accumulate=True, with_jacobi=True)
                res.append((final.view(np.uint64).copy(), rec.view(np.uint64).copy()))
            for level in (1, 2):
                assert np.array_equal(res[0][0], res[level][0]) and np.array_equal(res[0][1], res[level][1]), (seed, level)
            assert np.isfinite(res[0][0].view(np.float64)).all() and np.isfinite(res[0][1].view(np.float64)).all()
    finally:
        emu.set_statement_grouping(0)


@pytest.mark.parametrize("grouping_mode", [0, 1, 2])
@pytest.mark.parametrize("side", ["left", "right"])
def test_chebyshev_preconditioner_same_template_on_host_and_device_vector(square_nb, side, grouping_mode):
    """Storm::ChebyshevPreconditioner (host/Storm/B200/ChebyshevPreconditioner.hpp) is one template in the reference's
    vector vocabulary: compiled on the reference's host vector inside oracle/_ref (with the reference's own solver
    headers around it) it is the checker of itself on Storm::DeviceVector. Iterate, residual history, every reduction
    value (the power iterations of build() included) and the number of operator applies, bit for bit -- for the
    preconditioned branches of CG, BiCGStab, FGMRES and IDR(s), left and right, sequential and tree reductions, with
    and without statement grouping; and the preconditioner does what it is for: far fewer iterations."""
    emu.set_statement_grouping(grouping_mode)
    try:
        op = orc.FaceOp(square_nb, prefill=0, dt=-1.0, dirichlet=True)   # the Poisson operator of configs 2 / 4
        b = rhs(square_nb.n_cells)
        for solver in ("cg", "bicgstab", "fgmres", "idrs"):
            for mode in (orc.RED_SEQ, orc.RED_TREE):
                kw = dict(num_iterations=400, abs_tol=0.0, rel_tol=1e-9, pre_side=side, mode=mode, cheb_degree=5,
                          cheb_eig_ratio=20.0, cheb_power_iterations=8)
                want = orc.ref_solve(solver, op, b, pre="chebyshev", **kw)
                got = emu.solve(solver, emu.EmuOp(op), b, precond="chebyshev", **kw)
                assert want.converged and same(got, want), (solver, mode)
                assert got.n_apply == want.n_apply and got.n_apply > 4 * got.iterations
        plain = orc.ref_solve("cg", op, b, num_iterations=2000, abs_tol=0.0, rel_tol=1e-9)
        pre = orc.ref_solve("cg", op, b, num_iterations=2000, abs_tol=0.0, rel_tol=1e-9, pre="chebyshev", cheb_degree=5,
                            cheb_eig_ratio=20.0)
        # 526 iterations without, 127 with (a similar number of operator applies, a quarter of the reductions); same solution
        assert plain.converged and pre.converged and pre.iterations * 3 < plain.iterations
        assert np.linalg.norm(pre.x - plain.x) <= 1e-7 * np.linalg.norm(plain.x)
    finally:
        emu.set_statement_grouping(False)
