"""The statement stream of the reference's solvers (their own templates on Storm::DeviceVector, linked against a
logging stand-in of the C ABI: oracle/statement_trace.py) pins the pass counts SURVEY.md 8a/8d quote and the
measurement scripts divide by -- derived from what the solvers actually execute, not from reading their source."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def counts(tmp_path_factory):
    sys.path.insert(0, ROOT)
    from oracle import statement_trace as st
    if not st.available():
        pytest.skip("statement tracer not built (make -C oracle trace; needs the StormRuler sources)")
    out = tmp_path_factory.mktemp("trace") / "counts.json"
    # in a process of its own: the stand-in library exports the same symbols as libstormb200.so
    subprocess.run([sys.executable, "-m", "oracle.statement_trace", "--json", str(out)], check=True, cwd=ROOT,
                   capture_output=True)
    return json.load(open(out))


def test_as_written_pass_counts_match_the_survey(counts):
    want = {            # solver: (applies, reductions, vector passes as written) per iteration, SURVEY.md 8a
        "cg": (1, 2, 12), "bicgstab": (2, 5, 24), "cgs": (2, 3, 24), "tfqmr": (2, 4, 40),
        "bicgstabl": (4 / 2, None, 63 / 2), "idrs": (5 / 4, None, 173 / 4), "richardson": (1, 1, 7),
    }
    for solver, (applies, reds, passes) in want.items():
        c = counts[solver]
        assert c["applies"] == applies and c["passes_written"] == passes, (solver, c)
        if reds is not None:
            assert c["reductions"] == reds, (solver, c)
    assert counts["tfqmr1"]["passes_written"] <= 34 and counts["tfqmr1"]["applies"] == 2     # "<= 34V": x <- d is conditional
    # GMRES(50): inner step k costs (5k + 8) V as written, 1 apply, k + 2 reductions; + restart and solution update
    g = counts["gmres"]
    ks = range(50)
    assert abs(g["passes_written"] - (sum(5 * k + 8 for k in ks) + 50 * 1.0 + 7) / 50) < 2.0
    assert abs(g["reductions"] - (sum(k + 2 for k in ks) + 1) / 50) < 0.1
    assert counts["fgmres"] == g                                           # no preconditioner: the same stream


def test_fused_schedules_written_down(counts):
    """What a statement-fusing backend moves (element-wise statements + trailing reductions = one pass): equals the
    hand-fused schedules where they exist (CG 9V; GMRES (4k+6)V on average) and is within 2V of the BiCGStab one
    (15V: it also defers `x += alpha p` across an apply, which statement order alone does not allow)."""
    assert counts["cg"]["passes_fused"] == 9
    assert counts["bicgstab"]["passes_fused"] == 17
    g = counts["gmres"]["passes_fused"]
    assert abs(g - sum(4 * k + 6 for k in range(50)) / 50) < 2.0
    for solver in ("cgs", "bicgstabl", "tfqmr", "tfqmr1", "idrs", "richardson"):
        c = counts[solver]
        assert c["passes_fused"] < c["passes_written"] and c["launches_fused"] < c["launches_written"]
    # the numbers quoted in DESIGN.md
    assert (counts["cgs"]["passes_fused"], counts["tfqmr"]["passes_fused"], counts["bicgstabl"]["passes_fused"],
            counts["idrs"]["passes_fused"]) == (17, 35, 21.5, 29.25)


def test_grouped_solvers_move_fewer_bytes_in_fewer_launches(counts):
    """Storm::B200::IdrsSolver / BiCgStabLSolver (statement groups, host scalars): the numbers quoted in DESIGN.md."""
    gi, gb = counts["grouped_idrs"], counts["grouped_bicgstabl"]
    assert (gi["applies"], gi["reductions"]) == (counts["idrs"]["applies"], counts["idrs"]["reductions"])
    assert (gb["applies"], gb["reductions"]) == (counts["bicgstabl"]["applies"], counts["bicgstabl"]["reductions"])
    assert (gi["passes_written"], gi["launches_written"]) == (31.0, 6.5)          # 43.25 V, 18.25 launches as written
    assert (gb["passes_written"], gb["launches_written"]) == (25.5, 8.5)          # 31.5 V, 15 launches as written


def test_automatic_grouping_pass_and_launch_counts(tmp_path):
    """What the generic path moves with Storm::B200::set_statement_grouping(true): the numbers quoted in DESIGN.md."""
    sys.path.insert(0, ROOT)
    from oracle import statement_trace as st
    if not st.available():
        pytest.skip("statement tracer not built (make -C oracle trace; needs the StormRuler sources)")
    out = tmp_path / "grouped.json"
    subprocess.run([sys.executable, "-m", "oracle.statement_trace", "--grouping", "--json", str(out)], check=True, cwd=ROOT,
                   capture_output=True)
    g = json.load(open(out))
    want = {"cg": (9, 3), "cgs": (20, 7), "bicgstab": (18, 6), "bicgstabl": (25.5, 8.5), "tfqmr": (37, 12),
            "tfqmr1": (27, 8), "idrs": (31.75, 7.5), "richardson": (6, 3),
            # the grouped classes with their leading dots riding on the deferred apply
            "grouped_idrs": (29.75, 5.5), "grouped_bicgstabl": (25.0, 8.0)}
    for solver, (passes, launches) in want.items():
        assert (g[solver]["passes_written"], g[solver]["launches_written"]) == (passes, launches), (solver, g[solver])
    assert abs(g["gmres"]["passes_written"] - 105.1) < 0.1 and abs(g["gmres"]["launches_written"] - 27.6) < 0.1
    # CG from the reference's unmodified template now runs the hand-fused schedule: B + 9V in 3 launches
    assert (g["cg"]["applies"], g["cg"]["reductions"]) == (1, 2)


def test_dependency_aware_scheduling_pass_counts(tmp_path):
    """set_statement_grouping(true, reorder = true): consumers launch only the queued statements they depend on. The
    reference's unmodified templates then move fewer bytes than the hand-written schedules in two cases (CG 8 V vs 9 V,
    IDR(4) 26 V vs 29.75 V); the numbers quoted in DESIGN.md."""
    sys.path.insert(0, ROOT)
    from oracle import statement_trace as st
    if not st.available():
        pytest.skip("statement tracer not built (make -C oracle trace; needs the StormRuler sources)")
    out = tmp_path / "scheduled.json"
    subprocess.run([sys.executable, "-m", "oracle.statement_trace", "--grouping", "2", "--json", str(out)], check=True,
                   cwd=ROOT, capture_output=True)
    g = json.load(open(out))
    want = {"cg": (8, 3), "cgs": (20, 8), "bicgstab": (16, 6), "bicgstabl": (23.5, 8.5), "tfqmr": (35, 12),
            "tfqmr1": (25, 8), "idrs": (26.0, 7.5), "richardson": (6, 3)}
    for solver, (passes, launches) in want.items():
        assert (g[solver]["passes_written"], g[solver]["launches_written"]) == (passes, launches), (solver, g[solver])


def test_solver_sweep_contract_uses_the_traced_counts(counts):
    import importlib.util
    spec = importlib.util.spec_from_file_location("solver_sweep", os.path.join(ROOT, "scripts", "solver_sweep.py"))
    sweep = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sweep)
    B, V = 72.0, 8.0
    for solver in ("cgs", "tfqmr", "bicgstabl", "idrs", "richardson"):
        c = counts[solver]
        assert sweep.contract_bytes(solver, B, V, 1, 50) == c["applies"] * B + c["passes_written"] * V, solver
