"""The contract figures of SURVEY.md 8d that the measurement scripts divide by: pinned here so a script edit cannot
silently change a roofline denominator."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "scripts", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_contract_bytes_per_iteration_on_tetrahedra():
    sweep = load("solver_sweep")
    N = 1_000_000
    F = 2 * N                                   # tets: F = 2 N interior faces, 2 entries per face
    B, V = 24.0 * N + 12.0 * (2 * F), 8.0 * N   # 72 B/cell/apply
    assert B / N == 72.0
    per_cell = lambda s, steps=1: sweep.contract_bytes(s, B, V, steps, 50) / N / steps  # noqa: E731
    assert per_cell("cg") == per_cell("fused_cg") == 144.0            # B + 9V
    assert per_cell("bicgstab") == per_cell("fused_bicgstab") == 264.0   # 2B + 15V
    assert per_cell("cgs") == 2 * 72 + 24 * 8
    assert per_cell("tfqmr") == 2 * 72 + 40 * 8 and per_cell("tfqmr1") == 2 * 72 + 34 * 8
    assert per_cell("bicgstabl") == (4 * 72 + 63 * 8) / 2
    assert per_cell("idrs") == (5 * 72 + 173 * 8) / 4
    assert per_cell("richardson") == 72 + 7 * 8
    # GMRES inner step k: B + (4k + 6) V; a full cycle of m = 50 steps
    ks = np.arange(50)
    assert sweep.contract_bytes("gmres", B, V, 50, 50) == float(np.sum(B + (4 * ks + 6) * V))
    assert sweep.contract_bytes("fused_gmres", B, V, 100, 50) == 2 * sweep.contract_bytes("fgmres", B, V, 50, 50)


def test_sweep_axis_choice_hits_the_requested_sizes():
    sweep = load("apply_sweep")
    for kind, per in (("tet", 6), ("hex", 1), ("poly", 2), ("hexlat", 1)):
        for size in (1e5, 1e6, 1e7, 1e8):
            n = sweep.axis_for(kind, size)
            assert 0.9 * size < per * n ** 3 < 1.1 * size, (kind, size, n)
