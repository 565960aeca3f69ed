"""ctypes access to libstorm_dropin.so: StormRuler's own solver templates instantiated on
Storm::DeviceVector (stormruler_b200/host/dropin.cpp). Built by `make -C stormruler_b200/host` where the
reference tree is available; the built library travels to the GPU box. Test/bench plumbing only."""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import capi

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "host", "libstorm_dropin.so")

GENERIC_SOLVERS = ("cg", "cgs", "bicgstab", "bicgstabl", "gmres", "fgmres", "tfqmr", "tfqmr1", "idrs",
                   "richardson")
NONLINEAR_SOLVERS = ("jfnk",)   # SolverNewton.hpp:101-173, inner solve = the reference's BiCgStabSolver
FUSED_SOLVERS = ("fused_cg", "fused_bicgstab", "fused_gmres")
# Storm/B200/GroupedSolvers.hpp: IDR(s) / BiCGStab(l) with their statements issued as sb_eval_group launches
GROUPED_SOLVERS = {"grouped_idrs": "idrs", "grouped_bicgstabl": "bicgstabl"}


class Opts(C.Structure):
    _fields_ = [("num_iterations", C.c_int64), ("abs_tol", C.c_double), ("rel_tol", C.c_double),
                ("num_inner_iterations", C.c_int64), ("relaxation_factor", C.c_double),
                ("use_graph", C.c_int32), ("precond", C.c_int32), ("pre_side", C.c_int32),
                ("cheb_degree", C.c_int32), ("cheb_power_iterations", C.c_int32), ("cheb_eig_ratio", C.c_double)]


class Report(C.Structure):
    _fields_ = [("converged", C.c_int32), ("iterations", C.c_int64), ("abs_err", C.c_double),
                ("rel_err", C.c_double), ("n_hist", C.c_int64), ("n_trace", C.c_int64),
                ("n_apply", C.c_int64)]


class ChParams(C.Structure):   # dropin_ch_params: Playground.cpp:113 constants + solver limits (<= 0 / < 0: defaults)
    _fields_ = [("tau", C.c_double), ("Gamma", C.c_double), ("sigma", C.c_double),
                ("num_iterations", C.c_int64), ("abs_tol", C.c_double), ("rel_tol", C.c_double), ("uniformed", C.c_int32)]


_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def load():
    global _lib
    if _lib is None:
        capi.load()  # libstormb200.so first (also found through the rpath)
        if not available():
            raise ImportError(f"{LIB_PATH} is missing: run `make -C stormruler_b200/host` where the "
                              "StormRuler sources are available")
        L = C.CDLL(LIB_PATH)
        L.dropin_solve.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                   C.POINTER(Opts), C.POINTER(Report), capi.f64p, C.c_int64, capi.f64p,
                                   C.c_int64]
        L.dropin_solve.restype = C.c_int
        L.dropin_cahn_hilliard_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_size_t, C.POINTER(ChParams), C.POINTER(Report), capi.f64p,
                                                C.c_int64, capi.f64p, C.c_int64]
        L.dropin_cahn_hilliard_step.restype = C.c_int
        L.dropin_solve_non_uniform.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_size_t, C.POINTER(Opts), C.POINTER(Report), capi.f64p, C.c_int64]
        L.dropin_solve_non_uniform.restype = C.c_int
        L.dropin_last_error.restype = C.c_char_p
        L.dropin_reset_rng.restype = None
        L.dropin_selftest_errors.argtypes = [C.c_void_p]
        L.dropin_selftest_errors.restype = C.c_int
        _lib = L
    return _lib


def set_statement_grouping(on) -> None:   # False / True / 2 (= on + dependency-aware scheduling)
    """Storm::B200::set_statement_grouping (DeviceVector.hpp): queue chain-shaped statements of the generic path and launch
    them as one sb_eval_group together with the reduction that follows. Off by default."""
    load().dropin_set_statement_grouping(int(on))


@dataclass
class Result:
    converged: bool
    iterations: int
    abs_err: float
    rel_err: float
    hist: np.ndarray
    trace: np.ndarray
    n_apply: int


PRE_SIDES = {"left": 0, "right": 1, "symmetric": 2}


def solve(name: str, op, x, b, num_iterations=2000, abs_tol=1e-6, rel_tol=1e-6, num_inner=0,
          relaxation_factor=0.0, use_graph=True, reset_rng=True, trace_cap=None, precond=None,
          pre_side="right", cheb_degree=0, cheb_power_iterations=0, cheb_eig_ratio=0.0) -> Result:
    """Run solver `name` on DeviceVectors x (in/out) and b through the C++ drop-in.
    precond: None | "jacobi" (Storm::JacobiPreconditioner) | "identity" (the reference's own) | "chebyshev"
    (Storm::ChebyshevPreconditioner: degree / power iterations / eigenvalue ratio, 0 = class defaults) in the
    reference's pre_op slot."""
    L = load()
    if reset_rng:
        L.dropin_reset_rng()
    cap_h = num_iterations + 2
    cap_t = trace_cap or (64 * num_iterations + 256)
    hist, trace = np.zeros(cap_h), np.zeros(cap_t)
    opts = Opts(num_iterations, abs_tol, rel_tol, num_inner, relaxation_factor, int(use_graph),
                {None: 0, "jacobi": 1, "identity": 2, "chebyshev": 3}[precond], PRE_SIDES[pre_side], int(cheb_degree), int(cheb_power_iterations),
                float(cheb_eig_ratio))
    rep = Report()
    rc = L.dropin_solve(name.encode(), op.ctx.handle, op.handle, x.ptr, b.ptr, x.n, C.byref(opts),
                        C.byref(rep), hist.ctypes.data_as(capi.f64p), cap_h,
                        trace.ctypes.data_as(capi.f64p), cap_t)
    if rc != 0:
        raise capi.StormB200Error(f"dropin_solve({name}) failed ({rc}): {L.dropin_last_error().decode()}")
    return Result(bool(rep.converged), rep.iterations, rep.abs_err, rep.rel_err,
                  hist[:min(rep.n_hist, cap_h)].copy(), trace[:min(rep.n_trace, cap_t)].copy(), rep.n_apply)


def cahn_hilliard_step(faces, c, c_hat, w_hat, tau=1.0e-3, Gamma=1.0e-4, sigma=2.0, num_iterations=0,
                       abs_tol=-1.0, rel_tol=-1.0, uniformed=False) -> Result:
    """One time step of the playground's Cahn-Hilliard solver (Playground.cpp:133-175) through the C++ drop-in:
    `faces` a faithful-form FvmOperator over the mesh, c (in) / c_hat (out: the new c) / w_hat (workspace)
    DeviceVectors. Defaults = the playground's constants (:113) and IterativeSolver's limits (2000, 1e-6, 1e-6)."""
    L = load()
    iters = num_iterations if num_iterations > 0 else 2000
    cap_h, cap_t = iters + 2, 8 * iters + 64
    hist, trace = np.zeros(cap_h), np.zeros(cap_t)
    prm = ChParams(tau, Gamma, sigma, num_iterations, abs_tol, rel_tol, int(uniformed))
    rep = Report()
    rc = L.dropin_cahn_hilliard_step(faces.ctx.handle, faces.handle, c.ptr, c_hat.ptr, w_hat.ptr, c.n,
                                     C.byref(prm), C.byref(rep), hist.ctypes.data_as(capi.f64p), cap_h,
                                     trace.ctypes.data_as(capi.f64p), cap_t)
    if rc != 0:
        raise capi.StormB200Error(f"dropin_cahn_hilliard_step failed ({rc}): {L.dropin_last_error().decode()}")
    return Result(bool(rep.converged), rep.iterations, rep.abs_err, rep.rel_err,
                  hist[:min(rep.n_hist, cap_h)].copy(), trace[:min(rep.n_trace, cap_t)].copy(), rep.n_apply)


def solve_non_uniform(name: str, op, x, b, shift, num_iterations=2000, abs_tol=1e-6, rel_tol=1e-6) -> Result:
    """The reference's solve_non_uniform (Solver.hpp:271-292) on DeviceVectors for the affine operator
    A(x) = op(x) + shift; name: cg | bicgstab | gmres | idrs."""
    L = load()
    L.dropin_reset_rng()
    cap_t = 64 * num_iterations + 256
    trace = np.zeros(cap_t)
    opts = Opts(num_iterations, abs_tol, rel_tol, 0, 0.0, 0, 0, 1)
    rep = Report()
    rc = L.dropin_solve_non_uniform(name.encode(), op.ctx.handle, op.handle, x.ptr, b.ptr, shift.ptr, x.n, C.byref(opts),
                                    C.byref(rep), trace.ctypes.data_as(capi.f64p), cap_t)
    if rc != 0:
        raise capi.StormB200Error(f"dropin_solve_non_uniform({name}) failed ({rc}): {L.dropin_last_error().decode()}")
    return Result(bool(rep.converged), rep.iterations, rep.abs_err, rep.rel_err, np.zeros(0),
                  trace[:min(rep.n_trace, cap_t)].copy(), rep.n_apply)


def random_program(op, init, seed: int, steps: int, with_accumulate=False, with_jacobi=False):
    """dropin_random_program (stormruler_b200/host/dropin.cpp): a seeded random program over a pool of DeviceVectors under
    the current statement-grouping mode. init: [n_vecs, n] host array. Returns (final vectors, recorded values)."""
    L = load()
    init = np.ascontiguousarray(init, np.float64)
    n_vecs, n = init.shape
    final, rec, nrec = np.zeros_like(init), np.zeros(2 * steps + 8), C.c_int64(0)
    L.dropin_random_program.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_int, C.c_void_p, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.c_int, C.c_int]
    L.dropin_random_program.restype = C.c_int
    rc = L.dropin_random_program(op.ctx.handle, op.handle, n, seed, steps, init.ctypes.data_as(C.c_void_p), n_vecs,
                                 final.ctypes.data_as(C.c_void_p), rec.ctypes.data_as(C.c_void_p), rec.shape[0],
                                 C.byref(nrec), int(with_accumulate), int(with_jacobi))
    if rc != 0:
        raise capi.StormB200Error(f"dropin_random_program failed ({rc}): {L.dropin_last_error().decode()}")
    return final, rec[:nrec.value].copy()
