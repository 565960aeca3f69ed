"""TEST INFRASTRUCTURE -- not product code.

Runs the product's drop-in translation unit (stormruler_b200/host/dropin.cpp: StormRuler's own solver templates on
Storm::DeviceVector, and the playground's Cahn-Hilliard step) on a HOST-EXECUTING stand-in of the C ABI
(oracle/emu/emu_stormb200.c, built by `make -C oracle emu`; needs the StormRuler sources at build time only). Every
vector statement the C++23 host layer issues is executed on numpy-visible host arrays by the contract that
include/stormb200.h documents, so the CPU tests can compare the host layer with the reference's own run bit for bit:

    ref (reference headers on a host vector)  ==  drop-in TU on the emulator      [here, no GPU]
    drop-in TU on the emulator (tree mode)    ==  drop-in TU on libstormb200.so   [GPU tests]
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import orc

HERE = os.path.dirname(os.path.abspath(__file__))
DROPIN = os.path.join(HERE, "_ref", "emu", "liboracle_dropin_on_emulator.so")
EMU = os.path.join(HERE, "_ref", "emu", "liboracle_cabi_emulator.so")

COUNT_NAMES = ("eval", "fill", "copy", "dot", "norm", "apply", "accumulate", "jacobi")
PRE_SIDES = {"left": 0, "right": 1, "symmetric": 2}


class Opts(C.Structure):      # dropin_opts of stormruler_b200/host/dropin.cpp
    _fields_ = [("num_iterations", C.c_int64), ("abs_tol", C.c_double), ("rel_tol", C.c_double),
                ("num_inner_iterations", C.c_int64), ("relaxation_factor", C.c_double),
                ("use_graph", C.c_int32), ("precond", C.c_int32), ("pre_side", C.c_int32),
                ("cheb_degree", C.c_int32), ("cheb_power_iterations", C.c_int32), ("cheb_eig_ratio", C.c_double)]


class Report(C.Structure):
    _fields_ = [("converged", C.c_int32), ("iterations", C.c_int64), ("abs_err", C.c_double),
                ("rel_err", C.c_double), ("n_hist", C.c_int64), ("n_trace", C.c_int64), ("n_apply", C.c_int64)]


class ChParams(C.Structure):  # dropin_ch_params
    _fields_ = [("tau", C.c_double), ("Gamma", C.c_double), ("sigma", C.c_double),
                ("num_iterations", C.c_int64), ("abs_tol", C.c_double), ("rel_tol", C.c_double), ("uniformed", C.c_int32)]


def available() -> bool:
    return os.path.exists(DROPIN) and os.path.exists(EMU)


_libs = None


def _load():
    global _libs
    if _libs is None:
        # RTLD_LOCAL (ctypes' default): the stand-in's sb_* symbols stay private to the emulated drop-in
        dr = C.CDLL(DROPIN)
        em = C.CDLL(EMU)            # same instance: already loaded as the drop-in's dependency
        em.emu_ctx.restype = C.c_void_p
        em.emu_op_create.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        em.emu_op_create.restype = C.c_void_p
        em.emu_op_free.argtypes = [C.c_void_p]
        em.emu_set_reduction_mode.argtypes = [C.c_int]
        em.emu_get_counts.argtypes = [C.POINTER(C.c_int64)]
        dr.dropin_solve.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                    C.POINTER(Opts), C.POINTER(Report), orc._f64p, C.c_int64, orc._f64p, C.c_int64]
        dr.dropin_solve.restype = C.c_int
        dr.dropin_cahn_hilliard_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                 C.c_size_t, C.POINTER(ChParams), C.POINTER(Report), orc._f64p,
                                                 C.c_int64, orc._f64p, C.c_int64]
        dr.dropin_cahn_hilliard_step.restype = C.c_int
        dr.dropin_last_error.restype = C.c_char_p
        dr.dropin_reset_rng.restype = None
        dr.dropin_selftest_errors.argtypes = [C.c_void_p]
        dr.dropin_selftest_errors.restype = C.c_int
        _libs = (em, dr)
    return _libs


class EmuOp:
    """An emulator operator handle over an oracle operator object (FaceOp, RowsOp, ConvDiffOp, CallbackOp):
    sb_apply -> its callback; sb_apply_accumulate -> the oracle face loop (FaceOp only); sb_op_jacobi -> x / diag."""

    def __init__(self, op, diag=None):
        em, _ = _load()
        self.op, self.n = op, op.n
        fn, user = op.callback
        faces = C.cast(C.pointer(op.struct), C.c_void_p) if isinstance(op, orc.FaceOp) else None
        self._diag = None if diag is None else np.ascontiguousarray(diag, np.float64)
        dptr = None if self._diag is None else self._diag.ctypes.data_as(C.c_void_p)
        self.handle = em.emu_op_create(self.n, fn, user, faces, dptr)

    def __del__(self):
        try:
            _load()[0].emu_op_free(self.handle)
        except Exception:
            pass


def counts() -> dict:
    out = (C.c_int64 * 8)()
    _load()[0].emu_get_counts(out)
    return dict(zip(COUNT_NAMES, list(out)))


def set_statement_grouping(on) -> None:   # False / True / 2 (= on + dependency-aware scheduling)
    """Storm::B200::set_statement_grouping in the emulated drop-in (off by default)."""
    _load()[1].dropin_set_statement_grouping(int(on))


def group_count() -> int:
    """Number of sb_eval_group launches since the last solve started."""
    em = _load()[0]
    em.emu_group_count.restype = C.c_int64
    return int(em.emu_group_count())


def apply_dot_count() -> int:
    """Number of sb_apply_dot calls (an apply with the following dot riding on it) since the last solve started."""
    em = _load()[0]
    em.emu_apply_dot_count.restype = C.c_int64
    return int(em.emu_apply_dot_count())


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def solve(name: str, op: EmuOp, b, x0=None, num_iterations=2000, abs_tol=1e-6, rel_tol=1e-6, num_inner=0,
          relaxation_factor=0.0, mode=orc.RED_SEQ, precond=None, pre_side="right", reset_rng=True,
          trace_cap=None, cheb_degree=0, cheb_power_iterations=0, cheb_eig_ratio=0.0) -> orc.SolveResult:
    """dropin_solve on the emulator: the reference template `name` on Storm::DeviceVector, host-executed."""
    em, dr = _load()
    em.emu_set_reduction_mode(mode)
    em.emu_reset_counts()
    if reset_rng:
        dr.dropin_reset_rng()
    b = np.ascontiguousarray(b, np.float64)
    n = b.shape[0]
    x = np.zeros(n) if x0 is None else np.ascontiguousarray(x0, np.float64).copy()
    cap_h, cap_t = num_iterations + 2, trace_cap or (64 * num_iterations + 256)
    hist, trace = np.zeros(cap_h), np.zeros(cap_t)
    opts = Opts(num_iterations, abs_tol, rel_tol, num_inner, relaxation_factor, 0,
                {None: 0, "jacobi": 1, "identity": 2, "chebyshev": 3}[precond], PRE_SIDES[pre_side], int(cheb_degree),
                int(cheb_power_iterations), float(cheb_eig_ratio))
    rep = Report()
    rc = dr.dropin_solve(name.encode(), em.emu_ctx(), op.handle, _p(x), _p(b), n, C.byref(opts), C.byref(rep),
                         hist.ctypes.data_as(orc._f64p), cap_h, trace.ctypes.data_as(orc._f64p), cap_t)
    if rc != 0:
        raise RuntimeError(f"dropin_solve({name}) on the emulator failed ({rc}): {dr.dropin_last_error().decode()}")
    return orc.SolveResult(x, bool(rep.converged), rep.iterations, rep.abs_err, rep.rel_err,
                           hist[:min(rep.n_hist, cap_h)].copy(), trace[:min(rep.n_trace, cap_t)].copy(), rep.n_apply)


def cahn_hilliard_step(faces: EmuOp, c, mode=orc.RED_SEQ, tau=orc.CH_TAU, Gamma=orc.CH_GAMMA, sigma=orc.CH_SIGMA,
                       num_iterations=0, abs_tol=-1.0, rel_tol=-1.0, uniformed=False):
    """dropin_cahn_hilliard_step on the emulator. Returns (SolveResult with x = the new c, w_hat)."""
    em, dr = _load()
    em.emu_set_reduction_mode(mode)
    em.emu_reset_counts()
    c = np.ascontiguousarray(c, np.float64)
    n = c.shape[0]
    c_hat, w_hat = np.zeros(n), np.zeros(n)
    iters = num_iterations if num_iterations > 0 else 2000
    cap_h, cap_t = iters + 2, 8 * iters + 64
    hist, trace = np.zeros(cap_h), np.zeros(cap_t)
    prm = ChParams(tau, Gamma, sigma, num_iterations, abs_tol, rel_tol, int(uniformed))
    rep = Report()
    rc = dr.dropin_cahn_hilliard_step(em.emu_ctx(), faces.handle, _p(c), _p(c_hat), _p(w_hat), n, C.byref(prm),
                                      C.byref(rep), hist.ctypes.data_as(orc._f64p), cap_h,
                                      trace.ctypes.data_as(orc._f64p), cap_t)
    if rc != 0:
        raise RuntimeError(f"dropin_cahn_hilliard_step on the emulator failed ({rc}): {dr.dropin_last_error().decode()}")
    res = orc.SolveResult(c_hat, bool(rep.converged), rep.iterations, rep.abs_err, rep.rel_err,
                          hist[:min(rep.n_hist, cap_h)].copy(), trace[:min(rep.n_trace, cap_t)].copy(), rep.n_apply)
    return res, w_hat


def solve_non_uniform(name: str, op: EmuOp, b, shift, x0=None, num_iterations=2000, abs_tol=1e-6, rel_tol=1e-6,
                      mode=orc.RED_SEQ) -> orc.SolveResult:
    """dropin_solve_non_uniform on the emulator: the reference's solve_non_uniform template on Storm::DeviceVector."""
    em, dr = _load()
    em.emu_set_reduction_mode(mode)
    em.emu_reset_counts()
    dr.dropin_solve_non_uniform.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_size_t, C.POINTER(Opts), C.POINTER(Report), orc._f64p, C.c_int64]
    dr.dropin_solve_non_uniform.restype = C.c_int
    b, shift = np.ascontiguousarray(b, np.float64), np.ascontiguousarray(shift, np.float64)
    n = b.shape[0]
    x = np.zeros(n) if x0 is None else np.ascontiguousarray(x0, np.float64).copy()
    cap_t = 64 * num_iterations + 256
    trace = np.zeros(cap_t)
    opts = Opts(num_iterations, abs_tol, rel_tol, 0, 0.0, 0, 0, 1, 0, 0, 0.0)
    rep = Report()
    rc = dr.dropin_solve_non_uniform(name.encode(), em.emu_ctx(), op.handle, _p(x), _p(b), _p(shift), n, C.byref(opts),
                                     C.byref(rep), trace.ctypes.data_as(orc._f64p), cap_t)
    if rc != 0:
        raise RuntimeError(f"dropin_solve_non_uniform({name}) failed ({rc}): {dr.dropin_last_error().decode()}")
    return orc.SolveResult(x, bool(rep.converged), rep.iterations, rep.abs_err, rep.rel_err, np.zeros(0),
                           trace[:min(rep.n_trace, cap_t)].copy(), rep.n_apply)


def random_program(op: EmuOp, init, seed: int, steps: int, mode=orc.RED_SEQ, with_accumulate=False, with_jacobi=False):
    """dropin_random_program on the emulator under the current grouping mode: (final vectors [n_vecs, n], recorded values)."""
    em, dr = _load()
    em.emu_set_reduction_mode(mode)
    em.emu_reset_counts()
    init = np.ascontiguousarray(init, np.float64)
    n_vecs, n = init.shape
    final, rec, nrec = np.zeros_like(init), np.zeros(steps + 8), C.c_int64(0)
    dr.dropin_random_program.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_int, C.c_void_p, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.c_int, C.c_int]
    dr.dropin_random_program.restype = C.c_int
    final, rec = np.zeros_like(init), np.zeros(2 * steps + 8)
    rc = dr.dropin_random_program(em.emu_ctx(), op.handle, n, seed, steps, _p(init), n_vecs, _p(final), _p(rec), rec.shape[0],
                                  C.byref(nrec), int(with_accumulate), int(with_jacobi))
    if rc != 0:
        raise RuntimeError(f"dropin_random_program failed ({rc}): {dr.dropin_last_error().decode()}")
    return final, rec[:nrec.value].copy()


def selftest_errors() -> int:
    em, dr = _load()
    return dr.dropin_selftest_errors(em.emu_ctx())
