"""stormruler_b200 -- B200-native Krylov hot path behind StormRuler's Operator/Solver API.

Python is plumbing here (tests, bench.py): thin objects over the C ABI of libstormb200.so
(include/stormb200.h), named after the reference's own types so parity tests read like the
reference: `DeviceVector` stands in for `CellField` (Feathers/Field.hpp:60-114), `FvmOperator` for
the playground's `FunctionalOperator` around `stormDivGrad` (Playground.cpp:115-131,153-167),
`CgSolver` / `BiCgStabSolver` for Solvers/SolverCg.hpp / SolverBiCgStab.hpp with the public knobs
of IterativeSolver (Solver.hpp:66-76). The C++23 drop-in (reference solver headers compiled
unchanged against the device vector) lives in stormruler_b200/host/.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import capi
from .capi import (ADD_ASSIGN, ASSIGN, DIV_ASSIGN, FORM_COEF, FORM_FAITHFUL, MUL_ASSIGN, SUB_ASSIGN,
                   StormB200Error)

__all__ = ["Context", "DeviceVector", "FvmOperator", "ConvDiffOperator", "CgSolver", "BiCgStabSolver", "GmresSolver", "StormB200Error",
           "FORM_COEF", "FORM_FAITHFUL", "ASSIGN", "ADD_ASSIGN", "SUB_ASSIGN", "MUL_ASSIGN",
           "DIV_ASSIGN", "expr"]


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Context:
    """One device, one compute stream (sb_ctx)."""

    def __init__(self, device: int = 0):
        self.lib = capi.load()
        h = C.c_void_p()
        capi.check(self.lib.sb_ctx_create(int(device), C.byref(h)))
        self.handle = h
        self.device = device

    def close(self):
        if getattr(self, "handle", None):
            self.lib.sb_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        capi.check(self.lib.sb_sync(self.handle))

    @property
    def stream(self) -> int:
        return int(self.lib.sb_ctx_stream(self.handle) or 0)

    @property
    def launch_count(self) -> int:
        return int(self.lib.sb_ctx_launch_count(self.handle))

    # -- vectors --------------------------------------------------------------------------------
    def zeros(self, n: int) -> "DeviceVector":
        return DeviceVector(self, n)

    def vector(self, host) -> "DeviceVector":
        host = _f64(host)
        v = DeviceVector(self, host.shape[0])
        v.upload(host)
        return v

    # -- BLAS-1 ---------------------------------------------------------------------------------
    def eval(self, y: "DeviceVector", assign_op: int, ops, vecs=(), scals=()):
        """y (op)= expr, with expr given as a postfix opcode list (see stormb200.h)."""
        e = capi.Expr()
        e.n_ops = len(ops)
        for k, o in enumerate(ops):
            e.ops[k] = o
        for k, v in enumerate(vecs):
            e.vec[k] = v.ptr
        for k, s in enumerate(scals):
            e.scal[k] = float(s)
        capi.check(self.lib.sb_eval(self.handle, y.ptr, y.n, assign_op, C.byref(e)))

    def eval_group(self, stmts=(), dots=()) -> np.ndarray:
        """sb_eval_group: `stmts` = [(y, base | None, [(c, x, sub), ...]), ...] run in order in one launch, then the
        dot products of `dots` = [(a, b), ...] over the final values; returns the dot values."""
        S = (capi.Chain * max(len(stmts), 1))()
        n = None
        for k, (y, base, terms) in enumerate(stmts):
            S[k].y, S[k].base, S[k].n_terms = y.ptr, (base.ptr if base is not None else None), len(terms)
            for t, (c, x, sub) in enumerate(terms):
                S[k].c[t], S[k].x[t], S[k].sub[t] = float(c), x.ptr, int(bool(sub))
            n = y.n
        m = len(dots)
        A = (C.c_void_p * max(m, 1))(*[p[0].ptr for p in dots])
        B = (C.c_void_p * max(m, 1))(*[p[1].ptr for p in dots])
        if n is None:
            n = dots[0][0].n if m else 0     # an empty group is rejected by the library
        out = np.zeros(max(m, 1))
        capi.check(self.lib.sb_eval_group(self.handle, n, len(stmts), S, m, A, B, out.ctypes.data_as(capi.f64p)))
        return out[:m]

    def dot(self, a: "DeviceVector", b: "DeviceVector") -> float:
        out = C.c_double()
        capi.check(self.lib.sb_dot(self.handle, a.ptr, b.ptr, a.n, C.byref(out)))
        return out.value

    def norm2(self, a: "DeviceVector") -> float:
        out = C.c_double()
        capi.check(self.lib.sb_norm2(self.handle, a.ptr, a.n, C.byref(out)))
        return out.value

    def dot_batch(self, pairs) -> np.ndarray:
        m = len(pairs)
        A = (C.c_void_p * m)(*[p[0].ptr for p in pairs])
        B = (C.c_void_p * m)(*[p[1].ptr for p in pairs])
        out = np.zeros(m)
        capi.check(self.lib.sb_dot_batch(self.handle, m, A, B, pairs[0][0].n,
                                         out.ctypes.data_as(capi.f64p)))
        return out


class DeviceVector:
    """Device-resident fp64 vector (zero-filled on allocation, like Field::assign)."""

    def __init__(self, ctx: Context, n: int):
        self.ctx, self.n = ctx, int(n)
        p = C.c_void_p()
        capi.check(ctx.lib.sb_vec_alloc(ctx.handle, self.n, C.byref(p)))
        self.ptr = p

    def __del__(self):
        try:
            if self.ptr and self.ctx.handle:
                self.ctx.lib.sb_vec_free(self.ctx.handle, self.ptr)
        except Exception:
            pass

    def upload(self, host):
        host = _f64(host)
        assert host.shape[0] == self.n
        capi.check(self.ctx.lib.sb_vec_upload(self.ctx.handle, self.ptr, host.ctypes.data_as(capi.f64p), self.n))
        return self

    def numpy(self) -> np.ndarray:
        out = np.empty(self.n)
        capi.check(self.ctx.lib.sb_vec_download(self.ctx.handle, self.ptr, out.ctypes.data_as(capi.f64p), self.n))
        return out

    def fill(self, value: float):
        capi.check(self.ctx.lib.sb_fill(self.ctx.handle, self.ptr, self.n, float(value)))
        return self

    def copy_from(self, other: "DeviceVector"):
        capi.check(self.ctx.lib.sb_copy(self.ctx.handle, self.ptr, other.ptr, self.n))
        return self


class expr:
    """Tiny postfix builder mirroring the Bittern operators the solvers use (MatrixMath.hpp:233-301):
    expr.v(r) + beta * (expr.v(p) - omega * expr.v(v))  ->  V0 S0 V1 S1 V2 MUL SUB MUL ADD."""

    def __init__(self, ops, vecs, scals):
        self.ops, self.vecs, self.scals = ops, vecs, scals

    @staticmethod
    def v(vec: DeviceVector) -> "expr":
        return expr([("v", 0)], [vec], [])

    @staticmethod
    def _lift(x) -> "expr":
        if isinstance(x, expr):
            return x
        if isinstance(x, DeviceVector):
            return expr.v(x)
        return expr([("s", 0)], [], [float(x)])

    def _bin(self, other, code, swap=False) -> "expr":
        a, b = (expr._lift(other), self) if swap else (self, expr._lift(other))
        ops = list(a.ops)
        for kind, k in b.ops:
            if kind == "v":
                ops.append(("v", k + len(a.vecs)))
            elif kind == "s":
                ops.append(("s", k + len(a.scals)))
            else:
                ops.append((kind, k))
        ops.append(("o", code))
        return expr(ops, a.vecs + b.vecs, a.scals + b.scals)

    def __add__(self, o): return self._bin(o, capi.OP_ADD)
    def __radd__(self, o): return self._bin(o, capi.OP_ADD, swap=True)
    def __sub__(self, o): return self._bin(o, capi.OP_SUB)
    def __rsub__(self, o): return self._bin(o, capi.OP_SUB, swap=True)
    def __mul__(self, o): return self._bin(o, capi.OP_MUL)
    def __rmul__(self, o): return self._bin(o, capi.OP_MUL, swap=True)
    def __truediv__(self, o): return self._bin(o, capi.OP_DIV)
    def __neg__(self): return expr(self.ops + [("o", capi.OP_NEG)], self.vecs, self.scals)

    def encode(self):
        # merge identical vector operands so `a + beta*a` uses one slot
        uniq, remap = [], {}
        for k, v in enumerate(self.vecs):
            for q, u in enumerate(uniq):
                if u is v:
                    remap[k] = q
                    break
            else:
                remap[k] = len(uniq)
                uniq.append(v)
        code = []
        for kind, k in self.ops:
            code.append(capi.OP_VEC0 + remap[k] if kind == "v" else capi.OP_SCAL0 + k if kind == "s" else k)
        return code, uniq, self.scals

    def assign_to(self, y: DeviceVector, assign_op: int = ASSIGN):
        code, vecs, scals = self.encode()
        y.ctx.eval(y, assign_op, code, vecs, scals)


class FvmOperator:
    """y <- A(x): the matrix-free FVM operator as uploaded cell rows (sb_op)."""

    def __init__(self, ctx: Context, mesh, prefill: int, dt: float, form: int = FORM_COEF,
                 dirichlet: bool = False):
        """`mesh` needs attributes n_cells, face_cell [F,2], face_area, face_dist, cell_vol and
        bface_cell, bface_area, bface_dist (used when dirichlet=True)."""
        self.ctx = ctx
        fc = np.ascontiguousarray(mesh.face_cell, np.int32).reshape(-1)
        fa, fd, cv = _f64(mesh.face_area), _f64(mesh.face_dist), _f64(mesh.cell_vol)
        nb = int(len(mesh.bface_area)) if dirichlet else 0
        bc = np.ascontiguousarray(mesh.bface_cell, np.int32)
        ba, bd = _f64(mesh.bface_area), _f64(mesh.bface_dist)
        soa = capi.MeshSoa(int(mesh.n_cells), int(fa.shape[0]), fc.ctypes.data_as(capi.i32p),
                           fa.ctypes.data_as(capi.f64p), fd.ctypes.data_as(capi.f64p),
                           cv.ctypes.data_as(capi.f64p), nb, bc.ctypes.data_as(capi.i32p),
                           ba.ctypes.data_as(capi.f64p), bd.ctypes.data_as(capi.f64p))
        desc = capi.OpDesc(int(form), int(prefill), float(dt))
        h = C.c_void_p()
        capi.check(ctx.lib.sb_op_create(ctx.handle, C.byref(soa), C.byref(desc), C.byref(h)))
        self.handle = h
        info = capi.OpInfo()
        capi.check(ctx.lib.sb_op_get_info(h, C.byref(info)))
        self.info = info
        self.n = int(info.n_cells)

    def __del__(self):
        try:
            if self.handle and self.ctx.handle:
                self.ctx.lib.sb_op_destroy(self.ctx.handle, self.handle)
        except Exception:
            pass

    def mul(self, y: DeviceVector, x: DeviceVector):
        """Operator::mul (Operator.hpp:74)."""
        capi.check(self.ctx.lib.sb_apply(self.ctx.handle, self.handle, x.ptr, y.ptr))

    def mul_dot(self, y: DeviceVector, x: DeviceVector, u: DeviceVector | None = None) -> float:
        """y <- A(x) and <u, y> (u None: <x, y>) in one kernel (sb_apply_dot)."""
        out = C.c_double()
        capi.check(self.ctx.lib.sb_apply_dot(self.ctx.handle, self.handle, x.ptr, y.ptr, u.ptr if u is not None else None,
                                             C.byref(out)))
        return out.value

    def mul_dot_yy_yx(self, y: DeviceVector, x: DeviceVector):
        """y <- A(x); returns (<y, y>, <y, x>) from the same kernel (sb_apply_dot_yy_yx)."""
        out = np.zeros(2)
        capi.check(self.ctx.lib.sb_apply_dot_yy_yx(self.ctx.handle, self.handle, x.ptr, y.ptr, out.ctypes.data_as(capi.f64p)))
        return float(out[0]), float(out[1])

    def div_grad(self, u: DeviceVector, dt: float, c: DeviceVector):
        """u += dt * div grad c: `stormDivGrad(mesh, u, dt, c)` as the playground calls it (Playground.cpp:115-131);
        faithful-form operators only (sb_apply_accumulate)."""
        capi.check(self.ctx.lib.sb_apply_accumulate(self.ctx.handle, self.handle, float(dt), c.ptr, u.ptr))

    def jacobi(self, y: DeviceVector, x: DeviceVector):
        """y = D^-1 x with D the operator's diagonal (sb_op_jacobi): the Jacobi preconditioner's mul."""
        capi.check(self.ctx.lib.sb_op_jacobi(self.ctx.handle, self.handle, x.ptr, y.ptr))

    def rows(self):
        """Download the row layout (col, val0, val1|None, diag|None) for integer/bit checks."""
        w, ld = self.info.width, self.info.ld
        col = np.empty((w, ld), np.int32)
        v0 = np.empty((w, ld))
        faithful = self.info.form == FORM_FAITHFUL
        v1 = np.empty((w, ld)) if faithful else None
        diag = None if faithful else np.empty(ld)
        capi.check(self.ctx.lib.sb_op_download_rows(
            self.ctx.handle, self.handle, col.ctypes.data_as(capi.i32p), v0.ctypes.data_as(capi.f64p),
            v1.ctypes.data_as(capi.f64p) if faithful else None,
            None if faithful else diag.ctypes.data_as(capi.f64p)))
        return col, v0, v1, diag


class ConvDiffOperator(FvmOperator):
    """y = -nu div grad x + div(beta x), first-order upwind in the face-loop pattern of
    UpwindConvectionScheme (Feathers/ConvectionScheme.hpp:83-106), Dirichlet mirror ghosts: non-symmetric
    coefficient rows applied by the same kernels (sb_op_create_convdiff). `face_un` / `bface_un` = beta . n per
    interior / boundary face (Mesh.face_flux)."""

    def __init__(self, ctx: Context, mesh, nu: float, face_un, bface_un):
        self.ctx = ctx
        fc = np.ascontiguousarray(mesh.face_cell, np.int32).reshape(-1)
        fa, fd, cv = _f64(mesh.face_area), _f64(mesh.face_dist), _f64(mesh.cell_vol)
        bc = np.ascontiguousarray(mesh.bface_cell, np.int32)
        ba, bd = _f64(mesh.bface_area), _f64(mesh.bface_dist)
        fu, bu = _f64(face_un), _f64(bface_un)
        assert fu.shape[0] == fa.shape[0] and bu.shape[0] == ba.shape[0]
        soa = capi.MeshSoa(int(mesh.n_cells), int(fa.shape[0]), fc.ctypes.data_as(capi.i32p),
                           fa.ctypes.data_as(capi.f64p), fd.ctypes.data_as(capi.f64p),
                           cv.ctypes.data_as(capi.f64p), int(ba.shape[0]), bc.ctypes.data_as(capi.i32p),
                           ba.ctypes.data_as(capi.f64p), bd.ctypes.data_as(capi.f64p))
        desc = capi.ConvDiffDesc(float(nu), fu.ctypes.data_as(capi.f64p), bu.ctypes.data_as(capi.f64p))
        h = C.c_void_p()
        capi.check(ctx.lib.sb_op_create_convdiff(ctx.handle, C.byref(soa), C.byref(desc), C.byref(h)))
        self.handle = h
        info = capi.OpInfo()
        capi.check(ctx.lib.sb_op_get_info(h, C.byref(info)))
        self.info = info
        self.n = int(info.n_cells)


@dataclass
class _FusedSolver:
    """Public knobs and progress fields of IterativeSolver (Solver.hpp:66-76)."""
    num_iterations: int = 2000
    absolute_error_tolerance: float = 1.0e-6
    relative_error_tolerance: float = 1.0e-6
    iteration: int = 0
    absolute_error: float = 0.0
    relative_error: float = 0.0
    check_every: int = 0
    use_graph: bool = False
    profile: bool = False
    record: bool = True
    schedule: int = 0          # SCHEDULE_AUTO / SCHEDULE_STEPWISE / SCHEDULE_PERSISTENT (include/stormb200.h)
    timeline_iters: int = 0    # persistent schedule: in-kernel timeline of the first iterations -> self.timeline
    tuning: int = 0            # capi.TUNE_* bits of the stepwise schedule (0 = the library's defaults)
    history: np.ndarray = field(default_factory=lambda: np.zeros(0))
    trace: np.ndarray = field(default_factory=lambda: np.zeros(0))
    timeline: np.ndarray = field(default_factory=lambda: np.zeros((0, 20), np.uint64))
    solve_ms: float = 0.0
    iter_ms: float = 0.0
    kernel_ms: tuple = ()
    wait_ms: tuple = ()
    ar_wait_ms: tuple = ()
    final_ms: tuple = ()
    launches: int = 0
    schedule_used: int = 0
    _entry = ""
    _trace_per_iter = 2

    def solve(self, x: DeviceVector, b: DeviceVector, op: FvmOperator) -> bool:
        lib = op.ctx.lib
        tl = np.zeros((max(int(self.timeline_iters), 1), capi.TIMELINE_WORDS), np.uint64)
        opts = capi.SolverOpts(int(self.num_iterations), float(self.absolute_error_tolerance),
                               float(self.relative_error_tolerance), int(self.check_every),
                               int(bool(self.use_graph)), int(bool(self.profile)), int(self.schedule),
                               int(self.timeline_iters), tl.ctypes.data_as(C.POINTER(C.c_uint64)), int(self.tuning))
        rep = capi.SolverReport()
        cap_h = self.num_iterations + 2 if self.record else 0
        cap_t = self._trace_per_iter * self.num_iterations + 8 if self.record else 0
        hist, trace = np.zeros(max(cap_h, 1)), np.zeros(max(cap_t, 1))
        fn = getattr(lib, self._entry)
        capi.check(fn(op.ctx.handle, op.handle, x.ptr, b.ptr, C.byref(opts), C.byref(rep),
                      hist.ctypes.data_as(capi.f64p) if self.record else None, cap_h,
                      trace.ctypes.data_as(capi.f64p) if self.record else None, cap_t))
        self.iteration = int(rep.iterations)
        self.absolute_error, self.relative_error = rep.abs_err, rep.rel_err
        self.history, self.trace = hist[:rep.n_hist].copy(), trace[:rep.n_trace].copy()
        self.solve_ms, self.iter_ms, self.launches = rep.solve_ms, rep.iter_ms, int(rep.launches)
        self.kernel_ms = tuple(rep.kernel_ms[k] for k in range(rep.n_kernel_slots))
        self.wait_ms = tuple(rep.wait_ms[k] for k in range(rep.n_kernel_slots))
        self.ar_wait_ms = tuple(rep.ar_wait_ms[k] for k in range(rep.n_kernel_slots))
        self.final_ms = tuple(rep.final_ms[k] for k in range(rep.n_kernel_slots))
        self.schedule_used = int(rep.schedule)
        self.timeline = tl[:min(int(self.timeline_iters), self.iteration)] if rep.schedule == capi.SCHEDULE_PERSISTENT else tl[:0]
        return bool(rep.converged)


@dataclass
class CgSolver(_FusedSolver):
    _entry = "sb_cg_solve"
    _trace_per_iter = 2


@dataclass
class BiCgStabSolver(_FusedSolver):
    _entry = "sb_bicgstab_solve"
    _trace_per_iter = 5


@dataclass
class GmresSolver:
    """Fused restarted GMRES(m) (sb_gmres_solve): GmresSolver / FgmresSolver of SolverGmres.hpp without a
    preconditioner, with the public knobs of InnerOuterIterativeSolver (Solver.hpp:158-159)."""
    num_iterations: int = 2000
    absolute_error_tolerance: float = 1.0e-6
    relative_error_tolerance: float = 1.0e-6
    num_inner_iterations: int = 50
    lookahead: int = 0
    iteration: int = 0
    absolute_error: float = 0.0
    relative_error: float = 0.0
    record: bool = True
    history: np.ndarray = field(default_factory=lambda: np.zeros(0))
    trace: np.ndarray = field(default_factory=lambda: np.zeros(0))
    solve_ms: float = 0.0
    iter_ms: float = 0.0
    launches: int = 0

    def solve(self, x: DeviceVector, b: DeviceVector, op) -> bool:
        lib = op.ctx.lib
        opts = capi.GmresOpts(int(self.num_iterations), float(self.absolute_error_tolerance),
                              float(self.relative_error_tolerance), int(self.num_inner_iterations), int(self.lookahead))
        rep = capi.SolverReport()
        m = self.num_inner_iterations or 50
        cap_h = self.num_iterations + 2 if self.record else 0
        cap_t = (m + 3) * (self.num_iterations + 2) + 8 if self.record else 0
        hist, trace = np.zeros(max(cap_h, 1)), np.zeros(max(cap_t, 1))
        capi.check(lib.sb_gmres_solve(op.ctx.handle, op.handle, x.ptr, b.ptr, C.byref(opts), C.byref(rep),
                                      hist.ctypes.data_as(capi.f64p) if self.record else None, cap_h,
                                      trace.ctypes.data_as(capi.f64p) if self.record else None, cap_t))
        self.iteration = int(rep.iterations)
        self.absolute_error, self.relative_error = rep.abs_err, rep.rel_err
        self.history, self.trace = hist[:rep.n_hist].copy(), trace[:rep.n_trace].copy()
        self.solve_ms, self.iter_ms, self.launches = rep.solve_ms, rep.iter_ms, int(rep.launches)
        return bool(rep.converged)


def solve_host(ctx: Context, op: FvmOperator, solver: str, x_host: np.ndarray, b_host: np.ndarray,
               num_iterations=2000, abs_tol=1e-6, rel_tol=1e-6, use_graph=False, check_every=0, schedule=0, tuning=0):
    """sb_solve_host: host buffers in/out, copies inside the call. x_host is updated in place."""
    assert x_host.dtype == np.float64 and x_host.flags.c_contiguous
    b_host = _f64(b_host)
    opts = capi.SolverOpts(int(num_iterations), float(abs_tol), float(rel_tol), int(check_every), int(use_graph), 0,
                           int(schedule), 0, None, int(tuning))
    rep = capi.SolverReport()
    capi.check(ctx.lib.sb_solve_host(ctx.handle, op.handle, solver.encode(), x_host.ctypes.data_as(capi.f64p),
                                     b_host.ctypes.data_as(capi.f64p), C.byref(opts), C.byref(rep), None, 0))
    return rep
