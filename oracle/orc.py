"""ctypes bindings of the parity oracle -- TEST INFRASTRUCTURE, not product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` leg
may import this module. It wraps

* ``oracle/liboracle.so``            -- the plain-C restatement (oracle/sb_oracle.c), and
* ``oracle/_ref/libref_solvers.so``  -- the reference's own solver headers compiled verbatim
  (oracle/ref_build/ref_solvers.cpp), when it has been built (needs /root/reference at build
  time only; the built file travels to the GPU box).

Both are built by ``make -C oracle`` (``__graft_entry__.build()`` runs it).
"""
from __future__ import annotations

import ctypes as C
import os
import struct
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_LIB_PATH = os.path.join(HERE, "_ref", "libref_solvers.so")
REF_MESH_TOOL = os.path.join(HERE, "_ref", "ref_mesh_tool")

RED_SEQ, RED_TREE, RED_TREE_SEG = 0, 1, 2
COL_PAD = -(2**31)

_f64p = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)

APPLY_FN = C.CFUNCTYPE(None, C.c_void_p, _f64p, _f64p, C.c_size_t)


def build(ref: bool | None = None) -> None:
    """Run the oracle Makefile (liboracle.so always; _ref/ when the reference is mounted)."""
    target = "all" if ref is None else ("ref" if ref else "oracle")
    subprocess.run(["make", "-C", HERE, target], check=True, capture_output=True)


class FaceOpStruct(C.Structure):
    _fields_ = [
        ("n_cells", C.c_int64), ("n_faces", C.c_int64),
        ("face_cell", _i32p), ("face_area", _f64p), ("face_dist", _f64p), ("cell_vol", _f64p),
        ("n_bfaces", C.c_int64),
        ("bface_cell", _i32p), ("bface_area", _f64p), ("bface_dist", _f64p),
        ("prefill", C.c_int32), ("dt", C.c_double),
    ]


class ConvDiffStruct(C.Structure):
    _fields_ = [("base", FaceOpStruct), ("nu", C.c_double), ("face_un", _f64p), ("bface_un", _f64p)]


class RowsOpStruct(C.Structure):
    _fields_ = [("n", C.c_int64), ("width", C.c_int), ("ld", C.c_int64), ("col", _i32p), ("a", _f64p),
                ("diag", _f64p)]


class SolverOpts(C.Structure):
    _fields_ = [("num_iterations", C.c_int64), ("abs_tol", C.c_double), ("rel_tol", C.c_double),
                ("reduction_mode", C.c_int32)]


class SolverReport(C.Structure):
    _fields_ = [("converged", C.c_int32), ("iterations", C.c_int64), ("abs_err", C.c_double),
                ("rel_err", C.c_double), ("n_hist", C.c_int64), ("n_trace", C.c_int64)]


class RefOpts(C.Structure):
    _fields_ = [("num_iterations", C.c_int64), ("abs_tol", C.c_double), ("rel_tol", C.c_double),
                ("num_inner_iterations", C.c_int64), ("reduction_mode", C.c_int32),
                ("relaxation_factor", C.c_double), ("pre_fn", C.c_void_p), ("pre_user", C.c_void_p),
                ("pre_side", C.c_int32), ("pre_kind", C.c_int32), ("cheb_degree", C.c_int32),
                ("cheb_power_iterations", C.c_int32), ("cheb_eig_ratio", C.c_double)]


class RefReport(C.Structure):
    _fields_ = [("converged", C.c_int32), ("iterations", C.c_int64), ("abs_err", C.c_double),
                ("rel_err", C.c_double), ("n_hist", C.c_int64), ("n_trace", C.c_int64),
                ("n_apply", C.c_int64)]


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build(ref=False)
        L = C.CDLL(LIB_PATH)
        L.orc_apply_faces.argtypes = [C.POINTER(FaceOpStruct), _f64p, _f64p]
        L.orc_apply_faces.restype = None
        L.orc_divgrad_accumulate.argtypes = [C.POINTER(FaceOpStruct), C.c_double, _f64p, _f64p]
        L.orc_divgrad_accumulate.restype = None
        L.orc_rows_width.argtypes = [C.POINTER(FaceOpStruct)]
        L.orc_rows_width.restype = C.c_int
        L.orc_build_rows.argtypes = [C.POINTER(FaceOpStruct), C.c_int, C.c_int64, _i32p, _i64p]
        L.orc_build_rows.restype = None
        L.orc_rows_faithful.argtypes = [C.POINTER(FaceOpStruct), C.c_int, C.c_int64, _i64p, _f64p, _f64p]
        L.orc_rows_faithful.restype = None
        L.orc_apply_rows_faithful.argtypes = [C.c_int64, C.c_int, C.c_int64, _i32p, _f64p, _f64p,
                                              C.c_int32, C.c_double, _f64p, _f64p]
        L.orc_apply_rows_faithful.restype = None
        L.orc_rows_coef.argtypes = [C.POINTER(FaceOpStruct), C.c_int, C.c_int64, _i32p, _i64p, _i32p,
                                    _f64p, _f64p]
        L.orc_rows_coef.restype = None
        L.orc_apply_rows_coef.argtypes = [C.c_int64, C.c_int, C.c_int64, _i32p, _f64p, _f64p, _f64p, _f64p]
        L.orc_apply_rows_coef.restype = None
        L.orc_apply_convdiff_faces.argtypes = [C.POINTER(ConvDiffStruct), _f64p, _f64p]
        L.orc_apply_convdiff_faces.restype = None
        L.orc_rows_convdiff.argtypes = [C.POINTER(ConvDiffStruct), C.c_int, C.c_int64, _i32p, _f64p, _f64p]
        L.orc_rows_convdiff.restype = None
        L.orc_dot.argtypes = [C.c_int64, _f64p, _f64p, C.c_int]
        L.orc_dot.restype = C.c_double
        L.orc_norm2.argtypes = [C.c_int64, _f64p, C.c_int]
        L.orc_norm2.restype = C.c_double
        L.orc_set_segments.argtypes = [C.c_int, _i64p]
        L.orc_set_segments.restype = None
        L.orc_safe_divide.argtypes = [C.c_double, C.c_double]
        L.orc_safe_divide.restype = C.c_double
        for name in ("orc_cg", "orc_bicgstab"):
            fn = getattr(L, name)
            fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, _f64p, _f64p, C.POINTER(SolverOpts),
                           C.POINTER(SolverReport), _f64p, C.c_int64, _f64p, C.c_int64]
            fn.restype = C.c_int
        _lib = L
    return _lib


def have_ref() -> bool:
    return os.path.exists(REF_LIB_PATH)


def ref():
    global _ref
    if _ref is None:
        if not have_ref():
            raise FileNotFoundError(
                f"{REF_LIB_PATH} not built (needs /root/reference; run `make -C oracle ref`)")
        R = C.CDLL(REF_LIB_PATH)
        R.ref_solve.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_void_p, _f64p, _f64p,
                                C.POINTER(RefOpts), C.POINTER(RefReport), _f64p, C.c_int64, _f64p,
                                C.c_int64]
        R.ref_solve.restype = C.c_int
        R.ref_reset_rng.restype = None
        R.ref_fill_randomly_generic.argtypes = [C.c_size_t, _f64p]
        R.ref_fill_randomly_generic.restype = None
        R.ref_dot.argtypes = [C.c_size_t, _f64p, _f64p]
        R.ref_dot.restype = C.c_double
        R.ref_norm2.argtypes = [C.c_size_t, _f64p]
        R.ref_norm2.restype = C.c_double
        R.ref_build_info.restype = C.c_char_p
        _ref = R
    return _ref


def _p(a, typ):
    return a.ctypes.data_as(typ)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


@dataclass
class FaceMesh:
    """Face-list SoA of one mesh (the arrays sb_mesh_soa / orc_face_op point at)."""
    n_cells: int
    face_cell: np.ndarray   # int32 [F,2] inner, outer
    face_area: np.ndarray   # f64 [F]
    face_dist: np.ndarray   # f64 [F]
    cell_vol: np.ndarray    # f64 [N]
    bface_cell: np.ndarray  # int32 [B]
    bface_area: np.ndarray  # f64 [B]
    bface_dist: np.ndarray  # f64 [B]

    @property
    def n_faces(self) -> int:
        return int(self.face_area.shape[0])

    @property
    def n_bfaces(self) -> int:
        return int(self.bface_area.shape[0])

    def without_boundary(self) -> "FaceMesh":
        e = np.zeros(0)
        return FaceMesh(self.n_cells, self.face_cell, self.face_area, self.face_dist, self.cell_vol,
                        np.zeros(0, np.int32), e, e.copy())


class FaceOp:
    """y = prefill(x) + dt * div grad x over a FaceMesh (see sb_oracle.h: orc_face_op)."""

    def __init__(self, mesh: FaceMesh, prefill: int, dt: float, dirichlet: bool = False):
        self.mesh = mesh
        self.prefill, self.dt = int(prefill), float(dt)
        self._keep = [np.ascontiguousarray(mesh.face_cell, np.int32).reshape(-1), _f64(mesh.face_area),
                      _f64(mesh.face_dist), _f64(mesh.cell_vol),
                      np.ascontiguousarray(mesh.bface_cell, np.int32), _f64(mesh.bface_area),
                      _f64(mesh.bface_dist)]
        k = self._keep
        nb = mesh.n_bfaces if dirichlet else 0
        self.struct = FaceOpStruct(mesh.n_cells, mesh.n_faces, _p(k[0], _i32p), _p(k[1], _f64p),
                                   _p(k[2], _f64p), _p(k[3], _f64p), nb, _p(k[4], _i32p),
                                   _p(k[5], _f64p), _p(k[6], _f64p), self.prefill, self.dt)
        self.n = mesh.n_cells

    # --- face loop ---------------------------------------------------------------------------
    def apply(self, x):
        x = _f64(x)
        y = np.empty(self.n)
        lib().orc_apply_faces(C.byref(self.struct), _p(x, _f64p), _p(y, _f64p))
        return y

    def divgrad_accumulate(self, dt, c, u):
        """u += dt * div grad c in place: stormDivGrad as the playground calls it (Playground.cpp:115-131)."""
        c = _f64(c)
        assert u.dtype == np.float64 and u.flags.c_contiguous
        lib().orc_divgrad_accumulate(C.byref(self.struct), float(dt), _p(c, _f64p), _p(u, _f64p))
        return u

    @property
    def callback(self):
        """(function pointer, user pointer) pair for orc_cg / ref_solve."""
        return C.cast(lib().orc_apply_faces_cb, C.c_void_p), C.cast(C.pointer(self.struct), C.c_void_p)

    # --- cell rows ---------------------------------------------------------------------------
    def rows(self, ld: int | None = None):
        L = lib()
        w = L.orc_rows_width(C.byref(self.struct))
        ld = self.n if ld is None else int(ld)
        col = np.empty((w, ld), np.int32)
        face = np.empty((w, ld), np.int64)
        L.orc_build_rows(C.byref(self.struct), w, ld, _p(col, _i32p), _p(face, _i64p))
        return w, ld, col, face

    def rows_faithful(self, ld: int | None = None):
        w, ld, col, face = self.rows(ld)
        g = np.empty((w, ld))
        d = np.empty((w, ld))
        lib().orc_rows_faithful(C.byref(self.struct), w, ld, _p(face, _i64p), _p(g, _f64p), _p(d, _f64p))
        return w, ld, col, g, d

    def rows_coef(self, ld: int | None = None):
        w, ld, col, face = self.rows(ld)
        col_out = np.empty((w, ld), np.int32)
        a = np.empty((w, ld))
        diag = np.empty(ld)
        lib().orc_rows_coef(C.byref(self.struct), w, ld, _p(col, _i32p), _p(face, _i64p),
                            _p(col_out, _i32p), _p(a, _f64p), _p(diag, _f64p))
        return w, ld, col_out, a, diag

    def apply_rows_faithful(self, x, rows=None):
        w, ld, col, g, d = rows or self.rows_faithful()
        x = _f64(x)
        y = np.empty(self.n)
        lib().orc_apply_rows_faithful(self.n, w, ld, _p(col, _i32p), _p(g, _f64p), _p(d, _f64p),
                                      self.prefill, self.dt, _p(x, _f64p), _p(y, _f64p))
        return y

    def apply_rows_coef(self, x, rows=None):
        w, ld, col, a, diag = rows or self.rows_coef()
        x = _f64(x)
        y = np.empty(self.n)
        lib().orc_apply_rows_coef(self.n, w, ld, _p(col, _i32p), _p(a, _f64p), _p(diag, _f64p),
                                  _p(x, _f64p), _p(y, _f64p))
        return y


class RowsOp:
    """An operator given as coefficient rows (col, a, diag): y_i = diag_i x_i + sum_k a_ik x[col_ik]."""

    def __init__(self, n, width, ld, col, a, diag):
        self.n = int(n)
        self._keep = [np.ascontiguousarray(col, np.int32), _f64(a), _f64(diag)]
        k = self._keep
        self.rows = (int(width), int(ld), k[0], k[1], k[2])
        self.struct = RowsOpStruct(self.n, int(width), int(ld), _p(k[0], _i32p), _p(k[1], _f64p), _p(k[2], _f64p))

    def apply(self, x):
        x = _f64(x)
        y = np.empty(self.n)
        w, ld, col, a, diag = self.rows
        lib().orc_apply_rows_coef(self.n, w, ld, _p(col, _i32p), _p(a, _f64p), _p(diag, _f64p), _p(x, _f64p),
                                  _p(y, _f64p))
        return y

    @property
    def callback(self):
        return C.cast(lib().orc_apply_rows_cb, C.c_void_p), C.cast(C.pointer(self.struct), C.c_void_p)


class ConvDiffOp:
    """y = -nu div grad x + div(beta x), first-order upwind, Dirichlet mirror ghosts (sb_oracle.h:
    orc_convdiff_op). `face_un` / `bface_un`: beta . n per interior / boundary face."""

    def __init__(self, mesh: FaceMesh, nu: float, face_un, bface_un, dirichlet: bool = True):
        self.mesh, self.nu, self.n = mesh, float(nu), mesh.n_cells
        self._geo = FaceOp(mesh, prefill=0, dt=0.0, dirichlet=dirichlet)
        self._un = [_f64(face_un), _f64(bface_un)]
        assert self._un[0].shape[0] == mesh.n_faces and self._un[1].shape[0] == mesh.n_bfaces
        self.struct = ConvDiffStruct(self._geo.struct, self.nu, _p(self._un[0], _f64p), _p(self._un[1], _f64p))

    def apply(self, x):
        x = _f64(x)
        y = np.empty(self.n)
        lib().orc_apply_convdiff_faces(C.byref(self.struct), _p(x, _f64p), _p(y, _f64p))
        return y

    @property
    def callback(self):
        return C.cast(lib().orc_apply_convdiff_faces_cb, C.c_void_p), C.cast(C.pointer(self.struct), C.c_void_p)

    def rows_coef(self, ld: int | None = None) -> RowsOp:
        """Row form in the product's layout contract (width from the interior faces only)."""
        L = lib()
        interior = FaceOpStruct.from_buffer_copy(self._geo.struct)
        interior.n_bfaces = 0
        w = max(1, L.orc_rows_width(C.byref(interior)))
        ld = self.n if ld is None else int(ld)
        col, a, diag = np.empty((w, ld), np.int32), np.empty((w, ld)), np.empty(ld)
        L.orc_rows_convdiff(C.byref(self.struct), w, ld, _p(col, _i32p), _p(a, _f64p), _p(diag, _f64p))
        return RowsOp(self.n, w, ld, col, a, diag)


class CallbackOp:
    """Wrap a Python callable y = f(x) as an (fn, user) pair (slow; small cases only)."""

    def __init__(self, f, n):
        self.n = n

        def _cb(user, y, x, n_):
            xv = np.ctypeslib.as_array(x, shape=(n_,))
            yv = np.ctypeslib.as_array(y, shape=(n_,))
            yv[:] = f(xv.copy())

        self._cfn = APPLY_FN(_cb)

    @property
    def callback(self):
        return C.cast(self._cfn, C.c_void_p), C.c_void_p(None)


def set_segments(seg_ptr) -> None:
    """Rank-block boundaries for RED_TREE_SEG (the multi-GPU reduction order)."""
    seg_ptr = np.ascontiguousarray(seg_ptr, np.int64)
    lib().orc_set_segments(len(seg_ptr) - 1, _p(seg_ptr, _i64p))
    if have_ref():  # the compiled reference harness links its own copy of the oracle reductions
        ref().orc_set_segments.argtypes = [C.c_int, _i64p]
        ref().orc_set_segments(len(seg_ptr) - 1, _p(seg_ptr, _i64p))


def dot(a, b, mode=RED_SEQ) -> float:
    a, b = _f64(a), _f64(b)
    return float(lib().orc_dot(a.shape[0], _p(a, _f64p), _p(b, _f64p), mode))


def norm2(a, mode=RED_SEQ) -> float:
    a = _f64(a)
    return float(lib().orc_norm2(a.shape[0], _p(a, _f64p), mode))


@dataclass
class SolveResult:
    x: np.ndarray
    converged: bool
    iterations: int
    abs_err: float
    rel_err: float
    hist: np.ndarray    # hist[0] initial residual, hist[k] after iteration k
    trace: np.ndarray   # every dot/norm result in call order
    n_apply: int = -1


def solve(solver: str, op, b, x0=None, num_iterations=2000, abs_tol=1e-6, rel_tol=1e-6,
          mode=RED_SEQ) -> SolveResult:
    """Plain-C restated solvers (oracle/sb_oracle.c): 'cg' | 'bicgstab'."""
    L = lib()
    fn = {"cg": L.orc_cg, "bicgstab": L.orc_bicgstab}[solver]
    b = _f64(b)
    n = b.shape[0]
    x = np.zeros(n) if x0 is None else _f64(x0).copy()
    cap_h = num_iterations + 2
    cap_t = 8 * num_iterations + 16
    hist, trace = np.zeros(cap_h), np.zeros(cap_t)
    opts = SolverOpts(num_iterations, abs_tol, rel_tol, mode)
    rep = SolverReport()
    f, u = op.callback
    rc = fn(f, u, n, _p(b, _f64p), _p(x, _f64p), C.byref(opts), C.byref(rep), _p(hist, _f64p), cap_h,
            _p(trace, _f64p), cap_t)
    assert rc == 0
    return SolveResult(x, bool(rep.converged), rep.iterations, rep.abs_err, rep.rel_err,
                       hist[:rep.n_hist].copy(), trace[:min(rep.n_trace, cap_t)].copy())


REF_SOLVERS = ("cg", "cgs", "bicgstab", "bicgstabl", "gmres", "fgmres", "tfqmr", "tfqmr1", "idrs",
               "richardson")
REF_NONLINEAR = ("jfnk",)


class JacobiOp:
    """y = x / diag, the point-diagonal preconditioner (restates sb_op_jacobi): a CallbackOp for ref_solve(pre=...)."""

    def __init__(self, diag):
        d = _f64(diag).copy()
        self.n = d.shape[0]
        self._cb = CallbackOp(lambda x: x / d, self.n)

    @property
    def callback(self):
        return self._cb.callback


def ref_solve(solver: str, op, b, x0=None, num_iterations=2000, abs_tol=1e-6, rel_tol=1e-6,
              num_inner=0, mode=RED_SEQ, relaxation_factor=0.0, reset_rng=True,
              trace_cap=None, pre=None, pre_side="right", cheb_degree=0, cheb_power_iterations=0,
              cheb_eig_ratio=0.0) -> SolveResult:
    """The reference's own solver headers (oracle/_ref), on a host vector. `pre`: optional operator object
    with a .callback (e.g. JacobiOp) placed in the reference's pre_op slot, or the string "chebyshev": the product's
    Storm::ChebyshevPreconditioner template compiled on the host vector (parameters 0 = class defaults);
    pre_side: left | right | symmetric."""
    R = ref()
    if reset_rng:
        R.ref_reset_rng()
    b = _f64(b)
    n = b.shape[0]
    x = np.zeros(n) if x0 is None else _f64(x0).copy()
    cap_h = num_iterations + 2
    cap_t = trace_cap or (64 * num_iterations + 256)
    hist, trace = np.zeros(cap_h), np.zeros(cap_t)
    cheb = isinstance(pre, str) and pre == "chebyshev"
    pf, pu = pre.callback if (pre is not None and not cheb) else (None, None)
    opts = RefOpts(num_iterations, abs_tol, rel_tol, num_inner, mode, relaxation_factor, pf, pu,
                   {"left": 0, "right": 1, "symmetric": 2}[pre_side], 3 if cheb else 0, int(cheb_degree),
                   int(cheb_power_iterations), float(cheb_eig_ratio))
    rep = RefReport()
    f, u = op.callback
    rc = R.ref_solve(solver.encode(), n, f, u, _p(b, _f64p), _p(x, _f64p), C.byref(opts), C.byref(rep),
                     _p(hist, _f64p), cap_h, _p(trace, _f64p), cap_t)
    if rc != 0:
        raise ValueError(f"ref_solve: unknown solver {solver!r}")
    return SolveResult(x, bool(rep.converged), rep.iterations, rep.abs_err, rep.rel_err,
                       hist[:rep.n_hist].copy(), trace[:min(rep.n_trace, cap_t)].copy(), rep.n_apply)


def ref_solve_non_uniform(solver: str, op, b, shift, x0=None, num_iterations=2000, abs_tol=1e-6, rel_tol=1e-6,
                          mode=RED_SEQ) -> SolveResult:
    """The reference's solve_non_uniform (Solver.hpp:271-292) on a host vector for A(x) = op(x) + shift."""
    R = ref()
    R.ref_solve_non_uniform.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_void_p, _f64p, _f64p, _f64p,
                                        C.POINTER(RefOpts), C.POINTER(RefReport), _f64p, C.c_int64]
    R.ref_solve_non_uniform.restype = C.c_int
    b = _f64(b)
    shift = None if shift is None else _f64(shift)     # None: `op` is the affine operator itself
    n = b.shape[0]
    x = np.zeros(n) if x0 is None else _f64(x0).copy()
    cap_t = 64 * num_iterations + 256
    trace = np.zeros(cap_t)
    opts = RefOpts(num_iterations, abs_tol, rel_tol, 0, mode, 0.0, None, None, 1, 0, 0, 0, 0.0)
    rep = RefReport()
    f, u = op.callback
    rc = R.ref_solve_non_uniform(solver.encode(), n, f, u, _p(b, _f64p), None if shift is None else _p(shift, _f64p),
                                 _p(x, _f64p), C.byref(opts),
                                 C.byref(rep), _p(trace, _f64p), cap_t)
    if rc != 0:
        raise ValueError(f"ref_solve_non_uniform: unknown solver {solver!r}")
    return SolveResult(x, bool(rep.converged), rep.iterations, rep.abs_err, rep.rel_err, np.zeros(0),
                       trace[:min(rep.n_trace, cap_t)].copy(), rep.n_apply)


# ---- the playground's Cahn-Hilliard time step (Playground.cpp:133-175) ------------------------------------------
CH_TAU, CH_GAMMA, CH_SIGMA = 1.0e-3, 1.0e-4, 2.0   # Playground.cpp:113


def ch_dF_dc(c):
    """dF/dc of the double-well potential, in the reference's association order (Playground.cpp:142-144):
    ((2.0 * c) * (c - 1.0)) * (2.0 * c - 1.0), every operation rounded separately."""
    return ((2.0 * c) * (c - 1.0)) * (2.0 * c - 1.0)


class CahnHilliardOp:
    """The affine operator the playground hands to CG (Playground.cpp:153-167):
        w_hat <<= f + sigma * (c_in - c);  stormDivGrad(mesh, w_hat, -Gamma, c_in);
        c_hat <<= c_in;                    stormDivGrad(mesh, c_hat, -tau, w_hat)
    over the interior faces of `mesh` (a FaceMesh). `w_hat` keeps the last evaluation, like the reference's."""

    def __init__(self, mesh: FaceMesh, c, tau=CH_TAU, Gamma=CH_GAMMA, sigma=CH_SIGMA):
        self.faces = FaceOp(mesh.without_boundary(), prefill=0, dt=0.0)
        self.n = mesh.n_cells
        self.c = _f64(c).copy()
        self.f = ch_dF_dc(self.c)
        self.tau, self.Gamma, self.sigma = float(tau), float(Gamma), float(sigma)
        self.w_hat = np.zeros(self.n)
        self._cb = CallbackOp(self.apply, self.n)

    def apply(self, c_in):
        c_in = _f64(c_in)
        self.w_hat = self.f + self.sigma * (c_in - self.c)
        self.faces.divgrad_accumulate(-self.Gamma, c_in, self.w_hat)
        c_hat = c_in.copy()
        self.faces.divgrad_accumulate(-self.tau, self.w_hat, c_hat)
        return c_hat

    @property
    def callback(self):
        return self._cb.callback


def cahn_hilliard_step(mesh: FaceMesh, c, mode=RED_SEQ, num_iterations=2000, abs_tol=1e-6, rel_tol=1e-6,
                       uniformed=False, **constants) -> SolveResult:
    """One time step: `c_hat <<= c; solve<CgSolver>(c_hat, c, op)` (Playground.cpp:148-167; the defaults are those
    of IterativeSolver, Solver.hpp:67-72). Uses the reference's own CgSolver when oracle/_ref is built, the C
    restatement otherwise (the two are pinned against each other in tests/test_oracle_golden.py)."""
    op = CahnHilliardOp(mesh, c, **constants)
    if uniformed:   # the same step through the reference's solve_non_uniform (the operator is affine): CG converges
        return ref_solve_non_uniform("cg", op, op.c, None, x0=op.c, num_iterations=num_iterations, abs_tol=abs_tol,
                                     rel_tol=rel_tol, mode=mode)
    run = ref_solve if have_ref() else solve
    return run("cg", op, op.c, x0=op.c, num_iterations=num_iterations, abs_tol=abs_tol, rel_tol=rel_tol, mode=mode)


# ---- reading the binary dumps of oracle/_ref/ref_mesh_tool ------------------------------------
class _Reader:
    def __init__(self, path):
        with open(path, "rb") as f:
            self.buf = f.read()
        self.off = 0

    def i64(self):
        v = struct.unpack_from("<q", self.buf, self.off)[0]
        self.off += 8
        return v

    def f64(self):
        v = struct.unpack_from("<d", self.buf, self.off)[0]
        self.off += 8
        return v

    def arr(self, dtype):
        n = self.i64()
        a = np.frombuffer(self.buf, dtype=dtype, count=n, offset=self.off).copy()
        self.off += n * np.dtype(dtype).itemsize
        return a


def read_mesh_export(path):
    r = _Reader(path)
    n_cells, n_nodes, n_faces_total, n_labels = r.i64(), r.i64(), r.i64(), r.i64()
    face_cell = r.arr(np.int32).reshape(-1, 2)
    face_area, face_dist = r.arr(np.float64), r.arr(np.float64)
    bface_cell, bface_area, bface_dist, bface_label = (r.arr(np.int32), r.arr(np.float64),
                                                       r.arr(np.float64), r.arr(np.int32))
    cell_vol, cx, cy = r.arr(np.float64), r.arr(np.float64), r.arr(np.float64)
    mesh = FaceMesh(n_cells, face_cell, face_area, face_dist, cell_vol, bface_cell, bface_area, bface_dist)
    extra = dict(n_nodes=n_nodes, n_faces_total=n_faces_total, n_face_labels=n_labels,
                 bface_label=bface_label, cell_center=np.stack([cx, cy], 1))
    return mesh, extra


def read_cg_dump(path):
    r = _Reader(path)
    n, conv, it = r.i64(), r.i64(), r.i64()
    abs_err, rel_err = r.f64(), r.f64()
    b, x, hist = r.arr(np.float64), r.arr(np.float64), r.arr(np.float64)
    return dict(n=n, converged=bool(conv), iterations=it, abs_err=abs_err, rel_err=rel_err, b=b, x=x,
                hist=hist)


def read_ch_dump(path):
    """Output of `ref_mesh_tool ch`: initial c and, per time step, the new c with the CG solver's report."""
    r = _Reader(path)
    n, steps = r.i64(), r.i64()
    out = dict(n=n, c0=r.arr(np.float64), steps=[])
    for _ in range(steps):
        conv, it = r.i64(), r.i64()
        abs_err, rel_err = r.f64(), r.f64()
        c, hist = r.arr(np.float64), r.arr(np.float64)
        out["steps"].append(dict(converged=bool(conv), iterations=it, abs_err=abs_err, rel_err=rel_err, c=c,
                                 hist=hist))
    return out
