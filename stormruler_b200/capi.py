"""ctypes binding of libstormb200.so -- one declaration per symbol of include/stormb200.h.

There is no CPU fallback: if the library is missing, `load()` raises; if there is no CUDA device,
`Context()` raises with the library's own error string.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libstormb200.so")

SB_OK = 0
FORM_FAITHFUL, FORM_COEF = 0, 1
PART_METIS, PART_SLAB = 0, 1
COMM_NCCL, COMM_P2P = 0, 1
SCHEDULE_AUTO, SCHEDULE_STEPWISE, SCHEDULE_PERSISTENT, SCHEDULE_FOLDED = 0, 1, 2, 3
TIMELINE_WORDS = 20
TUNE_PUSH_ON_PRODUCE, TUNE_NO_ACK, TUNE_STREAM_OPERATOR, TUNE_PDL_FINAL, TUNE_PDL_AFTER_FINAL, TUNE_PDL_APPLY, TUNE_IN_KERNEL_REDUCER, TUNE_OFF = 1, 2, 4, 8, 16, 32, 64, 0x80000000
TUNE_PUSH_LAZY = 128
COMM_BLOB_BYTES = 256
ASSIGN, ADD_ASSIGN, SUB_ASSIGN, MUL_ASSIGN, DIV_ASSIGN = range(5)
OP_VEC0, OP_SCAL0 = 0, 8
OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_NEG = 16, 17, 18, 19, 20
EXPR_MAX_OPS, EXPR_MAX_VEC, EXPR_MAX_SCAL = 24, 4, 4

f64p = C.POINTER(C.c_double)
i32p = C.POINTER(C.c_int32)
vpp = C.POINTER(C.c_void_p)


class MeshSoa(C.Structure):
    _fields_ = [("n_cells", C.c_int64), ("n_faces", C.c_int64), ("face_cell", i32p),
                ("face_area", f64p), ("face_dist", f64p), ("cell_vol", f64p),
                ("n_bfaces", C.c_int64), ("bface_cell", i32p), ("bface_area", f64p),
                ("bface_dist", f64p)]


i64p = C.POINTER(C.c_int64)


class PartInfo(C.Structure):
    _fields_ = [("n_parts", C.c_int32), ("n_cells", C.c_int64), ("edge_cut", C.c_int64),
                ("max_owned", C.c_int64), ("min_owned", C.c_int64), ("max_halo", C.c_int64),
                ("vec_capacity", C.c_int64)]


class LocalMesh(C.Structure):
    _fields_ = [("rank", C.c_int32), ("n_parts", C.c_int32), ("n_owned", C.c_int64),
                ("n_interior", C.c_int64), ("n_halo", C.c_int64), ("halo_base", C.c_int64),
                ("local_to_global", i32p), ("soa", MeshSoa), ("face_global", i64p),
                ("n_nbr", C.c_int32), ("nbr_rank", i32p), ("send_ptr", i64p), ("send_idx", i32p),
                ("recv_ptr", i64p), ("send_dst", i64p), ("bface_global", i64p)]


class OpDesc(C.Structure):
    _fields_ = [("form", C.c_int32), ("prefill", C.c_int32), ("dt", C.c_double)]


class ConvDiffDesc(C.Structure):
    _fields_ = [("nu", C.c_double), ("face_un", f64p), ("bface_un", f64p)]


class OpInfo(C.Structure):
    _fields_ = [("n_cells", C.c_int64), ("n_entries", C.c_int64), ("width", C.c_int32),
                ("ld", C.c_int64), ("form", C.c_int32), ("device_bytes", C.c_int64),
                ("algorithmic_bytes_per_apply", C.c_int64)]


class Expr(C.Structure):
    _fields_ = [("n_ops", C.c_int32), ("ops", C.c_uint8 * EXPR_MAX_OPS),
                ("vec", C.c_void_p * EXPR_MAX_VEC), ("scal", C.c_double * EXPR_MAX_SCAL)]


GROUP_MAX_STMT, GROUP_MAX_TERMS, GROUP_MAX_DOTS = 8, 8, 8


class Chain(C.Structure):      # sb_chain: y = ((base +- c0*x0) +- c1*x1) ...
    _fields_ = [("y", C.c_void_p), ("base", C.c_void_p), ("n_terms", C.c_int32),
                ("x", C.c_void_p * GROUP_MAX_TERMS), ("c", C.c_double * GROUP_MAX_TERMS),
                ("sub", C.c_uint8 * GROUP_MAX_TERMS)]


class SolverOpts(C.Structure):
    _fields_ = [("num_iterations", C.c_int64), ("abs_tol", C.c_double), ("rel_tol", C.c_double),
                ("check_every", C.c_int32), ("use_graph", C.c_int32), ("profile", C.c_int32),
                ("schedule", C.c_int32), ("timeline_iters", C.c_int32), ("h_timeline", C.POINTER(C.c_uint64)),
                ("tuning", C.c_uint32)]


class GmresOpts(C.Structure):
    _fields_ = [("num_iterations", C.c_int64), ("abs_tol", C.c_double), ("rel_tol", C.c_double),
                ("num_inner_iterations", C.c_int32), ("lookahead", C.c_int32)]


class SolverReport(C.Structure):
    _fields_ = [("converged", C.c_int32), ("iterations", C.c_int64), ("initial_err", C.c_double),
                ("abs_err", C.c_double), ("rel_err", C.c_double), ("n_hist", C.c_int64),
                ("n_trace", C.c_int64), ("solve_ms", C.c_double), ("iter_ms", C.c_double),
                ("launches", C.c_int64), ("n_kernel_slots", C.c_int32), ("kernel_ms", C.c_double * 8),
                ("schedule", C.c_int32), ("wait_ms", C.c_double * 8), ("ar_wait_ms", C.c_double * 8),
                ("final_ms", C.c_double * 8)]


# name -> (restype, argtypes); the keys are exactly the SB_API symbols of include/stormb200.h
SIGNATURES = {
    "sb_last_error": (C.c_char_p, []),
    "sb_version": (C.c_int, []),
    "sb_ctx_create": (C.c_int, [C.c_int, vpp]),
    "sb_ctx_destroy": (C.c_int, [C.c_void_p]),
    "sb_sync": (C.c_int, [C.c_void_p]),
    "sb_ctx_stream": (C.c_void_p, [C.c_void_p]),
    "sb_ctx_launch_count": (C.c_int64, [C.c_void_p]),
    "sb_vec_alloc": (C.c_int, [C.c_void_p, C.c_size_t, vpp]),
    "sb_vec_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sb_vec_upload": (C.c_int, [C.c_void_p, C.c_void_p, f64p, C.c_size_t]),
    "sb_vec_download": (C.c_int, [C.c_void_p, C.c_void_p, f64p, C.c_size_t]),
    "sb_op_create": (C.c_int, [C.c_void_p, C.POINTER(MeshSoa), C.POINTER(OpDesc), vpp]),
    "sb_op_create_convdiff": (C.c_int, [C.c_void_p, C.POINTER(MeshSoa), C.POINTER(ConvDiffDesc), vpp]),
    "sb_dist_op_create_convdiff": (C.c_int, [C.c_void_p, C.POINTER(LocalMesh), C.POINTER(ConvDiffDesc), vpp]),
    "sb_op_destroy": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sb_op_get_info": (C.c_int, [C.c_void_p, C.POINTER(OpInfo)]),
    "sb_op_download_rows": (C.c_int, [C.c_void_p, C.c_void_p, i32p, f64p, f64p, f64p]),
    "sb_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sb_apply_dot": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, f64p]),
    "sb_apply_dot_yy_yx": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, f64p]),
    "sb_apply_accumulate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]),
    "sb_op_jacobi": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sb_mesh_generate_box": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_uint64, C.c_int,
                                       C.c_uint64, vpp]),
    "sb_mesh_from_cells": (C.c_int, [C.c_int, C.c_int64, f64p, C.c_int64, i32p, vpp]),
    "sb_mesh_from_faces": (C.c_int, [C.POINTER(MeshSoa), f64p, f64p, f64p, vpp]),
    "sb_mesh_read_tetgen": (C.c_int, [C.c_char_p, vpp]),
    "sb_mesh_read_tetgen_2d": (C.c_int, [C.c_char_p, vpp]),
    "sb_mesh_bface_labels": (C.c_int, [C.c_void_p, i32p]),
    "sb_mesh_destroy": (C.c_int, [C.c_void_p]),
    "sb_mesh_renumber_rcm": (C.c_int, [C.c_void_p, i32p]),
    "sb_mesh_permute_cells": (C.c_int, [C.c_void_p, i32p]),
    "sb_mesh_get_soa": (C.c_int, [C.c_void_p, C.POINTER(MeshSoa)]),
    "sb_mesh_cell_centers": (C.c_int, [C.c_void_p, f64p]),
    "sb_mesh_face_normals": (C.c_int, [C.c_void_p, f64p, f64p]),
    "sb_mesh_write_vtk": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(f64p)]),
    "sb_mesh_bandwidth": (C.c_int64, [C.c_void_p]),
    "sb_part_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, vpp]),
    "sb_part_from_array": (C.c_int, [C.c_void_p, C.c_int, i32p, vpp]),
    "sb_part_destroy": (C.c_int, [C.c_void_p]),
    "sb_part_get_info": (C.c_int, [C.c_void_p, C.POINTER(PartInfo)]),
    "sb_part_get_array": (C.c_int, [C.c_void_p, C.POINTER(i32p)]),
    "sb_part_local": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(LocalMesh)]),
    "sb_comm_prepare": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int32, C.c_void_p]),
    "sb_comm_connect": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sb_comm_destroy": (C.c_int, [C.c_void_p]),
    "sb_comm_status": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "sb_dist_op_create": (C.c_int, [C.c_void_p, C.POINTER(LocalMesh), C.POINTER(OpDesc), vpp]),
    "sb_eval": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.POINTER(Expr)]),
    "sb_eval_group": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(Chain), C.c_int, vpp, vpp, f64p]),
    "sb_fill": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_double]),
    "sb_copy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "sb_dot": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, f64p]),
    "sb_norm2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, f64p]),
    "sb_dot_batch": (C.c_int, [C.c_void_p, C.c_int, vpp, vpp, C.c_size_t, f64p]),
    "sb_cg_solve": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(SolverOpts),
                              C.POINTER(SolverReport), f64p, C.c_int64, f64p, C.c_int64]),
    "sb_bicgstab_solve": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(SolverOpts),
                                    C.POINTER(SolverReport), f64p, C.c_int64, f64p, C.c_int64]),
    "sb_gmres_solve": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(GmresOpts),
                                 C.POINTER(SolverReport), f64p, C.c_int64, f64p, C.c_int64]),
    "sb_solve_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_char_p, f64p, f64p, C.POINTER(SolverOpts),
                                C.POINTER(SolverReport), f64p, C.c_int64]),
}

_lib = None


def load():
    """Load libstormb200.so and declare every entry point. Raises if the library is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m stormruler_b200.build` "
                "(there is no CPU fallback for the Krylov path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def use_host_library(path: str):
    """Bind the HOST-ONLY entry points (sb_mesh_*, sb_part_*, sb_last_error) to another shared library that exports
    them -- oracle/libsb_meshprep.so, the product's mesh sources built without any device code. bench.py's reference
    arm prepares its inputs this way, so that the process timing the reference's CPU solver never maps the CUDA
    library. Must be called before anything else has loaded the library; the device entry points are then absent."""
    global _lib
    if _lib is not None:
        raise RuntimeError("use_host_library: a library is already loaded")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        if name.startswith(("sb_mesh_", "sb_part_")) or name == "sb_last_error":
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


class StormB200Error(RuntimeError):
    pass


def check(rc: int) -> None:
    if rc != SB_OK:
        msg = load().sb_last_error().decode(errors="replace")
        raise StormB200Error(f"stormb200 error {rc}: {msg}")
