"""Multi-GPU plumbing: one process per GPU (DESIGN.md "Multi-GPU").

What lives here is host-side glue over the C ABI: the `torch.distributed` rendezvous (blob all-gather,
partition broadcast, barriers, max-over-ranks timing), `DistContext` (a `Context` whose vectors come
from the symmetric pool of `sb_comm_prepare`), `DistOperator` (`sb_dist_op_create`) and the N>1 leg of
`bench.py`. The data path itself -- halo exchange over NVLink, the in-kernel all-reduce, or their NCCL
equivalents -- is inside libstormb200.so; nothing here touches vector data except scatter/gather of
test and benchmark inputs.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import sys
import time

import numpy as np

from . import Context, DeviceVector, BiCgStabSolver, CgSolver, FORM_COEF, solve_host
from . import capi
from .mesh import LOCAL_ARRAYS, LOCAL_SCALARS, LocalArrays, LocalView, Mesh, Partition


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ---- torch.distributed plumbing ---------------------------------------------------------------------
def local_device() -> int:
    """CUDA device of this rank: LOCAL_RANK modulo the number of visible devices. With more ranks than devices the
    ranks share GPUs (their kernels are time-sliced): slow, but the peer-memory protocol is exactly the one that
    runs with a GPU per rank, so a 1-GPU box can run the world-size-2 bit-exactness tests."""
    import torch
    return int(os.environ.get("LOCAL_RANK", "0")) % max(1, torch.cuda.device_count())


def oversubscribed() -> bool:
    import torch
    return int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1"))) > max(1, torch.cuda.device_count())


def init_process_group(cuda: bool):
    """RANK / WORLD_SIZE / MASTER_* come from torchrun. CPU tensors travel over gloo (blobs, the
    partition array, timings); NCCL is initialised too on a GPU box so the launch contract's backend is
    the one in use for device-side barriers."""
    import torch
    import torch.distributed as dist
    if dist.is_initialized():
        return dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    if cuda:
        torch.cuda.set_device(local_device())
        # several ranks on one device (protocol tests on a 1-GPU box): NCCL refuses duplicate devices, gloo carries
        # the host-side plumbing alone (the data path never goes through torch.distributed anyway)
        dist.init_process_group(backend="gloo" if oversubscribed() else "cpu:gloo,cuda:nccl")
    else:
        dist.init_process_group(backend="gloo")
    return dist


def broadcast_array(arr: np.ndarray | None, shape, dtype, src: int = 0) -> np.ndarray:
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(arr, dtype)) if dist.get_rank() == src else torch.empty(shape, dtype=getattr(torch, np.dtype(dtype).name))
    dist.broadcast(t, src=src)
    return t.numpy()


def partition_mesh(mesh: Mesh, world: int, method: int = capi.PART_METIS) -> Partition:
    """Rank 0 partitions (METIS k-way or RCM slabs), everyone receives the same `part` array and builds
    its own local mesh from it: one source of truth, whatever the library's determinism."""
    import torch.distributed as dist
    part = None
    if dist.get_rank() == 0:
        t = time.time()
        p0 = Partition(mesh, world, method)
        part = np.array(p0.part, dtype=np.int32, copy=True)
        log(f"[multigpu] partition ({'metis' if method == capi.PART_METIS else 'slab'}, {world} parts): "
            f"{time.time() - t:.1f}s, edge cut {p0.info.edge_cut}, owned {p0.info.min_owned}..{p0.info.max_owned}, "
            f"max halo {p0.info.max_halo}")
        del p0
    part = broadcast_array(part, (mesh.n_cells,), np.int32)
    return Partition(mesh, world, part=part)


def scatter_mesh(build_mesh, world: int, method: int = capi.PART_METIS, cell_fields=None):
    """Rank 0 builds the global mesh (`build_mesh()` -> Mesh), partitions it and ships every rank its local mesh;
    the other ranks never hold the global mesh (at 49.8 M hexahedra the global build peaks at 29 GB per process:
    with partition_mesh() every rank pays that, here only rank 0 does). `cell_fields(mesh)` -> {name: per-cell array}
    is evaluated on rank 0 and each rank receives the rows of its owned cells. Returns (local, info, fields) with
    `local` a LocalArrays (same surface as LocalView), `info` the partition summary as a dict (n_cells, edge_cut,
    min_owned, max_owned, max_halo, vec_capacity), `fields` this rank's slices; rank 0 additionally finds the global
    mesh under info["mesh"] (None elsewhere)."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank()
    info_keys = ("n_cells", "edge_cut", "min_owned", "max_owned", "max_halo", "vec_capacity")

    def to_tensor(a):
        return torch.from_numpy(np.ascontiguousarray(a))

    if rank == 0:
        t = time.time()
        mesh = build_mesh()
        part = Partition(mesh, world, method)
        fields_g = cell_fields(mesh) if cell_fields is not None else {}
        info = {k: int(getattr(part.info, k)) for k in info_keys}
        log(f"[multigpu] rank 0 built and partitioned the global mesh ({mesh.n_cells} cells, {world} parts) in "
            f"{time.time() - t:.1f}s, edge cut {info['edge_cut']}, owned {info['min_owned']}..{info['max_owned']}")
        mine, my_fields = None, {}
        for r in range(world):
            L = LocalArrays.from_view(part.local(r))
            fl = {k: np.ascontiguousarray(np.asarray(v)[L.owned_global]) for k, v in fields_g.items()}
            if r == 0:
                mine, my_fields = L, fl
                continue
            header = {"scalars": {k: getattr(L, k) for k in LOCAL_SCALARS},
                      "lengths": {k: int(L._keep[k].size) for k, _ in LOCAL_ARRAYS},
                      "fields": {k: (list(v.shape), str(v.dtype)) for k, v in fl.items()}, "info": info}
            blob = np.frombuffer(json.dumps(header).encode(), np.uint8).copy()   # CPU tensors: travel over gloo
            dist.send(torch.tensor([blob.size], dtype=torch.int64), dst=r)
            dist.send(torch.from_numpy(blob), dst=r)
            for k, _ in LOCAL_ARRAYS:
                if L._keep[k].size:
                    dist.send(to_tensor(L._keep[k]), dst=r)
            for k, v in fl.items():
                if v.size:
                    dist.send(to_tensor(v), dst=r)
            del L, fl
        del part
        info["mesh"] = mesh
        return mine, info, my_fields
    size = torch.zeros(1, dtype=torch.int64)
    dist.recv(size, src=0)
    blob = torch.empty(int(size[0]), dtype=torch.uint8)
    dist.recv(blob, src=0)
    header = json.loads(blob.numpy().tobytes().decode())
    arrays = {}
    for k, dt in LOCAL_ARRAYS:
        a = np.empty(header["lengths"][k], dt)
        if a.size:
            dist.recv(torch.from_numpy(a), src=0)
        arrays[k] = a
    fields = {}
    for k, (shape, dtype) in header["fields"].items():
        a = np.empty(tuple(shape), np.dtype(dtype))
        if a.size:
            dist.recv(torch.from_numpy(a), src=0)
        fields[k] = a
    info = dict(header["info"])
    info["mesh"] = None
    return LocalArrays(header["scalars"], arrays), info, fields


class DistContext(Context):
    """A Context with a communicator: vectors are blocks of the symmetric slab (same offset on every
    rank), so kernels can address the neighbours' copies directly."""

    def __init__(self, device: int, rank: int, world: int, vec_capacity: int, n_vectors: int,
                 mode: int = capi.COMM_P2P):
        import torch
        import torch.distributed as dist
        super().__init__(device)
        self.rank, self.world, self.mode = rank, world, mode
        blob = (C.c_ubyte * capi.COMM_BLOB_BYTES)()
        capi.check(self.lib.sb_comm_prepare(self.handle, rank, world, mode, int(vec_capacity), int(n_vectors), blob))
        mine = torch.frombuffer(bytearray(bytes(blob)), dtype=torch.uint8).clone()
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        allb = b"".join(bytes(t.numpy().tobytes()) for t in gathered)
        buf = (C.c_ubyte * len(allb)).from_buffer_copy(allb)
        capi.check(self.lib.sb_comm_connect(self.handle, buf))
        dist.barrier()   # every slab is mapped and initialised before the first peer store

    def status(self) -> int:
        e = C.c_uint64()
        capi.check(self.lib.sb_comm_status(self.handle, C.byref(e)))
        return int(e.value)

    def close(self):
        if getattr(self, "handle", None):
            try:
                import torch.distributed as dist
                self.sync()
                if dist.is_initialized():
                    dist.barrier()   # no peer may still be storing into my slab
            except Exception:
                pass
            self.lib.sb_comm_destroy(self.handle)
        super().close()


class DistOperator:
    """This rank's rows of the distributed operator (sb_dist_op_create). Same surface as FvmOperator."""

    def __init__(self, ctx: DistContext, local: LocalView, prefill: int, dt: float, form: int = FORM_COEF,
                 dirichlet: bool = True):
        self.ctx, self.local = ctx, local
        lm = local.struct
        if not dirichlet:
            lm = capi.LocalMesh.from_buffer_copy(lm)
            lm.soa.n_bfaces = 0
        desc = capi.OpDesc(int(form), int(prefill), float(dt))
        h = C.c_void_p()
        capi.check(ctx.lib.sb_dist_op_create(ctx.handle, C.byref(lm), C.byref(desc), C.byref(h)))
        self.handle = h
        info = capi.OpInfo()
        capi.check(ctx.lib.sb_op_get_info(h, C.byref(info)))
        self.info, self.n = info, int(info.n_cells)

    def __del__(self):
        try:
            if self.handle and self.ctx.handle:
                self.ctx.lib.sb_op_destroy(self.ctx.handle, self.handle)
        except Exception:
            pass

    def mul(self, y: DeviceVector, x: DeviceVector):
        capi.check(self.ctx.lib.sb_apply(self.ctx.handle, self.handle, x.ptr, y.ptr))


class DistConvDiffOperator(DistOperator):
    """This rank's rows of the convection-diffusion operator (sb_dist_op_create_convdiff). `face_un` / `bface_un`
    are GLOBAL per-face arrays (Mesh.face_flux); the local ones are gathered through face_global / bface_global."""

    def __init__(self, ctx: DistContext, local: LocalView, nu: float, face_un, bface_un):
        self.ctx, self.local = ctx, local
        fu = np.ascontiguousarray(np.asarray(face_un, np.float64)[local.face_global])
        bu = np.ascontiguousarray(np.asarray(bface_un, np.float64)[local.bface_global])
        desc = capi.ConvDiffDesc(float(nu), fu.ctypes.data_as(capi.f64p), bu.ctypes.data_as(capi.f64p))
        h = C.c_void_p()
        capi.check(ctx.lib.sb_dist_op_create_convdiff(ctx.handle, C.byref(local.struct), C.byref(desc), C.byref(h)))
        self.handle = h
        info = capi.OpInfo()
        capi.check(ctx.lib.sb_op_get_info(h, C.byref(info)))
        self.info, self.n = info, int(info.n_cells)


def gather_global(local, x_local: np.ndarray, n_global: int) -> np.ndarray:
    """All ranks' owned values -> the global vector (on every rank). Test / reporting helper."""
    import torch
    import torch.distributed as dist
    out = torch.zeros(n_global, dtype=torch.float64)
    out[torch.from_numpy(np.asarray(local.owned_global, dtype=np.int64))] = torch.from_numpy(np.ascontiguousarray(x_local))
    dist.all_reduce(out)   # disjoint supports: the sum is an exact scatter
    return out.numpy()


def max_over_ranks(v: float) -> float:
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(v)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def sum_over_ranks(v: float) -> float:
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(v)], dtype=torch.float64)
    dist.all_reduce(t)
    return float(t[0])


# ---- bench.py, N > 1 -----------------------------------------------------------------------------------
def bench_main(args, build_problem, workload_config, peaks, ClockSampler):
    """Strong scaling of the N=1 workload: the same 10.1 M-cell problem, METIS-partitioned over the
    ranks. Every rank times exactly K iterations on its own stream with CUDA events (the fused solver's
    iter_ms), bracketed by barrier + synchronize; rank 0 reports K / max over ranks."""
    import torch
    import bench as bench_mod   # phase_times, SLOTS, KERNEL_NAME (bench.py is the entry script: already imported)
    world, rank = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"])
    dist = init_process_group(cuda=True)
    local_rank = local_device()
    mode = capi.COMM_NCCL if args.comm == "nccl" else capi.COMM_P2P
    method = capi.PART_SLAB if args.partition == "slab" else capi.PART_METIS
    schedule = bench_mod.SCHEDULES(capi)[args.schedule]
    tuning = int(getattr(args, "tuning", 0))
    mesh, x_star = build_problem(args)
    part = partition_mesh(mesh, world, method)
    loc = part.local(rank)
    pinfo = part.info
    sampler = ClockSampler(local_rank).start() if rank == 0 else None   # early: nvidia-smi needs ~1 s to its first line
    ctx = DistContext(local_rank, rank, world, pinfo.vec_capacity, n_vectors=12, mode=mode)
    op = DistOperator(ctx, loc, prefill=0, dt=-1.0, form=FORM_COEF, dirichlet=True)
    n = loc.n_owned
    xs = ctx.vector(x_star[loc.owned_global])
    b = ctx.zeros(n)
    op.mul(b, xs)                                    # b = A x* (first halo exchange)
    Solver = BiCgStabSolver if args.solver == "bicgstab" else CgSolver
    applies_per_it, passes_per_it = (2, 15) if args.solver == "bicgstab" else (1, 9)

    def solve(iters, **kw):
        s = Solver(num_iterations=iters, absolute_error_tolerance=0.0, relative_error_tolerance=0.0,
                   record=False, **kw)
        x = ctx.zeros(n)
        dist.barrier()
        torch.cuda.synchronize()
        s.solve(x, b, op)
        torch.cuda.synchronize()
        dist.barrier()
        assert s.iteration == iters, (s.iteration, iters)
        return s, x

    solve(max(args.warmup, 3), use_graph=True, schedule=schedule, tuning=tuning)
    s, x = solve(args.steps, use_graph=True, schedule=schedule, tuning=tuning)
    iter_ms = max_over_ranks(s.iter_ms)
    launches = int(sum_over_ranks(s.launches))
    value = args.steps / (iter_ms * 1e-3)
    persistent = s.schedule_used == capi.SCHEDULE_PERSISTENT
    phases = None
    if persistent:   # the kernel's own timeline, per rank (rank 0's is printed; the waits are maxima over the ranks)
        k = min(args.steps, 64)
        st, _ = solve(k, schedule=schedule, timeline_iters=k, tuning=tuning)
        phases = bench_mod.phase_times(st.timeline, args.solver)
        if phases is not None:
            for key in ("us_in_barrier_cta0", "us_wait_for_other_ranks"):
                phases[key + "_max_over_ranks"] = {nm: max_over_ranks(v) for nm, v in phases[key].items()}
            phases["us_halo_wait_max_over_ranks"] = [max_over_ranks(v) for v in phases["us_halo_wait_max"]]
            phases["us_per_iteration_max_over_ranks"] = max_over_ranks(phases["us_per_iteration"])
    sp, _ = solve(args.steps, profile=True, tuning=tuning,
                  schedule=capi.SCHEDULE_FOLDED if s.schedule_used == capi.SCHEDULE_FOLDED else capi.SCHEDULE_STEPWISE)
    kms = [max_over_ranks(v) for v in sp.kernel_ms]
    wms = [max_over_ranks(v) for v in sp.wait_ms]
    ams = [max_over_ranks(v) for v in sp.ar_wait_ms]
    xg = gather_global(loc, x.numpy(), mesh.n_cells)
    err = float(np.linalg.norm(xg - x_star) / np.linalg.norm(x_star))

    # e2e: every rank passes HOST buffers of its shard through sb_solve_host; wall clock between barriers
    hx = torch.zeros(n, dtype=torch.float64).pin_memory()
    hb = torch.from_numpy(b.numpy()).pin_memory()
    hxn, hbn = hx.numpy(), hb.numpy()
    solve_host(ctx, op, args.solver, hxn, hbn, num_iterations=3, abs_tol=0.0, rel_tol=0.0, use_graph=True, schedule=schedule,
               tuning=tuning)
    hxn[:] = 0.0
    ctx.sync()
    dist.barrier()
    t = time.perf_counter()
    rep = solve_host(ctx, op, args.solver, hxn, hbn, num_iterations=args.steps, abs_tol=0.0, rel_tol=0.0, use_graph=True,
                     schedule=schedule, tuning=tuning)
    dist.barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t)
    assert rep.iterations == args.steps
    clocks = sampler.stop() if sampler else None

    alg_apply = int(sum_over_ranks(op.info.algorithmic_bytes_per_apply))
    n_glob = mesh.n_cells
    comm_err = ctx.status()
    if rank == 0:
        peak, peak_src = peaks()
        slots = bench_mod.SLOTS[args.solver]
        apply_slots = [k for k, nm in enumerate(slots) if nm.startswith("apply")]
        apply_ms = sum(kms[k] for k in apply_slots) / (len(apply_slots) * args.steps)
        alg_iter = applies_per_it * alg_apply + passes_per_it * 8 * n_glob
        alg_launch = alg_iter * args.steps if persistent else alg_apply
        launch_ms = iter_ms if persistent else apply_ms
        achieved = alg_launch / (launch_ms * 1e-3) / 1e9       # all ranks' bytes / slowest rank's launch time
        cfg = workload_config(args, mesh)
        cfg.update({"partition": args.partition, "comm": args.comm, "edge_cut": int(pinfo.edge_cut),
                    "cells_per_rank": [int(pinfo.min_owned), int(pinfo.max_owned)], "max_halo": int(pinfo.max_halo),
                    "schedule": bench_mod.schedule_name(capi, s.schedule_used), "tuning": tuning})
        line = {
            "metric": "krylov_iterations_per_sec", "value": value, "unit": "it/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": iter_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg,
            "roofline": {"bound": "hbm",
                         "kernel": (bench_mod.KERNEL_NAME[args.solver] + ", halo push + all-reduce inside, all ranks") if persistent
                         else "operator apply + fused dot(s) incl. halo pack/wait (all ranks)",
                         "achieved": achieved, "peak": peak * world, "unit": "GB/s", "frac": achieved / (peak * world),
                         "traffic": None, "peak_source": peak_src + f" x {world} GPUs",
                         "algorithmic_bytes_per_launch": int(alg_launch), "avg_launch_ms": launch_ms,
                         "share_of_step": 1.0 if persistent else
                         sum(kms[k] for k in apply_slots) / max(sum(kms[:len(slots)]), 1e-12)},
            "iteration_roofline": {"algorithmic_bytes_per_iteration": int(alg_iter),
                                   "achieved_gbs": alg_iter * value / 1e9,
                                   "frac_of_measured_peak": alg_iter * value / 1e9 / (peak * world),
                                   "frac_of_nominal_8TBs": alg_iter * value / (8e12 * world)},
            "phases": phases,
            "stepwise": {"kernel_ms_per_iteration": {nm: kms[k] / args.steps for k, nm in enumerate(slots)},
                         "in_kernel_wait_ms_per_iteration": {nm: wms[k] / args.steps for k, nm in enumerate(slots)},
                         "allreduce_wait_ms_per_iteration": {nm: ams[k] / args.steps for k, nm in enumerate(slots)},
                         "apply_avg_launch_ms": apply_ms,
                         "note": "profiled run (events around every launch, no graph), maxima over the ranks; in-kernel "
                                 "wait: apply slots = a boundary CTA waiting for a neighbour's halo values; all-reduce wait "
                                 "= the reducer of the slot waiting for the other ranks' sums (rank skew + one NVLink "
                                 "crossing)"},
            "applies_per_sec": applies_per_it * value,
            "cpu_baseline": None,
            "e2e": {"value": args.steps / e2e_s, "unit": "it/s", "h2d_bytes_per_step": 16 * n_glob / args.steps,
                    "d2h_bytes_per_step": 8 * n_glob / args.steps,
                    "note": f"per rank: one sb_solve_host call = H2D(x0,b shard) + init + {args.steps} iterations + "
                            f"D2H(x shard), pinned host buffers; wall clock, max over ranks"},
            "gpu_launches": launches, "clocks": clocks,
            "rel_error_vs_exact_after_steps": err, "residual_after_steps": float(s.absolute_error),
            "comm_error_word": comm_err,
        }
        print(json.dumps(line), flush=True)
    del op, xs, b, x, s, sp
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()
