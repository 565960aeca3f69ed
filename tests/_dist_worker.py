"""Worker of the world_size > 1 tests (launched by tests/test_multigpu.py through torch.distributed.run).

    cpu mode (gloo, no GPU): the host-side logic of the N>1 path -- partition broadcast, local meshes,
        halo maps -- is exercised end to end: the ranks exchange halo values with gloo send/recv following
        send_idx / recv_ptr / send_dst exactly as the device code does, apply their local face lists with
        the oracle, and every owned row must be bit-identical to the single-process result; the
        rank-ordered segmented tree dot is checked against per-rank tree sums combined in rank order.
    gpu mode (one rank per GPU): distributed apply / CG / BiCGStab through the C ABI, checked bit for bit
        against the single-process oracle run in the ranks' concatenated cell order with the segmented
        reduction tree (oracle RED_TREE_SEG).
Exit code 0 = all checks passed on this rank.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import orc  # noqa: E402
from stormruler_b200 import capi  # noqa: E402
from stormruler_b200 import multigpu as mg  # noqa: E402
from stormruler_b200.mesh import CELL_HEX, CELL_TET, Mesh, PolyMesh  # noqa: E402


def face_mesh(m):
    return orc.FaceMesh(m.n_cells, m.face_cell, m.face_area, m.face_dist, m.cell_vol, m.bface_cell, m.bface_area,
                        m.bface_dist)


def concat_order(part, world):
    """Global ids in the ranks' concatenated local order + segment pointers."""
    order, seg = [], [0]
    for r in range(world):
        L = part.local(r)
        order.append(np.asarray(L.owned_global, np.int64))
        seg.append(seg[-1] + L.n_owned)
    return np.concatenate(order), np.asarray(seg, np.int64)


def permuted_face_mesh(mesh, order):
    """The global mesh with cells relabelled to the concatenated order; faces keep their global order, so
    every cell still sums its face contributions in ascending global face index."""
    g2c = np.empty(mesh.n_cells, np.int32)
    g2c[order] = np.arange(mesh.n_cells, dtype=np.int32)
    return orc.FaceMesh(mesh.n_cells, g2c[mesh.face_cell], mesh.face_area, mesh.face_dist, mesh.cell_vol[order],
                        g2c[mesh.bface_cell] if mesh.n_bfaces else mesh.bface_cell, mesh.bface_area, mesh.bface_dist)


def host_halo_exchange(dist, loc, x_local):
    """What halo_pack + the peer stores do on the device, over gloo."""
    import torch
    reqs, recv_bufs = [], []
    for k in range(loc.n_nbr):
        q = int(loc.nbr_rank[k])
        sb = torch.from_numpy(np.ascontiguousarray(x_local[loc.send_idx[loc.send_ptr[k]:loc.send_ptr[k + 1]]]))
        rb = torch.empty(int(loc.recv_ptr[k + 1] - loc.recv_ptr[k]), dtype=torch.float64)
        recv_bufs.append(rb)
        reqs.append(dist.isend(sb, dst=q))
        reqs.append(dist.irecv(rb, src=q))
    for r in reqs:
        r.wait()
    for k in range(loc.n_nbr):
        x_local[loc.halo_base + loc.recv_ptr[k]: loc.halo_base + loc.recv_ptr[k + 1]] = recv_bufs[k].numpy()


def run_cpu(dist, rank, world):
    import torch
    for kind, dims, method in ((CELL_TET, (6, 5, 4), capi.PART_METIS), (CELL_HEX, (7, 6, 5), capi.PART_SLAB),
                               (CELL_TET, (9, 3, 3), capi.PART_SLAB), ("poly", (4,), capi.PART_METIS)):
        if kind == "poly":   # 14-face cells, ingested as a face list (sb_mesh_from_faces)
            mesh = PolyMesh.bcc(*dims, stretch=(1.0, 1.3, 0.7)).to_mesh()
            mesh.permute_cells(np.random.default_rng(43).permutation(mesh.n_cells).astype(np.int32))
        else:
            mesh = Mesh.box(kind, *dims, jitter=0.2, seed_jitter=42, shuffle=True, seed_shuffle=43)
        mesh.renumber_rcm()
        part = mg.partition_mesh(mesh, world, method)
        loc = part.local(rank)
        n = mesh.n_cells
        # every rank holds the same partition array
        ref_part = mg.broadcast_array(np.asarray(part.part) if rank == 0 else None, (n,), np.int32)
        assert np.array_equal(ref_part, part.part)
        # send_dst is the neighbour's halo_base + its recv_ptr for me: ask the neighbours
        mine = torch.zeros(world, dtype=torch.int64)
        for k in range(loc.n_nbr):
            mine[int(loc.nbr_rank[k])] = loc.halo_base + int(loc.recv_ptr[k])   # where nbr k's block starts in MY vectors
        table = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(table, mine)
        for k in range(loc.n_nbr):
            q = int(loc.nbr_rank[k])
            assert int(table[q][rank]) == int(loc.send_dst[k]), "send_dst disagrees with the neighbour's layout"
        # distributed apply on the host = halo exchange + local face loop; owned rows bit-identical
        rng = np.random.default_rng(7)
        xg = rng.standard_normal(n)
        gop = orc.FaceOp(face_mesh(mesh), prefill=1, dt=-0.05, dirichlet=True)
        yg = gop.apply(xg)
        xl = np.zeros(loc.n_cells)
        xl[:loc.n_owned] = xg[loc.owned_global]
        host_halo_exchange(dist, loc, xl)
        assert np.array_equal(xl[loc.halo_base:loc.halo_base + loc.n_halo], xg[loc.halo_global]), "halo values"
        lop = orc.FaceOp(face_mesh(loc), prefill=1, dt=-0.05, dirichlet=True)
        yl = lop.apply(xl)
        assert np.array_equal(yl[:loc.n_owned], yg[loc.owned_global]), "owned rows differ from the global apply"
        assert np.array_equal(mg.gather_global(loc, yl[:loc.n_owned], n), yg)
        # rank-ordered reduction: per-rank SB_TREE sums added in rank order == oracle RED_TREE_SEG
        order, seg = concat_order(part, world)
        orc.set_segments(seg)
        want = orc.dot(xg[order], yg[order], orc.RED_TREE_SEG)
        local_sum = torch.tensor([orc.dot(xl[:loc.n_owned], yl[:loc.n_owned], orc.RED_TREE)], dtype=torch.float64)
        sums = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(sums, local_sum)
        tot = float(sums[0][0])
        for r in range(1, world):
            tot = tot + float(sums[r][0])
        assert tot == want, (tot, want)
        assert mg.max_over_ranks(float(rank)) == world - 1 and mg.sum_over_ranks(1.0) == world
        if kind in (CELL_HEX, "poly"):
            # scatter mode: only rank 0 holds the global mesh; what arrives must be what part.local(rank) builds
            from stormruler_b200.mesh import LOCAL_ARRAYS, LOCAL_SCALARS
            the_mesh, the_part = mesh, np.array(part.part, np.int32, copy=True)

            class FixedPartition(mg.Partition):   # rank 0 reuses the partition every rank already agreed on
                def __init__(self, m, n_parts, method_):
                    super().__init__(m, n_parts, part=the_part)

            saved = mg.Partition
            mg.Partition = FixedPartition
            try:
                sl, info, fields = mg.scatter_mesh(lambda: the_mesh, world, method,
                                                   cell_fields=lambda m: {"vol": np.asarray(m.cell_vol), "ctr": m.cell_centers()})
            finally:
                mg.Partition = saved
            for k in LOCAL_SCALARS:
                assert getattr(sl, k) == getattr(loc, k), k
            for k, _ in LOCAL_ARRAYS:
                assert np.array_equal(np.asarray(getattr(sl, k)), np.asarray(getattr(loc, k))), k
            assert info["vec_capacity"] == part.info.vec_capacity and info["edge_cut"] == part.info.edge_cut
            assert (info["mesh"] is not None) == (rank == 0)
            assert np.array_equal(fields["vol"], np.asarray(mesh.cell_vol)[loc.owned_global])
            assert np.array_equal(fields["ctr"], mesh.cell_centers()[loc.owned_global])
            # the struct built from the received arrays feeds the oracle like the library's own
            yl2 = orc.FaceOp(face_mesh(sl), prefill=1, dt=-0.05, dirichlet=True).apply(xl)
            assert np.array_equal(yl2[:sl.n_owned], yg[sl.owned_global])
    return 0


def run_cpu_lattice(dist, rank, world):
    """Rank-local lattice slabs (mesh.HexLatticeSlab: config 4 / config 5 at full size): every rank builds ONLY its own
    local mesh from lattice arithmetic, the ranks exchange halos over gloo following those maps, and every owned row of
    the local operator must equal the row of the global operator -- which is built here for the check alone."""
    from stormruler_b200.mesh import HexLattice, HexLatticeSlab
    for dims in ((9, 7, 8), (5, 4, 23), (12, 12, 3)):
        slab = HexLatticeSlab(*dims, rank, world)
        loc = slab.local
        lat = HexLattice(*dims)
        n = lat.n_cells
        xg = np.random.default_rng(17).standard_normal(n)
        yg = orc.FaceOp(face_mesh(lat), prefill=1, dt=-0.05, dirichlet=True).apply(xg)
        xl = np.zeros(loc.n_cells)
        xl[:loc.n_owned] = xg[loc.owned_global]
        host_halo_exchange(dist, loc, xl)
        assert np.array_equal(xl[loc.halo_base:], xg[loc.halo_global]), "halo values landed in the wrong slots"
        yl = orc.FaceOp(face_mesh(loc), prefill=1, dt=-0.05, dirichlet=True).apply(xl)
        assert np.array_equal(yl[:loc.n_owned], yg[loc.owned_global]), f"rank-local slab rows differ from the global operator {dims}"
        # the summary every rank derives on its own agrees across ranks
        import torch
        cap = torch.tensor([slab.info()["vec_capacity"], slab.info()["edge_cut"]], dtype=torch.int64)
        caps = [torch.zeros_like(cap) for _ in range(world)]
        dist.all_gather(caps, cap)
        assert all(torch.equal(c, cap) for c in caps)
    return 0


def run_gpu(dist, rank, world, mode_name):
    import stormruler_b200 as sb
    mode = capi.COMM_NCCL if mode_name == "nccl" else capi.COMM_P2P
    local_rank = mg.local_device()
    cases = ((CELL_TET, (14, 12, 10), capi.PART_METIS), (CELL_HEX, (24, 20, 18), capi.PART_SLAB),
             ("poly", (16,), capi.PART_METIS))
    for kind, dims, method in cases:
        if kind == "poly":   # 14-wide rows: the wide instantiation of the apply kernel with the fused halo exchange
            mesh = PolyMesh.bcc(*dims, stretch=(1.0, 1.3, 0.7)).to_mesh()
            mesh.permute_cells(np.random.default_rng(43).permutation(mesh.n_cells).astype(np.int32))
        else:
            mesh = Mesh.box(kind, *dims, jitter=0.2, seed_jitter=42, shuffle=True, seed_shuffle=43)
        mesh.renumber_rcm()
        part = mg.partition_mesh(mesh, world, method)
        loc = part.local(rank)
        n = mesh.n_cells
        order, seg = concat_order(part, world)
        orc.set_segments(seg)
        pm = permuted_face_mesh(mesh, order)
        ctx = mg.DistContext(local_rank, rank, world, part.info.vec_capacity, n_vectors=30, mode=mode)
        rng = np.random.default_rng(11)
        xg = rng.standard_normal(n)
        bg = np.sin(0.37 * np.arange(n))
        for form in (sb.FORM_FAITHFUL, sb.FORM_COEF):
            op = mg.DistOperator(ctx, loc, prefill=1, dt=-0.05, form=form, dirichlet=True)
            cpu = orc.FaceOp(pm, prefill=1, dt=-0.05, dirichlet=True)
            # apply, several times in a row on alternating inputs (exercises the ack / sequence protocol)
            x, y = ctx.vector(xg[loc.owned_global]), ctx.zeros(loc.n_owned)
            if form == sb.FORM_FAITHFUL:
                want = cpu.apply(xg[order])
            else:
                want = cpu.apply_rows_coef(xg[order])
            for rep in range(5):
                op.mul(y, x)
                got = mg.gather_global(loc, y.numpy(), n)
                assert np.array_equal(got[order], want), f"distributed apply differs (form {form}, repeat {rep})"
            # ping-pong y = A x, x2 = A y: consecutive applies with different halo contents
            x2 = ctx.zeros(loc.n_owned)
            op.mul(x2, y)
            want2 = cpu.apply(want) if form == sb.FORM_FAITHFUL else cpu.apply_rows_coef(want)
            assert np.array_equal(mg.gather_global(loc, x2.numpy(), n)[order], want2)
            # distributed dot
            d = ctx.dot(x, y)
            assert d == orc.dot(xg[order], want, orc.RED_TREE_SEG), "rank-ordered dot differs"
            # faithful form: the stepwise schedule (with and without graph replay); coefficient form: the stepwise one
            # and the persistent whole-solve kernel (in-kernel halo push, grid barriers, mailbox all-reduce; P2P only),
            # alternating on one context so that the sequence numbers of the two protocols have to stay in step
            oracle_op = cpu if form == sb.FORM_FAITHFUL else orc.RowsOp(n, *cpu.rows_coef())
            # (schedule, graph replay, tuning bits): every combination must produce the same bits
            T = capi
            if form == sb.FORM_FAITHFUL:
                variants = [(capi.SCHEDULE_AUTO, False, 0), (capi.SCHEDULE_AUTO, True, 0)]
                if mode == capi.COMM_P2P:
                    variants += [(capi.SCHEDULE_STEPWISE, True, T.TUNE_PUSH_ON_PRODUCE | T.TUNE_PUSH_LAZY),
                                 (capi.SCHEDULE_STEPWISE, True, T.TUNE_PUSH_ON_PRODUCE | T.TUNE_PDL_FINAL),
                                 (capi.SCHEDULE_STEPWISE, True, T.TUNE_PUSH_ON_PRODUCE | T.TUNE_IN_KERNEL_REDUCER)]
            elif mode == capi.COMM_P2P:
                variants = [(capi.SCHEDULE_PERSISTENT, False, 0), (capi.SCHEDULE_STEPWISE, True, T.TUNE_OFF),
                            (capi.SCHEDULE_STEPWISE, True, T.TUNE_PUSH_ON_PRODUCE | T.TUNE_STREAM_OPERATOR),
                            (capi.SCHEDULE_PERSISTENT, False, 0), (capi.SCHEDULE_AUTO, False, 0),
                            (capi.SCHEDULE_STEPWISE, False, T.TUNE_PUSH_ON_PRODUCE | T.TUNE_PDL_FINAL | T.TUNE_PDL_AFTER_FINAL),
                            (capi.SCHEDULE_FOLDED, True, 0),
                            (capi.SCHEDULE_STEPWISE, True, T.TUNE_PUSH_ON_PRODUCE | T.TUNE_PUSH_LAZY | T.TUNE_STREAM_OPERATOR),
                            (capi.SCHEDULE_STEPWISE, False, T.TUNE_PUSH_ON_PRODUCE | T.TUNE_PUSH_LAZY | T.TUNE_IN_KERNEL_REDUCER),
                            (capi.SCHEDULE_STEPWISE, True, T.TUNE_PUSH_ON_PRODUCE | T.TUNE_PUSH_LAZY | T.TUNE_IN_KERNEL_REDUCER
                             | T.TUNE_STREAM_OPERATOR | T.TUNE_PDL_AFTER_FINAL | T.TUNE_PDL_APPLY),
                            (capi.SCHEDULE_STEPWISE, True, T.TUNE_IN_KERNEL_REDUCER),
                            (capi.SCHEDULE_STEPWISE, False, T.TUNE_IN_KERNEL_REDUCER | T.TUNE_PUSH_ON_PRODUCE),
                            (capi.SCHEDULE_STEPWISE, True, T.TUNE_IN_KERNEL_REDUCER | T.TUNE_PUSH_ON_PRODUCE | T.TUNE_STREAM_OPERATOR
                             | T.TUNE_PDL_AFTER_FINAL | T.TUNE_PDL_APPLY),
                            (capi.SCHEDULE_STEPWISE, True, T.TUNE_NO_ACK | T.TUNE_PDL_APPLY | T.TUNE_PDL_FINAL | T.TUNE_PDL_AFTER_FINAL),
                            (capi.SCHEDULE_STEPWISE, True, T.TUNE_PUSH_ON_PRODUCE | T.TUNE_PDL_APPLY | T.TUNE_PDL_FINAL
                             | T.TUNE_PDL_AFTER_FINAL | T.TUNE_STREAM_OPERATOR)]
            else:
                variants = [(capi.SCHEDULE_STEPWISE, True, 0)]
            for name, Solver in (("cg", sb.CgSolver), ("bicgstab", sb.BiCgStabSolver)):
                w = orc.solve(name, oracle_op, bg[order], num_iterations=60, abs_tol=0.0, rel_tol=1e-9, mode=orc.RED_TREE_SEG)
                for schedule, use_graph, tuning in variants:
                    s = Solver(num_iterations=60, absolute_error_tolerance=0.0, relative_error_tolerance=1e-9,
                               use_graph=use_graph, check_every=7, schedule=schedule, tuning=tuning)
                    xs = ctx.zeros(loc.n_owned)
                    conv = s.solve(xs, ctx.vector(bg[loc.owned_global]), op)
                    tag = f"{name} form {form} schedule {schedule}->{s.schedule_used} tuning {tuning:#x}"
                    if form == sb.FORM_COEF and mode == capi.COMM_P2P and schedule == capi.SCHEDULE_PERSISTENT:
                        assert s.schedule_used == capi.SCHEDULE_PERSISTENT, tag
                    assert conv == w.converged and s.iteration == w.iterations, (tag, conv, s.iteration, w.iterations)
                    assert np.array_equal(s.history, w.hist), f"{tag}: residual history differs from the oracle"
                    assert np.array_equal(mg.gather_global(loc, xs.numpy(), n)[order], w.x), f"{tag}: solution differs"
                    del xs
                # a plain apply between two solves: the ack / sequence protocol of the stepwise kernels picks up where
                # the persistent kernel left the counters
                op.mul(y, x)
                assert np.array_equal(mg.gather_global(loc, y.numpy(), n)[order], want), "apply after the solves differs"
            del op, x, y, x2
        # non-symmetric rows + the fused GMRES over the distributed operator (config 3)
        fu, bu = mesh.face_flux((1.0, 0.5, 0.25))
        cop = mg.DistConvDiffOperator(ctx, loc, 0.02, fu, bu)
        ccpu = orc.ConvDiffOp(pm, 0.02, fu, bu).rows_coef()
        xs = ctx.zeros(loc.n_owned)
        g = sb.GmresSolver(num_iterations=80, absolute_error_tolerance=0.0, relative_error_tolerance=1e-9, num_inner_iterations=11)
        conv = g.solve(xs, ctx.vector(bg[loc.owned_global]), cop)
        w = orc.ref_solve("gmres", ccpu, bg[order], num_iterations=80, abs_tol=0.0, rel_tol=1e-9, num_inner=11, mode=orc.RED_TREE_SEG)
        assert (conv, g.iteration) == (w.converged, w.iterations), (conv, g.iteration, w.converged, w.iterations)
        assert np.array_equal(g.trace, w.trace) and np.array_equal(g.history, w.hist), "distributed GMRES differs from the oracle"
        assert np.array_equal(mg.gather_global(loc, xs.numpy(), n)[order], w.x)
        del cop, xs
        assert ctx.status() == 0
        ctx.close()
    return 0


def main():
    mode = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist = mg.init_process_group(cuda=(mode != "cpu"))
    try:
        if mode == "cpu":
            rc = run_cpu(dist, rank, world) or run_cpu_lattice(dist, rank, world)
        else:
            rc = run_gpu(dist, rank, world, mode)
        dist.barrier()
    finally:
        dist.destroy_process_group()
    print(f"rank {rank}: OK", flush=True)
    sys.exit(rc)


if __name__ == "__main__":
    main()
