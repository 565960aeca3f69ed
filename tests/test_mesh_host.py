"""Host-side integer work of the path (no GPU): cell soup -> face-list SoA, RCM renumbering,
partitioning, halo maps. Bar (north_star): bit-exact. Oracle: the independent numpy restatement in
oracle/mesh_oracle.py (conventions from the reference's mesh layer, cited there)."""
import numpy as np
import pytest

from oracle import mesh_oracle as mo
from oracle import orc
from stormruler_b200 import capi
from stormruler_b200.mesh import (CELL_HEX, CELL_TET, LOCAL_ARRAYS, LOCAL_SCALARS, HexLattice, HexLatticeSlab, Mesh,
                                  Partition, PolyMesh)

KIND = {"tet": CELL_TET, "hex": CELL_HEX}
SOA_KEYS = ("face_cell", "face_area", "face_dist", "cell_vol", "bface_cell", "bface_area", "bface_dist")


def assert_soa_equal(mesh, want):
    assert mesh.n_cells == want["n_cells"]
    for k in SOA_KEYS:
        got = np.asarray(getattr(mesh, k))
        assert got.shape == want[k].shape, k
        assert np.array_equal(got, want[k]), f"{k} differs from the restatement"


def test_mt19937_64_restatement_known_answers():
    """std::mt19937_64 default seed: first output 14514284786278117030, 10000th 9981545732273789042
    (the values the C++ standard itself requires, [rand.predef])."""
    e = mo.MT19937_64()
    first = e()
    for _ in range(9998):
        e()
    assert first == 14514284786278117030 and e() == 9981545732273789042


@pytest.mark.parametrize("kind,dims", [("tet", (5, 4, 3)), ("hex", (6, 5, 4)), ("tet", (1, 1, 1)), ("hex", (1, 1, 2))])
@pytest.mark.parametrize("shuffle", [False, True])
def test_box_mesh_soa_bit_exact(kind, dims, shuffle):
    mesh = Mesh.box(KIND[kind], *dims, jitter=0.2, seed_jitter=42, shuffle=shuffle, seed_shuffle=43)
    xyz, cells = mo.box_cells(kind, *dims, jitter=0.2, seed_jitter=42, shuffle=shuffle, seed_shuffle=43)
    want = mo.face_list(xyz, cells)
    assert_soa_equal(mesh, want)
    assert np.array_equal(mesh.cell_centers(), want["cell_ctr"])
    # closedness: per cell, the area-weighted outward normals cancel <=> the Laplacian annihilates constants
    n_faces_expected = {"tet": 4, "hex": 6}[kind] * mesh.n_cells
    assert 2 * mesh.n_faces + mesh.n_bfaces == n_faces_expected


@pytest.mark.parametrize("kind,dims", [("tet", (5, 4, 3)), ("hex", (6, 5, 4))])
@pytest.mark.parametrize("shuffle", [False, True])
def test_face_normals_bit_exact_and_closed(kind, dims, shuffle):
    """Unit normals (what FaceView::normal() feeds the reference's flux schemes): bit-exact against the numpy
    restatement, unit length, oriented inner -> outer / outward; a tetrahedron's area vectors cancel."""
    mesh = Mesh.box(KIND[kind], *dims, jitter=0.2, seed_jitter=42, shuffle=shuffle, seed_shuffle=43)
    want = mo.face_list(*mo.box_cells(kind, *dims, jitter=0.2, seed_jitter=42, shuffle=shuffle, seed_shuffle=43))
    fn, bn = mesh.face_normals()
    assert np.array_equal(fn, want["face_normal"]) and np.array_equal(bn, want["bface_normal"])
    assert np.abs(np.linalg.norm(fn, axis=1) - 1.0).max() < 4e-16
    ctr = mesh.cell_centers()
    assert (np.einsum("ij,ij->i", fn, ctr[mesh.face_cell[:, 1]] - ctr[mesh.face_cell[:, 0]]) > 0).all()
    # boundary normals of the unit box are axis-aligned and point out of it
    assert np.allclose(np.abs(bn).max(axis=1), 1.0, atol=1e-12)
    if kind == "tet":
        acc = np.zeros((mesh.n_cells, 3))
        np.add.at(acc, mesh.face_cell[:, 0], mesh.face_area[:, None] * fn)
        np.add.at(acc, mesh.face_cell[:, 1], -mesh.face_area[:, None] * fn)
        np.add.at(acc, mesh.bface_cell, mesh.bface_area[:, None] * bn)
        assert np.abs(acc).max() < 1e-15
    fu, bu = mesh.face_flux((1.0, 0.5, 0.25))
    assert np.array_equal(fu, (1.0 * fn[:, 0] + 0.5 * fn[:, 1]) + 0.25 * fn[:, 2]) and bu.shape == (mesh.n_bfaces,)


@pytest.mark.parametrize("kind", ["tet", "hex"])
def test_ingestion_from_cells_matches_generator(kind):
    xyz, cells = mo.box_cells(kind, 4, 3, 3)
    a = Mesh.from_cells(KIND[kind], xyz, cells)
    b = Mesh.box(KIND[kind], 4, 3, 3)
    for k in SOA_KEYS:
        assert np.array_equal(getattr(a, k), getattr(b, k))
    with pytest.raises(capi.StormB200Error):
        bad = cells.copy()
        bad[0, 0] = len(xyz)  # node index out of range
        Mesh.from_cells(KIND[kind], xyz, bad)


@pytest.mark.parametrize("first_index", [0, 1])
def test_tetgen_reader_3d(tmp_path, first_index):
    """`.node` / `.ele` in the grammar the reference's reader accepts (IoTetgen.hpp:44-235): comments, attributes,
    labels, 0- or 1-based indices. The mesh read back is the mesh built from the same arrays."""
    xyz, cells = mo.box_cells("tet", 4, 3, 2)
    prefix = tmp_path / "box.1"
    with open(str(prefix) + ".node", "w") as f:
        f.write(f"# nodes\n{len(xyz)} 3 1 1\n")
        for k, p in enumerate(xyz):
            f.write(f"{k + first_index} {float(p[0])!r} {float(p[1])!r} {float(p[2])!r} 0.5 {k % 3}  # attribute, label\n")
    with open(str(prefix) + ".ele", "w") as f:
        f.write(f"{len(cells)} 4 1\n")
        for k, c in enumerate(cells):
            f.write(f"{k + first_index} " + " ".join(str(int(v) + first_index) for v in c) + " 7\n")
    a, b = Mesh.read_tetgen(prefix), Mesh.from_cells(CELL_TET, xyz, cells)
    assert a.n_cells == b.n_cells
    for k in SOA_KEYS:
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    with pytest.raises(capi.StormB200Error):
        Mesh.read_tetgen(tmp_path / "missing")
    with open(str(prefix) + ".ele", "w") as f:
        f.write("1 4 0\n0 0 1 2 99999\n")
    with pytest.raises(capi.StormB200Error):
        Mesh.read_tetgen(prefix)


@pytest.mark.parametrize("kind,dims", [("tet", (6, 5, 4)), ("hex", (7, 6, 5))])
def test_rcm_permutation_and_renumbered_mesh_bit_exact(kind, dims):
    mesh = Mesh.box(KIND[kind], *dims)
    xyz, cells = mo.box_cells(kind, *dims)
    before = mo.face_list(xyz, cells)
    bw0 = mesh.bandwidth
    perm = mesh.renumber_rcm()
    want_perm = mo.rcm(before["n_cells"], before["face_cell"])
    assert np.array_equal(perm, want_perm), "RCM permutation differs from the restatement"
    assert sorted(perm.tolist()) == list(range(mesh.n_cells))          # a bijection, perm[new] = old
    assert_soa_equal(mesh, mo.face_list(*mo.permute(xyz, cells, want_perm)))
    assert mesh.bandwidth < bw0 // 4                                    # and it does its job
    # permuting back restores the original face list (Utils/Permutations.hpp semantics)
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(perm), dtype=np.int32)
    mesh.permute_cells(inv)
    assert_soa_equal(mesh, before)
    with pytest.raises(capi.StormB200Error):
        mesh.permute_cells(np.zeros(mesh.n_cells, np.int32))            # not a permutation


def as_dict(mesh):
    d = {k: np.asarray(getattr(mesh, k)) for k in SOA_KEYS}
    d["n_cells"] = mesh.n_cells
    return d


LOCAL_KEYS = ("local_to_global", "nbr_rank", "send_ptr", "recv_ptr", "send_idx", "send_dst", "face_global", "bface_global",
              "face_cell", "face_area", "face_dist", "cell_vol", "bface_cell", "bface_area", "bface_dist")


@pytest.mark.parametrize("method,n_parts", [(capi.PART_SLAB, 2), (capi.PART_SLAB, 5), (capi.PART_METIS, 2),
                                            (capi.PART_METIS, 4), (capi.PART_METIS, 8)])
def test_partition_local_meshes_and_halo_maps_bit_exact(method, n_parts):
    mesh = Mesh.box(CELL_TET, 8, 7, 6)
    mesh.renumber_rcm()
    P = Partition(mesh, n_parts, method)
    gd = as_dict(mesh)
    if method == capi.PART_SLAB:
        assert np.array_equal(P.part, mo.slab_partition(mesh.n_cells, n_parts))
    else:
        # METIS is a third-party black box: its output is pinned by determinism and validity
        assert np.array_equal(P.part, Partition(mesh, n_parts, method).part)
        counts = np.bincount(P.part, minlength=n_parts)
        assert counts.min() > 0 and counts.max() <= 1.1 * mesh.n_cells / n_parts + 1
    fc = gd["face_cell"]
    assert P.info.edge_cut == int((P.part[fc[:, 0]] != P.part[fc[:, 1]]).sum())
    cap = 0
    for r in range(n_parts):
        L, want = P.local(r), mo.local_maps(gd, P.part, r, n_parts)
        for k in ("n_owned", "n_interior", "n_halo", "halo_base"):
            assert getattr(L, k) == want[k], (r, k)
        for k in LOCAL_KEYS:
            got = np.asarray(getattr(L, k))
            assert got.shape == want[k].shape and np.array_equal(got, want[k]), (r, k)
        cap = max(cap, mo.pad_up(L.n_owned) + mo.pad_up(max(L.n_halo, 1)))
    assert P.info.vec_capacity == cap
    # every cell is owned exactly once; what a sends to b is what b expects from a, in the same order
    owned = np.concatenate([P.local(r).owned_global for r in range(n_parts)])
    assert sorted(owned.tolist()) == list(range(mesh.n_cells))
    for a in range(n_parts):
        La = P.local(a)
        for k, b in enumerate(La.nbr_rank):
            Lb = P.local(int(b))
            kb = list(Lb.nbr_rank).index(a)
            sent = La.local_to_global[La.send_idx[La.send_ptr[k]:La.send_ptr[k + 1]]]
            expected = Lb.halo_global[Lb.recv_ptr[kb]:Lb.recv_ptr[kb + 1]]
            assert np.array_equal(sent, expected)
            assert La.send_dst[k] == Lb.halo_base + Lb.recv_ptr[kb]


@pytest.mark.parametrize("method", [capi.PART_SLAB, capi.PART_METIS])
def test_partitioned_apply_rows_are_bit_identical_to_global_rows(method):
    """Emulated exchange (numpy): every owned row of every rank reproduces the global face loop bit for bit,
    because the local face lists keep the global face order and orientation."""
    mesh = Mesh.box(CELL_HEX, 7, 6, 5)
    mesh.renumber_rcm()
    gm = orc.FaceMesh(mesh.n_cells, mesh.face_cell, mesh.face_area, mesh.face_dist, mesh.cell_vol,
                      mesh.bface_cell, mesh.bface_area, mesh.bface_dist)
    x = np.cos(0.11 * np.arange(mesh.n_cells)) + 0.3
    y = orc.FaceOp(gm, prefill=0, dt=-1.0, dirichlet=True).apply(x)
    P = Partition(mesh, 3, method)
    for r in range(3):
        L = P.local(r)
        xl = np.zeros(L.n_cells)
        xl[:L.n_owned] = x[L.owned_global]
        xl[L.halo_base:] = x[L.halo_global]            # what the neighbours' packs deliver
        lm = orc.FaceMesh(L.n_cells, L.face_cell, L.face_area, L.face_dist, L.cell_vol, L.bface_cell, L.bface_area,
                          L.bface_dist)
        yl = orc.FaceOp(lm, prefill=0, dt=-1.0, dirichlet=True).apply(xl)
        assert np.array_equal(yl[:L.n_owned], y[L.owned_global])


def test_partition_rejects_bad_input():
    mesh = Mesh.box(CELL_TET, 2, 2, 2)
    with pytest.raises(capi.StormB200Error):
        Partition(mesh, 2, part=np.full(mesh.n_cells, 5, np.int32))      # part id out of range
    with pytest.raises(capi.StormB200Error):
        Partition(mesh, 2, part=np.zeros(mesh.n_cells, np.int32))        # part 1 owns nothing
    with pytest.raises(capi.StormB200Error):
        Partition(mesh, 10 ** 6)                                         # more parts than cells


@pytest.mark.parametrize("n,stretch", [(1, (1.0, 1.0, 1.0)), (5, (1.0, 1.0, 1.0)), (6, (1.0, 1.3, 0.7))])
def test_polyhedral_mesh_geometry_is_closed_and_consistent(n, stretch):
    """Truncated-octahedra mesh (config 5's dual-polyhedra leg): 14 faces per cell, every cell's area vectors sum to
    zero, the pyramid sum over the faces gives the cell volume, the cells tile the stretched box lattice, faces are
    ordered by creating cell with inner < outer, and the oracle's row forms (14 wide) agree with its face loop."""
    m = PolyMesh.bcc(n, stretch)
    N = 2 * n ** 3
    assert m.n_cells == N and m.face_cell.dtype == np.int32 and m.bface_cell.dtype == np.int32
    assert 2 * m.n_faces + m.n_bfaces == 14 * N
    assert (m.face_cell[:, 0] < m.face_cell[:, 1]).all() and (np.diff(m.face_cell[:, 0]) >= 0).all()
    assert (np.diff(m.bface_cell) >= 0).all()
    fn, bn = m.face_normals()
    S, V = np.zeros((N, 3)), np.zeros(N)
    np.add.at(S, m.face_cell[:, 0], m.face_area[:, None] * fn)
    np.add.at(S, m.face_cell[:, 1], -m.face_area[:, None] * fn)
    np.add.at(S, m.bface_cell, m.bface_area[:, None] * bn)
    assert np.abs(S).max() < 1e-15
    # pyramid heights: the face plane bisects the segment to the neighbour site, h = (M d / 2) . n
    step = m.DIRS * np.asarray(stretch)[None, :] * (0.5 / n)
    nrm = m.DIRS / np.asarray(stretch)[None, :]
    nrm = nrm / np.linalg.norm(nrm, axis=1, keepdims=True)
    h = 0.5 * (step * nrm).sum(axis=1)
    assert np.allclose(np.linalg.norm(step, axis=1)[m.face_dir], m.face_dist, rtol=1e-15)
    for cells, area, d in ((m.face_cell[:, 0], m.face_area, m.face_dir), (m.face_cell[:, 1], m.face_area, m.face_dir),
                           (m.bface_cell, m.bface_area, m.bface_dir)):
        np.add.at(V, cells, area * h[d] / 3.0)
    assert np.allclose(V, m.cell_vol, rtol=1e-13)
    assert np.isclose(m.cell_vol.sum(), stretch[0] * stretch[1] * stretch[2], rtol=1e-13)
    c = m.cell_centers()
    d_centres = np.linalg.norm(c[m.face_cell[:, 1]] - c[m.face_cell[:, 0]], axis=1)
    assert np.allclose(d_centres, m.face_dist, rtol=1e-13)
    fm = orc.FaceMesh(N, m.face_cell, m.face_area, m.face_dist, m.cell_vol, m.bface_cell, m.bface_area, m.bface_dist)
    cpu = orc.FaceOp(fm, prefill=1, dt=-0.05, dirichlet=True)
    assert cpu.rows_faithful()[0] == 14
    x = np.random.default_rng(3).standard_normal(N)
    want = cpu.apply(x)
    assert np.array_equal(cpu.apply_rows_faithful(x), want)
    assert np.abs(cpu.apply_rows_coef(x) - want).max() <= 1e-13 * np.abs(want).max()


def face_list_dict(m):
    return dict(n_cells=m.n_cells, **{k: np.asarray(getattr(m, k)) for k in SOA_KEYS})


@pytest.mark.parametrize("source", ["poly", "square_nb"])
def test_face_list_mesh_handle_renumbering_and_partition_bit_exact(source):
    """sb_mesh_from_faces: the mesh IS its face list (a polyhedral mesh; the reference mesh classes' own export of
    tests/_data/mesh/square_nb.1). Verbatim on creation; permutations and RCM against the numpy restatement;
    partition local meshes and halo maps against the same restatement as the node-based meshes."""
    if source == "poly":
        src = PolyMesh.bcc(5, (1.0, 1.3, 0.7))
        ctr = src.cell_centers()
        fn, bn = src.face_normals()
    else:
        from conftest import golden_mesh
        src = golden_mesh("square_nb")
        ctr = fn = bn = None
    mesh = Mesh.from_faces(src, ctr, fn, bn)
    before = face_list_dict(src)
    assert_soa_equal(mesh, before)                                     # taken verbatim
    if fn is not None:
        got_fn, got_bn = mesh.face_normals()
        assert np.array_equal(got_fn, fn) and np.array_equal(got_bn, bn)
        assert np.array_equal(mesh.cell_centers(), ctr)
    else:
        with pytest.raises(capi.StormB200Error):
            mesh.face_normals()
        with pytest.raises(capi.StormB200Error):
            mesh.cell_centers()
    n = mesh.n_cells
    shuffle = np.random.default_rng(7).permutation(n).astype(np.int32)
    mesh.permute_cells(shuffle)
    want = mo.permute_face_list(before, shuffle, fn, bn, ctr)
    assert_soa_equal(mesh, want)
    if fn is not None:
        got_fn, got_bn = mesh.face_normals()
        assert np.array_equal(got_fn, want["face_normal"]) and np.array_equal(got_bn, want["bface_normal"])
        assert np.array_equal(mesh.cell_centers(), want["cell_ctr"])
    assert (mesh.face_cell[:, 0] < mesh.face_cell[:, 1]).all()
    bw_shuffled = mesh.bandwidth
    perm = mesh.renumber_rcm()
    assert np.array_equal(perm, mo.rcm(n, want["face_cell"])), "RCM permutation differs from the restatement"
    total = shuffle[perm]                                              # old id of every new cell, both steps
    assert_soa_equal(mesh, mo.permute_face_list(before, total, fn, bn, ctr))
    assert mesh.bandwidth < bw_shuffled // 3
    # the operator does not care how the cells are numbered: y_new[k] == y_old[total[k]] up to the sum order
    x = np.cos(0.11 * np.arange(n)) + 0.3
    y0 = orc.FaceOp(orc.FaceMesh(n, *[before[k] for k in SOA_KEYS]), prefill=1, dt=-0.05, dirichlet=True).apply(x)
    fm = orc.FaceMesh(n, mesh.face_cell, mesh.face_area, mesh.face_dist, mesh.cell_vol, mesh.bface_cell,
                      mesh.bface_area, mesh.bface_dist)
    y1 = orc.FaceOp(fm, prefill=1, dt=-0.05, dirichlet=True).apply(x[total])
    assert np.abs(y1 - y0[total]).max() <= 1e-13 * np.abs(y0).max()
    # partitioning works on the handle like on a node-based mesh
    gd = as_dict(mesh)
    for method, n_parts in ((capi.PART_SLAB, 3), (capi.PART_METIS, 4)):
        P = Partition(mesh, n_parts, method)
        for r in range(n_parts):
            L, wl = P.local(r), mo.local_maps(gd, P.part, r, n_parts)
            for k in ("n_owned", "n_interior", "n_halo", "halo_base"):
                assert getattr(L, k) == wl[k], (r, k)
            for k in LOCAL_KEYS:
                got = np.asarray(getattr(L, k))
                assert got.shape == wl[k].shape and np.array_equal(got, wl[k]), (r, k)


def test_face_list_mesh_rejects_bad_input():
    src = PolyMesh.bcc(2)
    bad = face_list_dict(src)

    class M:
        pass
    m = M()
    for k, v in bad.items():
        setattr(m, k, v.copy() if isinstance(v, np.ndarray) else v)
    m.face_cell[3, 1] = m.n_cells                                       # out of range
    with pytest.raises(capi.StormB200Error):
        Mesh.from_faces(m)
    m.face_cell[3, 1] = m.face_cell[3, 0]                               # a face between a cell and itself
    with pytest.raises(capi.StormB200Error):
        Mesh.from_faces(m)
    fn, bn = src.face_normals()
    with pytest.raises(capi.StormB200Error):
        Mesh.from_faces(src, None, fn, None)                            # one normal array without the other


@pytest.mark.parametrize("dims", [(1, 1, 1), (2, 1, 1), (1, 3, 2), (5, 4, 3), (7, 7, 7)])
def test_hex_lattice_face_list_matches_the_node_based_generator(dims):
    """The direct (numpy) generator used for the largest sweep points: integer layout bit-identical to the node-based
    generator without jitter and shuffle (face order, inner/outer, boundary-face order), geometry within rounding
    (analytic h products against triangle cross products)."""
    a, b = HexLattice(*dims), Mesh.box(CELL_HEX, *dims, jitter=0.0, shuffle=False)
    assert (a.n_cells, a.n_faces, a.n_bfaces) == (b.n_cells, b.n_faces, b.n_bfaces)
    assert a.face_cell.dtype == np.int32 and np.array_equal(a.face_cell, b.face_cell)
    assert a.bface_cell.dtype == np.int32 and np.array_equal(a.bface_cell, b.bface_cell)
    for k in ("face_area", "face_dist", "cell_vol", "bface_area", "bface_dist"):
        assert np.allclose(getattr(a, k), getattr(b, k), rtol=1e-13, atol=0.0), k
    assert np.allclose(a.cell_centers(), b.cell_centers(), rtol=1e-13, atol=1e-15)
    assert a.bandwidth == b.bandwidth
    # and it feeds the oracle / the face-list handle like any other mesh
    fm = orc.FaceMesh(a.n_cells, a.face_cell, a.face_area, a.face_dist, a.cell_vol, a.bface_cell, a.bface_area, a.bface_dist)
    gm = orc.FaceMesh(b.n_cells, b.face_cell, b.face_area, b.face_dist, b.cell_vol, b.bface_cell, b.bface_area, b.bface_dist)
    x = np.cos(0.3 * np.arange(a.n_cells))
    ya = orc.FaceOp(fm, prefill=0, dt=-1.0, dirichlet=True).apply(x)
    yb = orc.FaceOp(gm, prefill=0, dt=-1.0, dirichlet=True).apply(x)
    assert np.abs(ya - yb).max() <= 1e-12 * max(np.abs(yb).max(), 1.0)
    h = Mesh.from_faces(a, a.cell_centers())
    assert np.array_equal(h.face_cell, a.face_cell) and h.bandwidth == a.bandwidth


@pytest.mark.parametrize("dims,n_parts", [((7, 5, 6), 3), ((4, 4, 4), 2), ((9, 3, 2), 4), ((16, 16, 16), 8),
                                          ((5, 1, 1), 2), ((30, 20, 10), 5), ((3, 3, 40), 8), ((13, 11, 3), 7),
                                          ((6, 6, 6), 1)])
def test_rank_local_lattice_slab_equals_partitioning_the_global_mesh(dims, n_parts):
    """HexLatticeSlab builds one rank's local mesh from lattice arithmetic alone (config 4 at full size: no rank holds
    the 49.8 M-cell mesh). Every scalar and every array -- local order, global face numbers, halo send/receive maps,
    boundary faces -- must equal what the general path produces: the lattice as a face-list mesh handle, split by
    SB_PART_SLAB, sb_part_local. Slabs thinner than one lattice plane (neighbours beyond the adjacent ranks), slab
    boundaries inside a plane and inside a row, and a single part are all covered."""
    lat = HexLattice(*dims)
    part = Partition(Mesh.from_faces(lat, lat.cell_centers()), n_parts, capi.PART_SLAB)
    for r in range(n_parts):
        want, slab = part.local(r), HexLatticeSlab(*dims, r, n_parts)
        for k in LOCAL_SCALARS:
            assert int(getattr(slab.local, k)) == int(getattr(want, k)), (r, k)
        for k, dt in LOCAL_ARRAYS:
            got, ref = np.asarray(getattr(slab.local, k)).reshape(-1), np.asarray(getattr(want, k)).reshape(-1)
            assert got.dtype == dt and got.shape == ref.shape and np.array_equal(got, ref), (r, k)
        assert np.array_equal(slab.owned_centers(), lat.cell_centers()[want.owned_global])
    info = slab.info()
    for k, v in info.items():
        assert int(getattr(part.info, k)) == v, k


def _tiny_face_list(n, pairs, bcells, seed=0):
    rng = np.random.default_rng(seed)

    class M:
        pass
    m = M()
    m.n_cells = n
    m.face_cell = np.array(pairs, np.int32).reshape(-1, 2)
    F, B = m.face_cell.shape[0], len(bcells)
    m.face_area, m.face_dist, m.cell_vol = rng.uniform(.5, 1.5, F), rng.uniform(.5, 1.5, F), rng.uniform(.5, 1.5, n)
    m.bface_cell = np.array(bcells, np.int32)
    m.bface_area, m.bface_dist = rng.uniform(.5, 1.5, B), rng.uniform(.5, 1.5, B)
    return m


def test_degenerate_graphs_disconnected_isolated_duplicate_faces():
    """Host integer work on graphs a real mesh reader can produce by accident: several components, isolated cells,
    no faces at all, two faces between the same pair of cells, as many parts as cells."""
    src = _tiny_face_list(10, [(0, 1), (1, 2), (5, 6), (8, 9)], [0, 3, 3, 9])
    h = Mesh.from_faces(src)
    perm = h.renumber_rcm()
    assert np.array_equal(perm, mo.rcm(10, src.face_cell)) and sorted(perm.tolist()) == list(range(10))
    gd = as_dict(h)
    for method in (capi.PART_SLAB, capi.PART_METIS):
        for k in (1, 2, 3):
            P = Partition(h, k, method)
            assert np.bincount(P.part, minlength=k).min() > 0
            for r in range(k):
                L, want = P.local(r), mo.local_maps(gd, P.part, r, k)
                for key in LOCAL_KEYS:
                    assert np.array_equal(np.asarray(getattr(L, key)), want[key]), (method, k, r, key)
    empty = Mesh.from_faces(_tiny_face_list(4, [], [0, 1, 2, 3]))
    assert sorted(empty.renumber_rcm().tolist()) == [0, 1, 2, 3] and empty.bandwidth == 0
    for method in (capi.PART_SLAB, capi.PART_METIS):
        P = Partition(empty, 2, method)
        assert np.bincount(P.part, minlength=2).min() > 0 and P.info.edge_cut == 0 and P.local(1).n_halo == 0
    dup = Mesh.from_faces(_tiny_face_list(3, [(0, 1), (0, 1), (1, 2)], []))
    dup.renumber_rcm()
    P = Partition(dup, 2, capi.PART_METIS)          # METIS leaves a part empty here: the slab fallback takes over
    assert np.bincount(P.part, minlength=2).min() > 0
    assert sum(P.local(r).n_owned for r in range(2)) == 3
    P = Partition(Mesh.from_faces(_tiny_face_list(3, [(0, 1), (1, 2)], [])), 3, capi.PART_SLAB)
    assert P.part.tolist() == [0, 1, 2]


@pytest.mark.parametrize("seed", range(0, 60, 3))
def test_random_face_lists_renumbering_and_partition_against_the_restatement(seed):
    """Random multigraphs (duplicate faces, isolated cells, several components, random boundary faces): ingestion,
    an arbitrary permutation, RCM and both partitioners against the numpy restatement, every array bit for bit."""
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 60))
    pairs = [(int(a), int(b)) for a, b in rng.integers(0, n, (int(rng.integers(0, 4 * n)), 2)) if a != b]
    src = _tiny_face_list(n, pairs, np.sort(rng.integers(0, n, int(rng.integers(0, 2 * n)))).tolist(), seed)
    before = face_list_dict(src)
    h = Mesh.from_faces(src)
    assert_soa_equal(h, before)
    perm = rng.permutation(n).astype(np.int32)
    h.permute_cells(perm)
    want = mo.permute_face_list(before, perm)
    assert_soa_equal(h, want)
    p2 = h.renumber_rcm()
    assert np.array_equal(p2, mo.rcm(n, want["face_cell"]))
    assert_soa_equal(h, mo.permute_face_list(before, perm[p2]))
    gd = as_dict(h)
    for method in (capi.PART_SLAB, capi.PART_METIS):
        k = int(rng.integers(1, min(5, n) + 1))
        P = Partition(h, k, method)
        for r in range(k):
            L, wl = P.local(r), mo.local_maps(gd, P.part, r, k)
            for key in ("n_owned", "n_interior", "n_halo", "halo_base"):
                assert getattr(L, key) == wl[key], (method, k, r, key)
            for key in LOCAL_KEYS:
                got = np.asarray(getattr(L, key))
                assert got.shape == wl[key].shape and np.array_equal(got, wl[key]), (method, k, r, key)


@pytest.mark.parametrize("kind", ["tet", "hex"])
def test_vtk_dump_follows_the_playground_grammar(tmp_path, kind):
    """sb_mesh_write_vtk against the file grammar of the reference's save_vtk (Playground.cpp:65-109): literal header
    lines, POINTS / CELLS / CELL_TYPES / CELL_DATA blocks, 16 significant digits, fields in the current cell order
    (after a renumbering too)."""
    mesh = Mesh.box(KIND[kind], 3, 2, 2, jitter=0.2, seed_jitter=5, shuffle=True, seed_shuffle=6)
    mesh.renumber_rcm()
    n = mesh.n_cells
    c = mesh.cell_centers()
    fields = {"c": np.sin(c[:, 0]) + 0.1 * np.arange(n), "third": np.full(n, 1.0 / 3.0)}
    path = tmp_path / "out-00001.vtk"
    mesh.write_vtk(str(path), fields)
    lines = path.read_text().split("\n")
    assert lines[:4] == ["# vtk DataFile Version 2.0", "# Generated by Feathers/StormRuler/Mesh2VTK", "ASCII",
                         "DATASET UNSTRUCTURED_GRID"]
    xyz, cells = mo.box_cells(kind, 3, 2, 2, jitter=0.2, seed_jitter=5, shuffle=True, seed_shuffle=6)
    n_nodes, npc = xyz.shape[0], cells.shape[1]
    assert lines[4] == f"POINTS {n_nodes} double"
    pts = np.array([[float(v) for v in ln.split()] for ln in lines[5:5 + n_nodes]])
    assert np.allclose(pts, xyz, rtol=1e-15, atol=0.0)
    k = 5 + n_nodes
    assert lines[k] == "" and lines[k + 1] == f"CELLS {n} {n * (npc + 1)}"
    rows = [ln.split() for ln in lines[k + 2:k + 2 + n]]
    assert all(r[0] == str(npc) and len(r) == npc + 1 for r in rows)
    got_cells = np.array([[int(v) for v in r[1:]] for r in rows])
    # cell k of the dump is the k-th cell of the renumbered mesh: its centre is the mean of its nodes (tets) or close
    ctr = xyz[got_cells].mean(axis=1)
    assert np.abs(ctr - c).max() < (1e-12 if kind == "tet" else 0.05)
    assert sorted(map(tuple, np.sort(got_cells, axis=1).tolist())) == sorted(map(tuple, np.sort(cells, axis=1).tolist()))
    k += 2 + n
    assert lines[k] == "" and lines[k + 1] == f"CELL_TYPES {n}"
    assert set(lines[k + 2:k + 2 + n]) == {"10" if kind == "tet" else "12"}
    k += 2 + n
    assert lines[k] == "" and lines[k + 1] == f"CELL_DATA {n}"
    k += 2
    for name, values in fields.items():
        assert lines[k] == f"SCALARS {name} double 1" and lines[k + 1] == "LOOKUP_TABLE default"
        got = np.array([float(v) for v in lines[k + 2:k + 2 + n]])
        assert np.allclose(got, values, rtol=1e-15, atol=0.0)
        k += 2 + n
    assert lines[k] == ""
    assert "0.3333333333333333" in lines            # digits10 + 1 = 16 significant digits, like the reference
    with pytest.raises(capi.StormB200Error):
        Mesh.from_faces(PolyMesh.bcc(2)).write_vtk(str(tmp_path / "x.vtk"))     # no nodes
    with pytest.raises(capi.StormB200Error):
        mesh.write_vtk(str(tmp_path / "no_such_dir" / "x.vtk"))


REF_MESH_DIR = "/root/reference/tests/_data/mesh"


@pytest.mark.parametrize("name", ["square_nb", "rectangle", "step"])
def test_tetgen_2d_reader_reproduces_the_reference_mesh_classes(tmp_path, name):
    """sb_mesh_read_tetgen_2d on the reference's own test meshes against what the reference's reader +
    UnstructuredMesh<2,2> + views export (golden npz made by oracle/_ref/ref_mesh_tool; `step` is exported live by
    the tool): face order, inner/outer, geometry, boundary-face order -- every array bit for bit. Labels: the file's
    labels, which the reference's label() reports shifted by one at the first face of every label range (lower_bound
    instead of upper_bound, MeshUnstructured.hpp:185-192)."""
    import os
    import subprocess
    prefix = os.path.join(REF_MESH_DIR, name + ".1")
    if not os.path.exists(prefix + ".node"):
        pytest.skip("the reference's test meshes are not mounted here")
    mesh = Mesh.read_tetgen_2d(prefix)
    golden = os.path.join(os.path.dirname(__file__), "golden", f"mesh_{name}.npz")
    if os.path.exists(golden):
        g = dict(np.load(golden))
        centers = None
    else:
        tool = os.path.join(os.path.dirname(os.path.dirname(__file__)), "oracle", "_ref", "ref_mesh_tool")
        if not os.path.exists(tool):
            pytest.skip("oracle/_ref/ref_mesh_tool not built")
        out = tmp_path / "export.bin"
        # trailing dot: the reference swaps the path's extension (path.replace_extension, IoTetgen.hpp:54)
        subprocess.run([tool, "export", prefix + ".", str(out)], check=True, capture_output=True)
        fm, extra = orc.read_mesh_export(str(out))
        g = dict(n_cells=fm.n_cells, face_cell=fm.face_cell, face_area=fm.face_area, face_dist=fm.face_dist,
                 cell_vol=fm.cell_vol, bface_cell=fm.bface_cell, bface_area=fm.bface_area, bface_dist=fm.bface_dist,
                 bface_label=extra["bface_label"])
        centers = extra["cell_center"]
    assert mesh.n_cells == int(g["n_cells"])
    for k in SOA_KEYS:
        got = np.asarray(getattr(mesh, k))
        assert got.shape == g[k].shape and np.array_equal(got, g[k]), f"{k} differs from the reference's export"
    if centers is not None:
        assert np.array_equal(mesh.cell_centers()[:, :2], centers) and not mesh.cell_centers()[:, 2].any()
    labels, ref_labels = mesh.bface_labels(), np.asarray(g["bface_label"])
    assert (np.diff(labels) >= 0).all() and labels.min() >= 1
    starts = np.r_[0, np.flatnonzero(np.diff(labels)) + 1]
    rest = np.setdiff1d(np.arange(len(labels)), starts)
    assert np.array_equal(labels[rest], ref_labels[rest]) and np.array_equal(ref_labels[starts], labels[starts] - 1)
    # normals point from the inner to the outer cell / out of the domain, unit length
    fn, bn = mesh.face_normals()
    c = mesh.cell_centers()
    assert (np.einsum("ij,ij->i", fn, c[mesh.face_cell[:, 1]] - c[mesh.face_cell[:, 0]]) > 0).all()
    assert np.allclose(np.linalg.norm(fn, axis=1), 1.0, rtol=1e-14) and np.allclose(np.linalg.norm(bn, axis=1), 1.0, rtol=1e-14)
    S = np.zeros((mesh.n_cells, 3))
    np.add.at(S, mesh.face_cell[:, 0], mesh.face_area[:, None] * fn)
    np.add.at(S, mesh.face_cell[:, 1], -mesh.face_area[:, None] * fn)
    np.add.at(S, mesh.bface_cell, mesh.bface_area[:, None] * bn)
    assert np.abs(S).max() < 1e-13           # every triangle's edge vectors close
    # and the handle renumbers / partitions like any other (labels travel with the faces)
    mesh.renumber_rcm()
    assert sorted(mesh.bface_labels().tolist()) == sorted(labels.tolist())
    assert Partition(mesh, 3, capi.PART_METIS).info.edge_cut > 0


@pytest.mark.parametrize("name", ["square_nb", "rectangle"])
def test_vtk_of_a_2d_mesh_is_the_playground_file_byte_for_byte(tmp_path, name):
    """The playground's output step (Playground.cpp:205-208 -> save_vtk :65-109) on its own input format: a mesh read by
    sb_mesh_read_tetgen_2d, written by sb_mesh_write_vtk, against the file the reference's mesh classes produce through
    the same statements (oracle/_ref/ref_mesh_tool vtk): node order, per-cell node lists, VTK_TRIANGLE, number format.
    After a renumbering the cells (and their field values) move together."""
    import os
    import subprocess
    prefix = os.path.join(REF_MESH_DIR, name + ".1")
    tool = os.path.join(os.path.dirname(os.path.dirname(__file__)), "oracle", "_ref", "ref_mesh_tool")
    if not os.path.exists(prefix + ".node") or not os.path.exists(tool):
        pytest.skip("needs the reference's test meshes and oracle/_ref/ref_mesh_tool")
    want = tmp_path / "ref.vtk"
    subprocess.run([tool, "vtk", prefix + ".", str(want)], check=True, capture_output=True)
    mesh = Mesh.read_tetgen_2d(prefix)
    c = np.sin(0.37 * np.arange(mesh.n_cells))
    got = tmp_path / "got.vtk"
    mesh.write_vtk(str(got), {"c": c})
    assert got.read_bytes() == want.read_bytes()
    perm = mesh.renumber_rcm()
    mesh.write_vtk(str(got), {"c": c[perm]})
    ref_lines, new_lines = want.read_text().split("\n"), got.read_text().split("\n")
    assert len(ref_lines) == len(new_lines)
    cells0 = ref_lines.index(next(ln for ln in ref_lines if ln.startswith("CELLS "))) + 1
    data0 = ref_lines.index("LOOKUP_TABLE default") + 1
    assert new_lines[:cells0] == ref_lines[:cells0]                                     # header and points unchanged
    for k in (0, 1, mesh.n_cells // 2, mesh.n_cells - 1):
        assert new_lines[cells0 + k] == ref_lines[cells0 + int(perm[k])]
        assert new_lines[data0 + k] == ref_lines[data0 + int(perm[k])]


def test_playground_driver_host_side(tmp_path):
    """scripts/playground_cahn_hilliard.py, the parts that run without a GPU: the Triangle files it can generate are read
    alike by the library and by the reference's own reader + mesh classes (every SoA array, and the VTK file byte for
    byte), and its initial condition is the playground's glibc rand() stream (Playground.cpp:182-184; the committed
    Cahn-Hilliard fixture holds the reference's)."""
    import importlib.util
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(__file__))
    spec = importlib.util.spec_from_file_location("playground_driver", os.path.join(root, "scripts", "playground_cahn_hilliard.py"))
    drv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(drv)
    g = np.load(os.path.join(root, "tests", "golden", "cahn_hilliard_square_nb.npz"))
    assert np.array_equal(drv.initial_condition(len(g["c0"])), g["c0"])
    prefix = str(tmp_path / "gen.1")
    drv.write_triangle_files(prefix, 7, 5)
    mesh = Mesh.read_tetgen_2d(prefix)
    assert (mesh.n_cells, mesh.n_faces, mesh.n_bfaces) == (70, 93, 24)
    assert sorted(set(mesh.bface_labels().tolist())) == [1, 2, 3, 4]
    assert np.isclose(np.asarray(mesh.cell_vol).sum(), 4.0, rtol=1e-14)
    tool = os.path.join(root, "oracle", "_ref", "ref_mesh_tool")
    if not os.path.exists(tool):
        pytest.skip("oracle/_ref/ref_mesh_tool not built")
    subprocess.run([tool, "export", prefix + ".", str(tmp_path / "exp.bin")], check=True, capture_output=True)
    fm, _ = orc.read_mesh_export(str(tmp_path / "exp.bin"))
    for k in SOA_KEYS:
        assert np.array_equal(np.asarray(getattr(mesh, k)), getattr(fm, k)), k
    subprocess.run([tool, "vtk", prefix + ".", str(tmp_path / "ref.vtk")], check=True, capture_output=True)
    mesh.write_vtk(str(tmp_path / "got.vtk"), {"c": np.sin(0.37 * np.arange(mesh.n_cells))})
    assert (tmp_path / "got.vtk").read_bytes() == (tmp_path / "ref.vtk").read_bytes()


def test_tetgen_2d_reader_rejects_malformed_files(tmp_path):
    def write(name, node, edge, ele):
        for ext, text in ((".node", node), (".edge", edge), (".ele", ele)):
            if text is not None:
                (tmp_path / (name + ext)).write_text(text)
        return str(tmp_path / name)
    node = "4 2 0 0\n0 0 0\n1 1 0\n2 1 1\n3 0 1\n"
    edge = "5 1\n0 0 1 1\n1 1 2 1\n2 2 3 2\n3 3 0 2\n4 0 2 0\n"
    ele = "2 3 0\n0 0 1 2\n1 0 2 3\n"
    ok = Mesh.read_tetgen_2d(write("ok", node, edge, ele))
    assert (ok.n_cells, ok.n_faces, ok.n_bfaces) == (2, 1, 4) and ok.bface_labels().tolist() == [1, 1, 2, 2]
    assert ok.face_cell.tolist() == [[0, 1]] and np.array_equal(ok.cell_vol, [0.5, 0.5])
    # edges the file does not list are created by the cells (find_or_insert), with label 0
    part = Mesh.read_tetgen_2d(write("partial", node, "4 1\n0 0 1 1\n1 1 2 1\n2 2 3 2\n3 3 0 2\n", ele))
    assert (part.n_faces, part.n_bfaces) == (1, 4)
    bad = {
        "missing_edge_file": (node, None, ele),
        "dim3": ("4 3 0 0\n0 0 0 0\n1 1 0 0\n2 1 1 0\n3 0 1 0\n", edge, ele),
        "node_out_of_range": (node, edge, "2 3 0\n0 0 1 7\n1 0 2 3\n"),
        "clockwise_second_cell": (node, edge, "2 3 0\n0 0 1 2\n1 0 3 2\n"),      # both cells see edge (0,2) the same way
        "interior_label_on_boundary": (node, "5 1\n0 0 1 0\n1 1 2 1\n2 2 3 2\n3 3 0 2\n4 0 2 0\n", ele),
        "duplicate_edge": (node, "6 1\n0 0 1 1\n1 1 2 1\n2 2 3 2\n3 3 0 2\n4 0 2 0\n5 2 0 0\n", ele),
        "truncated": (node, edge, "2 3 0\n0 0 1\n"),
    }
    for name, (n_, e_, c_) in bad.items():
        with pytest.raises(capi.StormB200Error):
            Mesh.read_tetgen_2d(write(name, n_, e_, c_))
