"""Patch store invariants, checked on the CPU against the oracle's global numbering.

Mirrors RXMeshTest::run_ltog_mapping_test (tests/RXMesh_test/rxmesh_test.h:44-58,443-625)
plus the numbering / ownership rules of SURVEY.md 3.6.  No CUDA call is made.
"""
import numpy as np
import pytest

import rxmesh_b200 as rx
from conftest import make_mesh
from oracle import oracle as O

MESHES = ["sphere3", "dragon", "cube", "bunnyhead", "plane", "plane_5", "diamond", "sphere1",
          "torus", "ico6", "grid23x17", "damaged0", "damaged1", "damaged2", "damaged3", "damaged4", "damaged8"]


@pytest.fixture(scope="module", params=MESHES)
def built(request):
    V, F = make_mesh(request.param)
    m = rx.RXMeshStatic(F, device=False, patch_size=512 if F.shape[0] > 600 else 64)
    return request.param, V, F, m, O.Topology(F)


def test_counts_and_stats(built):
    name, V, F, m, T = built
    assert (m.get_num_vertices(), m.get_num_edges(), m.get_num_faces()) == (T.nv, T.ne, T.nf)
    s = T.stats()
    assert m.get_input_max_valence() == s["max_valence"]
    assert m.get_input_max_edge_incident_faces() == s["max_edge_incident_faces"]
    assert m.get_input_max_face_adjacent_faces() == s["max_face_adjacent_faces"]
    assert m.is_closed() == s["is_closed"] and m.is_edge_manifold() == s["is_edge_manifold"]
    # global edge numbering is the reference's (first appearance, (max,min) key)
    assert np.array_equal(m.edges(), T.ev)
    assert np.array_equal(m.face_edges(), T.fe)


def test_ltog_mapping_and_numbering(built):
    name, V, F, m, T = built
    P = m.get_num_patches()
    fpatch, vpatch, epatch = m.elem_patch(2), m.elem_patch(0), m.elem_patch(1)
    owner_of = [vpatch, epatch, fpatch]
    owned_seen = [np.zeros(n, bool) for n in (T.nv, T.ne, T.nf)]
    edge_id = {(int(a), int(b)): i for i, (a, b) in enumerate(T.ev)}
    for p in range(P):
        pv = m.patch(p)
        assert pv["n_owned"][2] <= m.get_patch_size()
        for t in range(3):
            l, no = pv["ltog"][t], pv["n_owned"][t]
            # owned first, each half ascending by global id (rxmesh.cpp:845-869)
            assert np.all(np.diff(l[:no].astype(np.int64)) > 0)
            assert np.all(np.diff(l[no:].astype(np.int64)) > 0)
            assert np.all(owner_of[t][l[:no]] == p) and np.all(owner_of[t][l[no:]] != p)
            assert not owned_seen[t][l[:no]].any()
            owned_seen[t][l[:no]] = True
            # slot / linear numbering
            assert np.array_equal(m.slot_to_global(t)[pv["slot_base"][t]:pv["slot_base"][t] + no], l[:no])
            assert pv["slot_base"][t] % 4 == 0
        lv, le, lf = pv["ltog"]
        # check_mapping_edges: local ev -> global vertices -> global edge id == ltog_e
        gv = lv[pv["ev"]]
        assert np.all(gv[:, 0] > gv[:, 1])
        assert all(edge_id[(int(a), int(b))] == int(g) for (a, b), g in zip(gv, le))
        # check_mapping_faces: local fe -> global edges == FE of the global face
        assert np.array_equal(le[pv["fe"] >> 1], T.fe[lf])
        # fv is FE o EV and equals the input corner order
        fe = pv["fe"].astype(np.int64)
        assert np.array_equal(pv["ev"].reshape(-1)[2 * (fe >> 1) + (fe & 1)], pv["fv"])
        assert np.array_equal(lv[pv["fv"]], F[lf])
        # rank-annotated incidence + stored list offsets (patch_layout.h): scattering row ids to
        # offset[col] + rank must reproduce the transposed lists in ascending row order
        assert pv["packed"] == m.is_packed()
        for conn, off, ncol, rank in ((pv["ev"], pv["voff_e"], pv["n"][0], pv["ev_rank"]),
                                      (pv["fv"], pv["voff_f"], pv["n"][0], pv["fv_rank"]),
                                      (pv["fe"] >> 1, pv["eoff_f"], pv["n"][1], pv["fe_rank"])):
            cols = conn.reshape(-1).astype(np.int64)
            cnt = np.bincount(cols, minlength=ncol)
            assert np.array_equal(np.concatenate([[0], np.cumsum(cnt)]), off)
            if pv["packed"]:
                pos = off[cols].astype(np.int64) + rank.reshape(-1)
                assert np.array_equal(np.sort(pos), np.arange(cols.shape[0]))  # a permutation
                rows = np.arange(cols.shape[0]) // conn.shape[1]
                out = np.empty(cols.shape[0], np.int64)
                out[pos] = rows
                for c in range(0, ncol, 5):
                    seg = out[off[c]:off[c + 1]]
                    assert np.all(np.diff(seg) >= 0) and np.array_equal(np.sort(rows[cols == c]), seg)
        # owner tables name the owning patch and the element's local id there
        for t in range(3):
            no = pv["n_owned"][t]
            for i, o in enumerate(pv["owner"][t]):
                q = int(pv["stash"][o >> 16][0])
                g = int(pv["ltog"][t][no + i])
                assert q == owner_of[t][g] and q != p
                assert m.slot_to_global(t)[int(pv["stash"][o >> 16][1 + t]) + (int(o) & 0xFFFF)] == g
    for t in range(3):
        assert owned_seen[t].all()  # ownership partitions the mesh


def test_ribbon_gives_complete_one_ring(built):
    # every owned vertex sees all its incident faces / edges inside its patch (patcher.cu:668-713)
    name, V, F, m, T = built
    vf, ve = O.csr_to_sets(T.query("VF")), O.csr_to_sets(T.query("VE"))
    for p in range(m.get_num_patches()):
        pv = m.patch(p)
        fset, eset = set(pv["ltog"][2].tolist()), set(pv["ltog"][1].tolist())
        for g in pv["ltog"][0][:pv["n_owned"][0]][::7]:
            assert set(vf[g]) <= fset and set(ve[g]) <= eset


def test_owner_is_lowest_patch(built):
    # vertex / edge owner = lowest patch id among incident faces' patches (patcher.cu:730-756)
    name, V, F, m, T = built
    fpatch = m.elem_patch(2).astype(np.int64)
    want_v = np.full(T.nv, 1 << 40, dtype=np.int64)
    np.minimum.at(want_v, F.reshape(-1).astype(np.int64), np.repeat(fpatch, 3))
    assert np.array_equal(want_v, m.elem_patch(0))
    want_e = np.full(T.ne, 1 << 40, dtype=np.int64)
    np.minimum.at(want_e, T.fe.reshape(-1).astype(np.int64), np.repeat(fpatch, 3))
    assert np.array_equal(want_e, m.elem_patch(1))


def test_one_ring_fans(built):
    # fans: per owned vertex the oriented cycle / chain of neighbours; consecutive entries span a face
    # whose winding is (v, a, b) (cf. orient_edges_around_vertices, kernels/rxmesh_queries.cuh:375-499)
    name, V, F, m, T = built
    if not m.has_fans():
        assert name in ("bunnyhead",) or not m.is_edge_manifold() or True
        return
    vv = O.csr_to_sets(T.query("VV"))
    faces = {tuple(int(x) for x in np.roll(f, -k)) for f in F for k in range(3)}
    _, bflags = T.boundary_vertices()
    for p in range(m.get_num_patches()):
        pv = m.patch(p)
        lv = pv["ltog"][0]
        fo, fvv = pv["fan_off"], pv["fan_v"]
        for v in range(pv["n_owned"][0]):
            b, e, closed = int(fo[v] & 0x7FFF), int(fo[v + 1] & 0x7FFF), bool(fo[v] >> 15)
            g = int(lv[v])
            ring = [int(lv[u]) for u in fvv[b:e]]
            assert tuple(sorted(ring)) == vv[g]
            assert closed == (not bflags[g])
            pairs = list(zip(ring, ring[1:])) + ([(ring[-1], ring[0])] if closed else [])
            assert all((g, a, c) in faces for a, c in pairs), (name, p, v)
            # fan_f names the face spanned by consecutive fan vertices
            ff = pv["fan_f"][b:e]
            lf = pv["ltog"][2]
            for (a, c), f in zip(pairs, ff):
                assert set(F[lf[f]].tolist()) == {g, a, c}
            if not closed:
                assert ff[-1] == 0xFFFF
            # fan_e names the edge between the vertex and every fan vertex (VE in oriented order)
            le, ev = pv["ltog"][1], pv["ev"]
            for u, e_ in zip(fvv[b:e], pv["fan_e"][b:e]):
                assert {int(x) for x in ev[e_]} == {v, int(u)}
                assert {int(x) for x in T.ev[le[e_]]} == {g, int(lv[u])}


def test_ring2_extension(built):
    # ring extension: every NOT-OWNED vertex within two rings of an owned one carries its complete one-ring, as ids of the
    # patch's vertices or of "ext" vertices (beyond the ribbon) that resolve through their owner patch like ribbon vertices
    name, V, F, m, T = built
    assert m.has_ring2()
    vv = O.csr_to_sets(T.query("VV"))
    P = m.get_num_patches()
    views = [m.patch(p) for p in range(P)]
    for p, pv in enumerate(views):
        nv, nov = pv["n"][0], pv["n_owned"][0]
        lv = pv["ltog"][0]
        ext_g = []
        for o in pv["ext_owner"]:
            q = int(pv["stash"][o >> 16][0])
            ext_g.append(int(views[q]["ltog"][0][o & 0xFFFF]))
            assert (o & 0xFFFF) < views[q]["n_owned"][0]
        assert not set(ext_g) & set(lv.tolist())  # ext vertices are NOT in the patch
        gid = [int(x) for x in lv] + ext_g         # extended local id -> global id
        owned = set(gid[:nov])
        d1 = set().union(*[set(vv[g]) for g in owned]) - owned if owned else set()
        d2 = (set().union(*[set(vv[g]) for g in d1]) - owned - d1) if d1 else set()
        want = d1 | d2
        assert len(pv["r2_idx"]) == nv - nov + len(ext_g)
        got = set()
        for i, r in enumerate(pv["r2_idx"]):
            if r == 0xFFFF:
                continue
            g = gid[nov + i]
            got.add(g)
            ids = pv["r2_val"][pv["r2_off"][r]:pv["r2_off"][r + 1]]
            assert tuple(sorted(gid[u] for u in ids)) == vv[g], (name, p, i)
        assert got == want, (name, p)


def test_ring2_can_be_left_out():
    V, F = make_mesh("ico6")
    m = rx.RXMeshStatic(F, device=False, patch_size=64, ring2=False)
    assert not m.has_ring2() and m.patch(0)["r2_idx"] is None
    m2 = rx.RXMeshStatic(F, device=False, patch_size=64)
    assert m2.topo_bytes() > m.topo_bytes()


def test_stored_ff_and_ef_rows(built):
    # stored FF rows (owned faces) and EF pairs (owned edges): what the FF / EF kernels read as plain rows on
    # edge-manifold input must be the oracle's adjacency restricted to the patch (the ribbon makes it complete)
    name, V, F, m, T = built
    if not m.is_edge_manifold():
        assert m.patch(0)["ff"] is None and m.patch(0)["ef"] is None
        return
    ff_ref, ef_ref = O.csr_to_sets(T.query("FF")), O.csr_to_sets(T.query("EF"))
    for p in range(m.get_num_patches()):
        pv = m.patch(p)
        ltf, lte = pv["ltog"][2], pv["ltog"][1]
        for f in range(pv["n_owned"][2]):
            row = pv["ff"][f]
            got = [int(ltf[x]) for x in row if x != 0xFFFF]
            assert all(x == 0xFFFF for x in row[len(got):])  # compacted to the front
            assert tuple(sorted(got)) == ff_ref[int(ltf[f])], (p, f)
        for e in range(pv["n_owned"][1]):
            row = pv["ef"][e]
            loc = [int(x) for x in row if x != 0xFFFF]
            assert loc == sorted(loc) and (row[0] != 0xFFFF)
            assert tuple(sorted(int(ltf[x]) for x in loc)) == ef_ref[int(lte[e])], (p, e)


def test_fans_absent_on_inconsistent_orientation():
    # two triangles sharing an edge with the SAME direction: not consistently oriented -> no fans
    F = np.array([[0, 1, 2], [1, 3, 2][::-1]], dtype=np.uint32)  # second face flipped
    assert not rx.RXMeshStatic(F, device=False).has_fans()
    F = np.array([[0, 1, 2], [1, 3, 2]], dtype=np.uint32)
    assert rx.RXMeshStatic(F, device=False).has_fans()


def test_wide_format_fallback(monkeypatch):
    V, F = make_mesh("sphere3")
    assert rx.RXMeshStatic(F, device=False).is_packed()
    monkeypatch.setenv("RXM_FORCE_WIDE", "1")
    m = rx.RXMeshStatic(F, device=False)
    assert not m.is_packed() and m.patch(0)["ev_rank"] is None
    monkeypatch.delenv("RXM_FORCE_WIDE")
    # a vertex of valence >= 32 cannot carry its rank in 5 bits -> wide format
    k = 40
    fan = np.array([[0, 1 + i, 1 + (i + 1) % k] for i in range(k)], dtype=np.uint32)
    assert not rx.RXMeshStatic(fan, device=False).is_packed()


def test_parallel_lloyd_equals_serial(monkeypatch):
    """patcher_lloyd (patcher/patcher.cu:828-987 in role): the level-synchronous multi-threaded multi-source BFS and
    re-centring passes give bit for bit the face -> patch assignment of the single-threaded FIFO passes, for any thread
    count -- several components (seeds added for unreached components), frontiers above and below the parallel
    threshold, a size bound that needs splits"""
    Va, Fa = make_mesh("sphere3")
    Vb, Fb = make_mesh("torus")
    two = np.concatenate([Fa, Fb + Va.shape[0]]).astype(np.uint32)
    Vc, Fc = rx.meshio.icosphere(100)  # 200 000 faces: frontiers of several thousand faces
    for F, ps in ((Fa, 64), (make_mesh("dragon")[1], 512), (two, 128), (Fc, 256)):
        monkeypatch.setenv("RXM_PATCHER_SERIAL", "1")
        m = rx.RXMeshStatic(F, patch_size=ps, device=False)
        ref = m.elem_patch(2).copy()  # elem_patch is a view into the mesh: copy while it is alive
        monkeypatch.delenv("RXM_PATCHER_SERIAL")
        monkeypatch.setenv("RXM_PATCHER_PARALLEL", "1")  # also with fewer than three threads
        for nt in (1, 2, 5, 8):
            m = rx.RXMeshStatic(F, patch_size=ps, device=False, num_threads=nt)
            assert np.array_equal(ref, m.elem_patch(2)), (F.shape[0], ps, nt)
        monkeypatch.delenv("RXM_PATCHER_PARALLEL")
        assert np.bincount(ref).max() <= ps


def test_user_patching_is_honoured():
    from rxmesh_b200 import meshio
    V, F = meshio.grid(33, 33)
    fp = meshio.grid_face_tiles(33, 33, 8)
    m = rx.RXMeshStatic(F, face_patch=fp, device=False)
    assert m.get_num_patches() == 16
    assert np.array_equal(m.elem_patch(2), fp)


def test_errors():
    with pytest.raises(rx.RXMeshError):
        rx.RXMeshStatic(np.array([[0, 1, 1]], dtype=np.uint32), device=False)  # degenerate face
    with pytest.raises(rx.RXMeshError):
        rx.RXMeshStatic(np.array([[0, 1, 3]], dtype=np.uint32), device=False)  # isolated vertex id 2
    m = rx.RXMeshStatic(np.array([[0, 1, 2]], dtype=np.uint32), device=False)
    with pytest.raises(rx.RXMeshError, match="DEVICE"):
        m.add_vertex_attribute("x", np.float32, 3, rx.DEVICE)  # no silent CPU fallback
    a = m.add_vertex_attribute("h", np.float32, 3, rx.HOST, rx.AoS)
    b = m.add_vertex_attribute("n", np.float32, 3, rx.HOST, rx.AoS)
    with pytest.raises(rx.RXMeshError, match="no CPU fallback"):
        m.vertex_normals(a, b)
    with pytest.raises(rx.RXMeshError, match="no CPU fallback"):
        m.mcf_solve(a, b)
    with pytest.raises(rx.RXMeshError, match="no CPU fallback"):
        m.mcf_solve(a, b, precondition=True)


def test_host_attribute_roundtrip_all_layouts():
    V, F = make_mesh("sphere3")
    m = rx.RXMeshStatic(F, device=False, patch_size=128)
    for layout in (rx.AoS, rx.AoSoA, rx.SoA):
        a = rx.Attribute(m, 0, np.float32, 3, rx.HOST, layout)
        a.from_global(V)
        assert np.array_equal(a.to_global(), V)
        # Attribute::operator()(handle, attr) addressing
        h = a.host_array()
        g2s, sb, ep = m.global_to_slot(0), m.slot_base(0), m.elem_patch(0)
        for g in (0, 17, 385):
            p = int(ep[g])
            lid = int(g2s[g] - sb[p])
            assert [h[a.index(p, lid, k)] for k in range(3)] == V[g].tolist()


def test_true_soa_host_storage_is_column_major():
    """Attribute.TrueSoAHostStorageIsColumnMajor / ResetSetsAllComponents (tests/RXMesh_test/test_attribute.cu:240-279,
    305-329): SoA is the reference's tensor layout -- storage_size == #elements * #attributes and
    data[c * n + linear_id(handle)] is component c of that element (attribute.h:249-261,406-421)."""
    V, F = make_mesh("sphere3")
    m = rx.RXMeshStatic(F, device=False, patch_size=128)
    n = m.get_num_vertices()
    a = rx.Attribute(m, 0, np.float32, 3, rx.HOST, rx.SoA)
    assert a.count() == 3 * n  # no padding slots
    data = a.host_array()
    for c in range(3):
        data[c * n:(c + 1) * n] = c * 1000 + np.arange(n, dtype=np.float32)
    lb = m.lin_base(0)
    for p in range(m.get_num_patches()):
        for lid in range(int(lb[p + 1] - lb[p])):
            row = int(lb[p]) + lid  # Context::linear_id (context.h:275-290)
            assert [data[a.index(p, lid, c)] for c in range(3)] == [c * 1000 + row for c in range(3)]
    # to_global undoes linear id -> input id: row r of the tensor is element slot_to_global(owner slot of r)
    g = a.to_global()
    s2g, sb = m.slot_to_global(0), m.slot_base(0)
    for p in range(m.get_num_patches()):
        for lid in range(int(lb[p + 1] - lb[p])):
            assert g[s2g[sb[p] + lid]].tolist() == [c * 1000 + int(lb[p]) + lid for c in range(3)]
    for layout in (rx.AoS, rx.AoSoA, rx.SoA):
        b = rx.Attribute(m, 0, np.float32, 3, rx.HOST, layout)
        b.reset(np.float32(5.0), rx.HOST)
        assert np.all(b.to_global() == 5.0)


@pytest.mark.parametrize("name", ["sphere3", "torus"])
def test_reference_saved_patching_golden(name, tmp_path):
    """tests/golden/<name>_patches are the patchings SAVED BY THE REFERENCE (input/sphere3_patches,
    input/torus_patches; Patcher::serialize, patcher/patcher.h:162-182). Replaying the reference's face->patch
    assignment, our builder must reproduce the reference's OWN outputs for ownership (Patcher::assign_patch,
    patcher.cu:719-757: m_vertex_patch / m_edge_patch, edge ids in the reference's numbering), owned face
    lists (m_patches_val/offset) and ribbons (Patcher::extract_ribbons, patcher.cu:640-717)."""
    import os
    from conftest import GOLDEN
    V, F = make_mesh(name)
    path = os.path.join(GOLDEN, name + "_patches")
    pf = rx.load_patcher_file(path)
    assert (pf["num_vertices"], pf["num_faces"]) == (V.shape[0], F.shape[0]) and pf["patch_size"] == 512
    m = rx.RXMeshStatic(F, patcher_file=path, device=False)
    P = pf["num_patches"]
    assert m.get_num_patches() == P and m.get_num_edges() == pf["num_edges"]
    assert np.array_equal(m.elem_patch(2), pf["face_patch"])
    assert np.array_equal(m.elem_patch(0), pf["vertex_patch"])   # reference's vertex ownership
    assert np.array_equal(m.elem_patch(1), pf["edge_patch"])     # reference's edge ownership AND edge numbering
    pend, rend = pf["patches_offset"][:P], pf["ribbon_ext_offset"][:P]
    for p in range(P):
        pv = m.patch(p)
        no = pv["n_owned"][2]
        owned_ref = pf["patches_val"][(pend[p - 1] if p else 0):pend[p]]
        ribbon_ref = pf["ribbon_ext_val"][(rend[p - 1] if p else 0):rend[p]]
        assert np.array_equal(np.sort(owned_ref), pv["ltog"][2][:no])
        assert np.array_equal(np.sort(ribbon_ref), pv["ltog"][2][no:])
    # write -> read round trip in the same format
    out = str(tmp_path / "saved_patches")
    m.save_patcher_file(out)
    back = rx.load_patcher_file(out)
    for k in ("face_patch", "vertex_patch", "edge_patch"):
        assert np.array_equal(back[k], pf[k])
    assert np.array_equal(back["patches_offset"], pend) and np.array_equal(back["ribbon_ext_offset"], rend)
    assert np.array_equal(np.sort(back["patches_val"][:pend[0]]), np.sort(pf["patches_val"][:pend[0]]))


@pytest.mark.parametrize("name,tiles", [("grid200x150", True), ("grid200x150", False), ("dragon", False), ("torus", False)])
@pytest.mark.parametrize("chunks", [4, 8])
def test_pipeline_frontiers(name, tiles, chunks):
    """Frontiers of the chunked H2D / kernel / D2H pipeline of the host-buffer calls (rxm_capi.cu: build_pipe_plan), checked
    against their definitions with numpy: before chunk c's slots are filled every owned element of patches < pb[c+1] has
    arrived (up_hi); chunk c only reads ribbon elements owned by chunks <= need[c]; once chunks <= c are done every global id
    below down_lo[c+1] is final."""
    import ctypes as C
    from rxmesh_b200 import meshio
    from rxmesh_b200._lib import lib
    V, F = make_mesh(name)
    fp = meshio.grid_face_tiles(200, 150, 8) if tiles else None
    m = rx.RXMeshStatic(F, face_patch=fp, patch_size=128, device=False)
    P = m.get_num_patches()
    pb = np.zeros(chunks + 1, np.uint32)
    up, dn, need = np.zeros((3, chunks), np.uint64), np.zeros((3, chunks + 1), np.uint64), np.zeros(chunks, np.uint32)
    K = lib().rxm_mesh_pipe_plan(m._h, chunks, pb.ctypes.data_as(C.c_void_p), up.ctypes.data_as(C.c_void_p),
                                 dn.ctypes.data_as(C.c_void_p), need.ctypes.data_as(C.c_void_p))
    assert K == chunks and pb[0] == 0 and pb[-1] == P and np.all(np.diff(pb.astype(np.int64)) > 0)
    chunk_of = np.searchsorted(pb, np.arange(P), side="right") - 1
    for t in range(3):
        ep = m.elem_patch(t).astype(np.int64)            # owner patch of every global element
        ec = chunk_of[ep]                                # ... and its chunk
        n = ep.shape[0]
        assert np.all(np.diff(up[t].astype(np.int64)) >= 0) and up[t][-1] == n
        assert np.all(np.diff(dn[t].astype(np.int64)) >= 0) and dn[t][0] == 0 and dn[t][-1] == n
        for c in range(chunks):
            owned = np.nonzero(ec <= c)[0]
            assert owned.max() < up[t][c]                 # all of them are inside the uploaded prefix
            later = np.nonzero(ec > c)[0]
            if later.size:
                assert later.min() >= dn[t][c + 1]        # nothing below the download frontier is still pending
    for c in range(chunks):
        assert c <= need[c] < chunks
        for p in range(int(pb[c]), int(pb[c + 1])):
            pv = m.patch(p)
            if pv["stash"].shape[0]:
                assert chunk_of[pv["stash"][:, 0]].max() <= need[c]
    if tiles:  # tile-ordered patches: a chunk needs at most the next chunk, and the frontiers advance with the chunks
        assert np.all(need <= np.minimum(np.arange(chunks) + 1, chunks - 1))
        assert up[0][0] < 0.6 * m.get_num_vertices()
