"""All eight static queries on the GPU, through the C ABI, against the oracle.

Verifier semantics of tests/RXMesh_test/rxmesh_test.h:343-440 for every OWNED source element:
input(h) == h; every valid output handle is owned by the patch it names; its global id is in the
ground truth; counts match (with multiplicity); every ground-truth neighbour is present.
FV / FE / EV additionally keep the reference's order.
"""
import numpy as np
import pytest

import rxmesh_b200 as rx
from conftest import make_mesh
from oracle import oracle as O

pytestmark = pytest.mark.gpu

MESHES = ["sphere3", "dragon", "cube", "bunnyhead", "plane", "diamond", "sphere1", "torus", "ico12",
          "grid40x31"]
OPS = ["VV", "VE", "VF", "EV", "EF", "FV", "FE", "FF"]


@pytest.fixture(scope="module", params=MESHES)
def built(request):
    rx.rx_init(0)
    V, F = make_mesh(request.param)
    m = rx.RXMeshStatic(F, patch_size=512 if F.shape[0] > 600 else 64)
    return request.param, V, F, m, O.Topology(F)


def verify(m, op, inp, out, src, dst, csr, ordered):
    off, val = csr
    n_src = m._num(src)
    ih = inp.host_array()
    oh = out.host_array()
    W = out.num_attributes
    sb = m.slot_base(src).astype(np.int64)
    lb = m.lin_base(src).astype(np.int64)
    s2g = m.slot_to_global(src)
    P = m.get_num_patches()
    checked = 0
    for p in range(P):
        b, cap, no = int(sb[p]), int(sb[p + 1] - sb[p]), int(lb[p + 1] - lb[p])
        lids = np.arange(no)
        handles = ih[b + lids]  # 1 attribute: AoSoA == plain
        assert np.array_equal(handles, (np.uint64(p) << np.uint64(32)) | lids.astype(np.uint64)), (op, p)
        # AoSoA: out[b*W + a*cap + lid]
        rows = oh[b * W:b * W + W * cap].reshape(W, cap)[:, :no].T  # [no, W]
        valid = rows != np.uint64(rx.INVALID64)
        # owned-by-the-named-patch check: local id < owned count of that patch
        hp = (rows >> np.uint64(32)).astype(np.int64)
        hl = (rows & np.uint64(0xFFFF)).astype(np.int64)
        dlb = m.lin_base(dst).astype(np.int64)
        owned_cnt = dlb[1:] - dlb[:-1]
        assert np.all(hp[valid] < P)
        assert np.all(hl[valid] < owned_cnt[hp[valid]]), (op, p, "output handle not owned")
        g_out = m.map_to_global(dst, rows)
        gsrc = s2g[b:b + no]
        for i in range(no):
            want = val[off[gsrc[i]]:off[gsrc[i] + 1]]
            got = g_out[i][valid[i]]
            assert valid[i].sum() == want.shape[0], (op, p, i, got, want)
            if ordered:
                assert np.array_equal(got, want), (op, p, i, got, want)
            else:
                assert np.array_equal(np.sort(got), np.sort(want)), (op, p, i, got, want)
            # valid entries are a prefix of the row (iter[0..size))
            assert valid[i][:want.shape[0]].all()
        checked += no
    assert checked == n_src


@pytest.mark.parametrize("op", OPS)
def test_query(built, op):
    name, V, F, m, T = built
    inp, out, src, dst = m.query_global(rx.Op[op])
    ordered = op in ("EV", "FV", "FE") or (op == "FF" and m.is_edge_manifold())
    csr = T.query(op)
    if op == "FF" and ordered:
        # reference order on manifold input: neighbour across edge 0, 1, 2, boundary edges skipped
        ef = T.query("EF")
        vals, offs = [], [0]
        for f in range(T.nf):
            for e in T.fe[f]:
                vals += [g for g in ef[1][ef[0][e]:ef[0][e + 1]] if g != f]
            offs.append(len(vals))
        csr = (np.asarray(offs, np.uint32), np.asarray(vals, np.uint32))
    verify(m, op, inp, out, src, dst, csr, ordered)


@pytest.mark.parametrize("op", ["VV", "EF", "FV", "FF"])
def test_query_output_layouts(built, op):
    """the query kernels write through Attribute::operator(): handles stored into AoS and SoA (tensor layout, by linear
    id) attributes are the ones stored into the default AoSoA attributes, element by element"""
    name, V, F, m, T = built
    ref_in, ref_out, src, dst = m.query_global(rx.Op[op])
    sb, lb = m.slot_base(src), m.lin_base(src)
    W = ref_out.num_attributes
    for layout in (rx.AoS, rx.SoA):
        inp, out, _, _ = m.query_global(rx.Op[op], layout=layout)
        a, b, ra, rb = inp.host_array(), out.host_array(), ref_in.host_array(), ref_out.host_array()
        if layout == rx.SoA:
            assert out.count() == W * m._num(src)
        for p in range(0, m.get_num_patches(), 3):
            for lid in range(0, int(lb[p + 1] - lb[p]), 5):
                assert a[inp.index(p, lid, 0)] == ra[ref_in.index(p, lid, 0)]
                assert [b[out.index(p, lid, k)] for k in range(W)] == [rb[ref_out.index(p, lid, k)] for k in range(W)]


def test_launch_box(built):
    name, V, F, m, T = built
    blocks, threads, smem = m.launch_box(rx.Op.VV)
    assert blocks == m.get_num_patches() and threads == 256 and 0 < smem < 227 * 1024


EDGE4_MESHES = ["sphere3", "torus", "dragon", "bunnyhead", "plane_5", "ico12", "grid40x31"]


@pytest.mark.parametrize("name", EDGE4_MESHES)
@pytest.mark.parametrize("op", ["EVDiamond", "EE"])
def test_edge4_queries(name, op):
    """Op::EVDiamond / Op::EE (kernels/rxmesh_queries.cuh:198-342; test_ev_diamond.cu): fixed width 4, invalid
    handles on mesh boundaries, exact slot order."""
    rx.rx_init(0)
    V, F = make_mesh(name)
    m = rx.RXMeshStatic(F, patch_size=512 if F.shape[0] > 600 else 64)
    T = O.Topology(F)
    want = T.ev_diamond() if op == "EVDiamond" else T.ee()
    inp, out, src, dst = m.query_global(rx.Op[op])
    assert out.num_attributes == 4
    oh = out.host_array()
    sb, lb = m.slot_base(1).astype(np.int64), m.lin_base(1).astype(np.int64)
    s2g = m.slot_to_global(1)
    seen = 0
    for p in range(m.get_num_patches()):
        b, cap, no = int(sb[p]), int(sb[p + 1] - sb[p]), int(lb[p + 1] - lb[p])
        rows = m.map_to_global(dst, oh[b * 4:b * 4 + 4 * cap].reshape(4, cap)[:, :no].T)
        assert np.array_equal(rows, want[s2g[b:b + no]]), (op, p)
        seen += no
    assert seen == T.ne
    if op == "EVDiamond" and name == "plane_5":
        # the reference's own check (test_ev_diamond.cu:62-76): the two triangles of an interior diamond tile a unit quad
        full = want[(want != 0xFFFFFFFF).all(axis=1)]
        x = V.astype(np.float64)
        area = lambda a, b, c: 0.5 * np.linalg.norm(np.cross(x[b] - x[a], x[c] - x[a]), axis=1)
        assert np.allclose(area(full[:, 0], full[:, 1], full[:, 2]) + area(full[:, 0], full[:, 2], full[:, 3]), 1.0, atol=1e-5)


def test_edge4_rejects_non_manifold():
    rx.rx_init(0)
    F = np.array([[0, 1, 2], [0, 1, 3], [0, 1, 4]], np.uint32)  # three faces on edge (0, 1)
    m = rx.RXMeshStatic(F, patch_size=64)
    with pytest.raises(rx.RXMeshError):
        m.query_global(rx.Op.EVDiamond)


def test_ee_inconsistent_orientation():
    """Op::EE when neighbouring faces disagree on the winding (both traverse a shared edge the same way): the second face
    takes the other side (the reference's atomicCAS fallback, rxmesh_queries.cuh:318-337), so every edge still reports
    the {next, previous} pair of each of its faces -- which side holds which face is not defined."""
    rx.rx_init(0)
    V, F = make_mesh("sphere3")
    F = F.copy()
    F[::3] = F[::3][:, [0, 2, 1]]  # flip every third face
    m = rx.RXMeshStatic(F, patch_size=256)
    T = O.Topology(F)
    inp, out, src, dst = m.query_global(rx.Op.EE)
    oh = out.host_array()
    sb, lb = m.slot_base(1).astype(np.int64), m.lin_base(1).astype(np.int64)
    s2g = m.slot_to_global(1)
    want = {}
    for f in range(T.nf):
        for j in range(3):
            want.setdefault(int(T.fe[f, j]), []).append((int(T.fe[f, (j + 1) % 3]), int(T.fe[f, (j + 2) % 3])))
    for p in range(m.get_num_patches()):
        b, cap, no = int(sb[p]), int(sb[p + 1] - sb[p]), int(lb[p + 1] - lb[p])
        rows = m.map_to_global(dst, oh[b * 4:b * 4 + 4 * cap].reshape(4, cap)[:, :no].T)
        for i in range(no):
            got = sorted((int(rows[i][2 * k]), int(rows[i][2 * k + 1])) for k in range(2) if rows[i][2 * k] != 0xFFFFFFFF)
            assert got == sorted(want[int(s2g[b + i])]), (p, i, got, want[int(s2g[b + i])])
