"""Every kernel family on the GPU, not only the default one: the same checks run with
  wide    : plain 16-bit ids + shared-atomic transposes (meshes beyond the packed-format limits)
  nofans  : packed rank-scatter transposes, no one-ring fans (non-manifold / inconsistently oriented input)
  persist : the persistent software-pipelined kernels (rxm_persistent.cuh)
so that a regression in a fallback path cannot hide behind the fast path."""
import numpy as np
import pytest

import rxmesh_b200 as rx
from conftest import make_mesh
from oracle import oracle as O
from rxmesh_b200.mesh import _DST, _SRC

pytestmark = pytest.mark.gpu

MODES = {"wide": {"RXM_FORCE_WIDE": "1", "RXM_NO_FANS": "1"}, "nofans": {"RXM_NO_FANS": "1"},
         "persist": {"RXM_PERSIST": "1"}, "persist_nofans": {"RXM_PERSIST": "1", "RXM_NO_FANS": "1"}}


@pytest.fixture(params=[(m, k) for m in ("dragon", "bunnyhead", "grid40x31") for k in MODES])
def built(request, monkeypatch):
    name, mode = request.param
    for k, v in MODES[mode].items():
        monkeypatch.setenv(k, v)
    rx.rx_init(0)
    V, F = make_mesh(name)
    m = rx.RXMeshStatic(F, patch_size=256)
    assert m.is_packed() == (mode != "wide") and m.has_fans() == (mode == "persist")
    return name, mode, V, F, m, O.Topology(F)


def test_all_queries(built):
    name, mode, V, F, m, T = built
    for op in ("VV", "VE", "VF", "EV", "EF", "FV", "FE", "FF"):
        o = rx.Op[op]
        off, val = m.query_csr(o)
        s2g_src, s2g_dst = m.slot_to_global(_SRC[o]), m.slot_to_global(_DST[o])
        src = np.repeat(s2g_src.astype(np.uint64), np.diff(off.astype(np.int64)))
        got = np.sort((src << np.uint64(32)) | s2g_dst[val].astype(np.uint64))
        roff, rval = T.query(op)
        rs = np.repeat(np.arange(roff.shape[0] - 1, dtype=np.uint64), np.diff(roff.astype(np.int64)))
        want = np.sort((rs << np.uint64(32)) | rval.astype(np.uint64))
        assert np.array_equal(got, want), (mode, op)
        # consume variant
        vals = np.random.RandomState(1).rand(m._num(_DST[o])).astype(np.float32)
        ref = O.consume_sum((roff, rval), vals)
        assert np.abs(m.query_consume_host(o, vals) - ref).max() < 1e-5 * max(1.0, np.abs(ref).max()), (mode, op)


def test_apps(built):
    name, mode, V, F, m, T = built
    got = m.vertex_normals_host(V)
    ref = O.vertex_normals(F, V, np.float64)
    assert (np.linalg.norm(got - ref, axis=1) / np.linalg.norm(ref, axis=1)).max() < 1e-5, mode
    y = m.laplacian_smooth_host(V, 0.01, 3)
    r = V.astype(np.float64)
    for _ in range(3):
        r = O.laplacian_step(T.query("VV"), r, 0.01, np.float64)
    assert np.abs(y - r).max() < 3e-5 * np.abs(V).max(), mode
    flag = rx.Attribute(m, 0, np.uint32, 1, rx.LOCATION_ALL, rx.AoS)
    m.boundary_vertices(flag)
    assert np.array_equal(flag.to_global().reshape(-1).astype(bool), T.boundary_vertices()[1])


@pytest.mark.parametrize("mesh,tiles", [("grid200x150", True), ("grid200x150", False), ("dragon", False), ("torus64x48", False)])
def test_pipelined_host_calls_match_plain(monkeypatch, mesh, tiles):
    """The host-buffer entry points run as a chunked H2D / kernel / D2H pipeline (rxm_capi.cu: pipelined_host_call)
    whenever the mesh has enough patches.  They must return exactly what the plain upload -> kernel -> download path
    (RXM_PIPE_CHUNKS=0) returns, for tile-ordered patches (real overlap) and for Lloyd patches (degenerate frontiers)."""
    from rxmesh_b200 import meshio
    rx.rx_init(0)
    V, F = make_mesh(mesh)
    fp = meshio.grid_face_tiles(200, 150, 8) if tiles else None
    rng = np.random.RandomState(3)
    res = {}
    for chunks in ("0", "8", "5"):
        monkeypatch.setenv("RXM_PIPE_CHUNKS", chunks)
        m = rx.RXMeshStatic(F, face_patch=fp, patch_size=128)
        assert m.get_num_patches() >= 16
        sv = rng.rand(m.get_num_vertices()).astype(np.float32) if "sv" not in res else res["sv"]
        sf = rng.rand(m.get_num_faces()).astype(np.float32) if "sf" not in res else res["sf"]
        res.setdefault("sv", sv), res.setdefault("sf", sf)
        got = (m.vertex_normals_host(V), m.query_consume_host(rx.Op.VV, sv), m.query_consume_host(rx.Op.VF, sf),
               m.query_consume_host(rx.Op.FV, sv), m.query_consume_host(rx.Op.EF, sf))
        # twice on the same mesh: staging buffers and events are reused
        again = m.vertex_normals_host(V)
        assert np.array_equal(got[0], again)
        if "plain" not in res:
            res["plain"] = got
        else:
            for a, b in zip(res["plain"], got):
                assert np.array_equal(a, b), chunks
    ref = O.vertex_normals(F, V, np.float64)
    err = np.linalg.norm(res["plain"][0] - ref, axis=1) / np.maximum(np.linalg.norm(ref, axis=1), 1e-30)
    assert err.max() < 1e-5
