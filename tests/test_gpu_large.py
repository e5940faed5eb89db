"""BASELINE.json configs[1] at full size: all eight static queries on a 10 M-face subdivided sphere
(class-I icosphere, nu = 707 -> 9 996 980 faces), bit-exact against the oracle.

Size-independent form of the reference verifier (tests/RXMesh_test/rxmesh_test.h:343-440): the multiset of
(source global id, neighbour global id) pairs produced by the GPU must equal the oracle's, which implies
per-element count equality, correctness and completeness; ownership of every output handle is implied by
the slot -> global map being defined for it."""
import numpy as np
import pytest

import rxmesh_b200 as rx
from oracle import oracle as O
from rxmesh_b200 import meshio
from rxmesh_b200.mesh import _DST, _SRC

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sphere10m():
    rx.rx_init(0)
    V, F = meshio.icosphere(707)
    assert F.shape[0] == 9_996_980 and V.shape[0] == 4_998_492
    m = rx.RXMeshStatic(F)  # built-in Lloyd patcher
    return V, F, m, O.Topology(F)


def _pairs(off, val, src_ids, dst_ids):
    src = np.repeat(src_ids.astype(np.uint64), np.diff(off.astype(np.int64)))
    return np.sort((src << np.uint64(32)) | dst_ids[val].astype(np.uint64))


@pytest.mark.parametrize("op", ["VV", "VE", "VF", "EV", "EF", "FV", "FE", "FF"])
def test_all_queries_10m_sphere(sphere10m, op):
    V, F, m, T = sphere10m
    o = rx.Op[op]
    off, val = m.query_csr(o)
    s2g_src, s2g_dst = m.slot_to_global(_SRC[o]), m.slot_to_global(_DST[o])
    assert not np.any(s2g_dst[val] == 0xFFFFFFFF), "an output handle is not an owned element of its patch"
    got = _pairs(off, val, s2g_src, s2g_dst)  # padding slots have empty lists
    roff, rval = T.query(op)
    want = _pairs(roff, rval, np.arange(roff.shape[0] - 1), np.arange(max(T.nv, T.ne, T.nf), dtype=np.uint32))
    assert got.shape == want.shape and np.array_equal(got, want)
    if op in ("EV", "FV", "FE"):  # order preserved: check on a strided sample of slots
        for s in range(0, s2g_src.shape[0], 50021):
            g = s2g_src[s]
            if g != 0xFFFFFFFF:
                assert np.array_equal(s2g_dst[val[off[s]:off[s + 1]]], rval[roff[g]:roff[g + 1]])


def test_normals_and_patch_stats_10m_sphere(sphere10m):
    V, F, m, T = sphere10m
    assert m.get_num_patches() > 19000 and m.get_per_patch_max_faces() < 2048
    sizes = np.diff(m.lin_base(2).astype(np.int64))
    assert sizes.max() <= 512
    got = m.vertex_normals_host(V)
    ref = O.vertex_normals(F, V, np.float64)
    rel = np.linalg.norm(got - ref, axis=1) / np.linalg.norm(ref, axis=1)
    assert rel.max() < 1e-5
    # on the unit sphere the normal is parallel to the position: a size-independent sanity property
    cosang = np.sum(got * V, axis=1) / np.linalg.norm(got, axis=1)
    assert cosang.min() > 0.9999
