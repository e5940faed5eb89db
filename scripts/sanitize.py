"""Workload for compute-sanitizer (the counterpart of the reference's scripts/sanatize.cmd):

  compute-sanitizer --tool memcheck  python scripts/sanitize.py
  compute-sanitizer --tool racecheck python scripts/sanitize.py

Runs every kernel family (fans / packed / wide-atomic / persistent) of the hot path once on a small mesh."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rxmesh_b200 as rx  # noqa: E402
from conftest import make_mesh  # noqa: E402

rx.rx_init(0)
DST = {"V": 0, "E": 1, "F": 2}
ONLY_MULTI = bool(os.environ.get("SANITIZE_MULTI"))  # just the multi-shard part at the end
for mode in () if ONLY_MULTI else ({}, {"RXM_NO_FANS": "1"}, {"RXM_NO_FANS": "1", "RXM_FORCE_WIDE": "1"}, {"RXM_PERSIST": "1"},
                                   {"RXM_VN_SCALAR": "1", "RXM_CONSUME_BT256": "1", "RXM_PIPE_CHUNKS": "0"}):
    for k in ("RXM_NO_FANS", "RXM_FORCE_WIDE", "RXM_PERSIST", "RXM_VN_SCALAR", "RXM_CONSUME_BT256", "RXM_PIPE_CHUNKS"):
        os.environ.pop(k, None)
    os.environ.update(mode)
    V, F = make_mesh("bunnyhead")
    m = rx.RXMeshStatic(F, patch_size=256)
    m.vertex_normals_host(V)
    m.laplacian_smooth_host(V, 0.01, 2)
    for op in ("VV", "VE", "VF", "EV", "EF", "FV", "FE", "FF"):
        m.query_global(rx.Op[op])
        m.query_consume_host(rx.Op[op], np.ones(m._num(DST[op[1]]), np.float32))
    for op in ("EVDiamond", "EE"):  # fixed-width-4 results (bunnyhead is edge-manifold with a boundary)
        m.query_global(rx.Op[op])
    x = rx.Attribute(m, 0, np.float32, 3, rx.LOCATION_ALL, rx.AoS)
    y = rx.Attribute(m, 0, np.float32, 3, rx.LOCATION_ALL, rx.AoS)
    x.from_global(V)
    if not mode:
        m.bilateral_filter(x, y, 2)
        flag = rx.Attribute(m, 0, np.uint32, 1, rx.LOCATION_ALL, rx.AoS)
        m.boundary_vertices(flag)
    print("mode", mode, "ok", flush=True)
# round 2: the Lloyd passes on the GPU (rxm_patcher_gpu.cu) and the single-process multi-GPU mode with three shards on this
# device -- the fused compute + halo kernel (peer stores into ghost slots, flag words, reader check-in)
for k in ("RXM_NO_FANS", "RXM_FORCE_WIDE", "RXM_PERSIST", "RXM_VN_SCALAR", "RXM_CONSUME_BT256", "RXM_PIPE_CHUNKS"):
    os.environ.pop(k, None)
os.environ["RXM_PATCHER_GPU"] = "1"
V, F = make_mesh("bunnyhead")
m = rx.RXMeshStatic(F, patch_size=128)
os.environ["RXM_PATCHER_GPU"] = "0"
m0 = rx.RXMeshStatic(F, device=False, patch_size=128)
assert np.array_equal(m.elem_patch(2), m0.elem_patch(2))
del os.environ["RXM_PATCHER_GPU"]
print("gpu patcher ok", flush=True)
if os.environ.get("SANITIZE_MULTI"):
    # three shards on this device: their fused kernels wait for each other's flags, i.e. they must be able to run
    # CONCURRENTLY -- run this part on its own and under a timeout (a tool that serialises kernels would make it wait forever)
    from rxmesh_b200.multi import RXMeshMulti  # noqa: E402
    mm = RXMeshMulti(F, [0, 0, 0], patch_size=64)
    a = mm.laplacian_smooth(V.astype(np.float32), 0.01, 6)
    assert np.array_equal(a, rx.RXMeshStatic(F, patch_size=64).laplacian_smooth_host(V.astype(np.float32), 0.01, 6))
    mm.vertex_normals(V.astype(np.float32))
    print("multi ok", flush=True)
    sys.exit(0)
# the user-kernel path through the drop-in headers (Query::dispatch, higher_query_block_dispatcher, split API)
if os.environ.get("SANITIZE_SKIP_SHIM"):
    sys.exit(0)
import ctypes as C  # noqa: E402
shim = C.CDLL(os.path.join(ROOT, "tests", "cpp", "libshim_apps.so"))
V, F = make_mesh("bunnyhead")
p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
out = np.zeros_like(V)
assert shim.shim_vertex_normals(p(F), F.shape[0], p(V), V.shape[0], 256, p(out)) == 0
assert shim.shim_filtering(p(F), F.shape[0], p(V), V.shape[0], 512, 1, p(out)) == 0
print("shim ok", flush=True)
