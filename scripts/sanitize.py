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
for mode in ({}, {"RXM_NO_FANS": "1"}, {"RXM_NO_FANS": "1", "RXM_FORCE_WIDE": "1"}, {"RXM_PERSIST": "1"}):
    for k in ("RXM_NO_FANS", "RXM_FORCE_WIDE", "RXM_PERSIST"):
        os.environ.pop(k, None)
    os.environ.update(mode)
    V, F = make_mesh("bunnyhead")
    m = rx.RXMeshStatic(F, patch_size=256)
    m.vertex_normals_host(V)
    m.laplacian_smooth_host(V, 0.01, 2)
    for op in ("VV", "VE", "VF", "EV", "EF", "FV", "FE", "FF"):
        m.query_global(rx.Op[op])
        m.query_consume_host(rx.Op[op], np.ones(m._num(DST[op[1]]), np.float32))
    x = rx.Attribute(m, 0, np.float32, 3, rx.LOCATION_ALL, rx.AoS)
    y = rx.Attribute(m, 0, np.float32, 3, rx.LOCATION_ALL, rx.AoS)
    x.from_global(V)
    if not mode:
        m.bilateral_filter(x, y, 2)
        flag = rx.Attribute(m, 0, np.uint32, 1, rx.LOCATION_ALL, rx.AoS)
        m.boundary_vertices(flag)
    print("mode", mode, "ok", flush=True)
