"""compute-sanitizer workload for the MCF solver (rxm_mcf.cu): both Laplacians, several patches, a few iterations.

  compute-sanitizer --tool memcheck  python scripts/sanitize_mcf.py
  compute-sanitizer --tool racecheck python scripts/sanitize_mcf.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rxmesh_b200 as rx  # noqa: E402
from conftest import make_mesh  # noqa: E402

rx.rx_init(0)
for name, ps in (("sphere3", 64), ("ico10", 128)):
    V, F = make_mesh(name)
    V = np.ascontiguousarray(V, np.float32)
    m = rx.RXMeshStatic(F, patch_size=ps)
    x0 = rx.Attribute(m, 0, np.float32, 3, rx.LOCATION_ALL, rx.AoS)
    x = rx.Attribute(m, 0, np.float32, 3, rx.LOCATION_ALL, rx.AoS)
    x0.from_global(V)
    for uniform, dt in ((True, 10.0), (False, 1e-2)):
        for pc in (False, True):
            info = m.mcf_solve(x0, x, time_step=dt, use_uniform_laplace=uniform, max_iter=12, tol_abs=1e-6, tol_rel=0.0, precondition=pc)
            assert np.isfinite(x.to_global()).all()
            print(name, "uniform" if uniform else "cotangent", "pcg" if pc else "cg", info, flush=True)
print("mcf ok", flush=True)
