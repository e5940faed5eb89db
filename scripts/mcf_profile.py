"""MCF solve on the 10 M-face noisy torus of config 3, both Laplacians, for ncu launch lists / captures:

  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_mcf --csv --log-file out.csv python scripts/mcf_profile.py
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rxmesh_b200 as rx  # noqa: E402
from rxmesh_b200 import meshio  # noqa: E402

nu = int(os.environ.get("MCF_NU", "2236"))
rx.rx_init(0)
V, F = meshio.torus(nu, nu, noise=0.2)
q = np.arange(F.shape[0] // 2, dtype=np.uint32)
fp = (q // nu // 16 * ((nu + 31) // 32) + q % nu // 32).repeat(2)
m = rx.RXMeshStatic(F, face_patch=fp, patch_size=1024, num_threads=os.cpu_count() or 8)
x0 = rx.Attribute(m, 0, np.float32, 3, rx.DEVICE, rx.AoS)
x = rx.Attribute(m, 0, np.float32, 3, rx.DEVICE, rx.AoS)
x0.from_global(np.ascontiguousarray(V, np.float32))
for uniform, dt, ta, tr in ((True, 10.0, 1e-6, 0.0), (False, 1e-5, 0.0, 1e-6)):
    for rep in range(2):
        t0 = time.perf_counter()
        info = m.mcf_solve(x0, x, time_step=dt, use_uniform_laplace=uniform, max_iter=100, tol_abs=ta, tol_rel=tr)
        print("uniform" if uniform else "cotangent", rep, info, "%.1f ms wall" % (1e3 * (time.perf_counter() - t0)), flush=True)
