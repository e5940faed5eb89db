"""The MCF solve three ways on one mesh and patching (noisy torus, Lloyd patches): the drop-in CGMatFreeAttrSolver over the
REFERENCE'S OWN init_B / matvec kernels (oracle/_ref/libshim_refsrc1.so: the six-kernel structure of the app), the same solver
over the restated kernel (tests/cpp/libshim_apps.so, cotangent only), and the fixed-function rxm_mcf_solve."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rxmesh_b200 as rx  # noqa: E402
from rxmesh_b200 import meshio  # noqa: E402

nu = int(os.environ.get("MCF_NU", "2236"))
ps = 1024
rx.rx_init(0)
V, F = meshio.torus(nu, nu, noise=0.2)
V = np.ascontiguousarray(V, np.float32)
F = np.ascontiguousarray(F, np.uint32)
p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
rx.lib()
out = {"mesh": "%d-face noisy torus, Lloyd patches of <= %d faces" % (F.shape[0], ps)}
libs = {"reference_kernels_under_dropin_solver": os.path.join(ROOT, "oracle", "_ref", "libshim_refsrc1.so")}
cases = (("uniform", 1, 10.0, 1e-6, 0.0), ("cotangent", 0, 1e-5, 0.0, 1e-6))
for label, path in libs.items():
    if not os.path.exists(path):
        out[label] = "not built"
        continue
    shim = C.CDLL(path)
    for cname, uni, dt, ta, tr in cases:
        res, info = np.zeros_like(V), np.zeros(4, np.float32)
        t0 = time.perf_counter()
        rc = shim.shim_mcf_cg(p(F), F.shape[0], p(V), V.shape[0], ps, C.c_float(dt), uni, 0, 100, C.c_float(ta), C.c_float(tr), p(res), p(info))
        steps = int(info[0]) + 1
        out[label + "_" + cname] = {"rc": rc, "iterations": int(info[0]), "ms_solve": float(info[3]), "ms_per_iteration": float(info[3]) / steps,
                                    "wall_s_with_mesh_build": round(time.perf_counter() - t0, 2)}
        print(label, cname, out[label + "_" + cname], flush=True)
m = rx.RXMeshStatic(F, patch_size=ps, num_threads=os.cpu_count() or 8)
x0 = rx.Attribute(m, 0, np.float32, 3, rx.DEVICE, rx.AoS)
x = rx.Attribute(m, 0, np.float32, 3, rx.DEVICE, rx.AoS)
x0.from_global(V)
import torch  # noqa: E402  (events)
for cname, uni, dt, ta, tr in cases:
    kw = dict(time_step=dt, use_uniform_laplace=bool(uni), max_iter=100, tol_abs=ta, tol_rel=tr)
    m.mcf_solve(x0, x, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    info = m.mcf_solve(x0, x, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    steps = info["iterations"] + (1 if info["converged"] else 0)
    out["fixed_function_" + cname] = {"iterations": info["iterations"], "ms_solve": ms, "ms_per_iteration": ms / max(steps, 1)}
    print("fixed_function", cname, out["fixed_function_" + cname], flush=True)
for cname in ("uniform", "cotangent"):
    a, b = out.get("reference_kernels_under_dropin_solver_" + cname), out.get("fixed_function_" + cname)
    if isinstance(a, dict) and isinstance(b, dict) and a["rc"] == 0:
        out["speedup_" + cname] = a["ms_per_iteration"] / b["ms_per_iteration"]
print(json.dumps(out))
