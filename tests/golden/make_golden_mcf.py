"""Generate tests/golden/ref_mcf.npz: the MCF solves of the REFERENCE'S OWN kernels (apps/MCF/mcf_kernels.cuh: init_B,
matvec, precond_matvec, compiled unmodified into oracle/_ref/libshim_refsrc1.so) under the drop-in CGMatFreeAttrSolver /
PCGMatFreeAttrSolver headers, run ON THE B200 BOX:

  gpurun -- 'python tests/golden/make_golden_mcf.py'        # writes gpurun_out/ref_mcf.npz
  cp gpurun_out/ref_mcf.npz tests/golden/                   # then commit

tests/test_mcf.py::test_oracle_vs_reference_kernels_golden pins the float64 oracle solve to these on the CPU.  Per mesh (sphere3,
torus40x30, dragon), Laplacian (uniform: dt 10, tol_abs 1e-6 -- the app's defaults; cotangent: dt 1e-2, tol_rel 1e-9) and solver
(cg / pcg; pcg always with tol_rel 1e-9): X [V, 3] fp32 in input vertex order, iterations, start and final residual.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rxmesh_b200 as rx  # noqa: E402
from conftest import make_mesh  # noqa: E402

CASES = [(name, uni, pcg) for name in ("sphere3", "torus40x30", "dragon") for uni in (1, 0) for pcg in (0, 1)]


def params(uni, pcg):
    dt, ta, tr, mi = (10.0, 1e-6, 0.0, 200) if uni else (1e-2, 0.0, 1e-9, 500)
    if pcg:
        ta, tr = 0.0, 1e-9
    return dt, ta, tr, mi


if __name__ == "__main__":
    rx.rx_init(0)
    rx.lib()
    shim = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libshim_refsrc1.so"))
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    out = {}
    for name, uni, pcg in CASES:
        V, F = make_mesh(name)
        V = np.ascontiguousarray(V, np.float32)
        dt, ta, tr, mi = params(uni, pcg)
        X, info = np.zeros_like(V), np.zeros(4, np.float32)
        rc = shim.shim_mcf_cg(p(F), F.shape[0], p(V), V.shape[0], 512, C.c_float(dt), uni, pcg, mi, C.c_float(ta), C.c_float(tr), p(X), p(info))
        assert rc == 0, (name, uni, pcg, rc)
        key = "%s_%s_%s" % (name, "uniform" if uni else "cotangent", "pcg" if pcg else "cg")
        out[key + "_X"] = X
        out[key + "_info"] = info[:3].astype(np.float64)
        print(key, info, flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "ref_mcf.npz"), **out)
    print("wrote gpurun_out/ref_mcf.npz")
