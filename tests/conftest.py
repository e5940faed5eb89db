import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # make sure the native pieces exist (nvcc cross-compiles without a GPU)
    so = os.path.join(ROOT, "rxmesh_b200", "librxmesh_b200.so")
    if not os.path.exists(so) and os.path.exists("/usr/local/cuda/bin/nvcc"):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "rxmesh_b200", "csrc")])


def load_golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: g[k] for k in g.files}


@pytest.fixture(scope="session")
def golden():
    return load_golden


def make_mesh(name):
    """(verts, faces) for a named test mesh: golden OBJ fixtures or procedural ones."""
    from rxmesh_b200 import meshio
    if name.startswith("ico"):
        return meshio.icosphere(int(name[3:]))
    if name.startswith("grid"):
        nx, ny = name[4:].split("x")
        return meshio.grid(int(nx), int(ny))
    if name.startswith("torus") and "x" in name:
        nu, nv = name[5:].split("x")
        return meshio.torus(int(nu), int(nv), noise=0.2)
    g = load_golden(name)
    return g["V"], g["F"]
