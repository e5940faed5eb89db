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
    if name.startswith("damaged"):
        return damaged_mesh(int(name[7:]))
    g = load_golden(name)
    return g["V"], g["F"]


def damaged_mesh(seed):
    """An icosphere put through what real inputs do (deterministic per seed): a third of the faces removed at random
    (holes, boundary chains, pieces that hang together by one vertex, several components), some faces flipped (inconsistent
    orientation: no one-ring fans), and for odd seeds one extra face glued onto an existing edge (a non-manifold edge with
    three faces: no stored FF / EF rows, EE / EVDiamond unsupported)."""
    from rxmesh_b200 import meshio
    rng = np.random.RandomState(1000 + seed)
    V, F = meshio.icosphere(6)
    keep = rng.rand(F.shape[0]) > (0.33 if seed < 4 else 0.04)  # seeds >= 4: a few holes, orientation kept -> open fans
    F = F[keep].copy()
    flip = rng.rand(F.shape[0]) < (0.1 if seed < 4 else 0.0)
    F[flip] = F[flip][:, [0, 2, 1]]
    if seed % 2:
        a, b = int(F[0, 0]), int(F[0, 1])
        V = np.concatenate([V, (1.3 * V[a:a + 1] + 0.2).astype(V.dtype)])
        F = np.concatenate([F, np.asarray([[a, b, V.shape[0] - 1]], dtype=F.dtype)])
    F = np.ascontiguousarray(F, dtype=np.uint32)
    # vertices no face references any more are dropped (the builder rejects isolated vertex ids, tested in test_errors)
    used = np.unique(F)
    remap = np.full(V.shape[0], -1, np.int64)
    remap[used] = np.arange(used.shape[0])
    F = remap[F].astype(np.uint32)
    V = V[used]
    return np.ascontiguousarray(V, dtype=np.float32), F
