"""Single-process multi-GPU mode behind the C ABI (rxm_multi_*, rxmesh_b200/csrc/rxm_multi.cu).

CPU: the host-only plan (shards, ghost rings, owner matching) against the global mesh and against the planner of the
one-process-per-GPU mode (rxmesh_b200/distributed.py); locality ordering of the Lloyd patch ids.
GPU: the shards may share a device, so the fused compute + halo kernel (k_laplacian_fan2<true>: peer stores into ghost
slots, flag words, reader check-in) runs on a ONE-GPU box too: results bit-identical to the unsharded mesh.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _mesh(kind):
    from rxmesh_b200 import meshio
    if kind == "torus":
        return meshio.torus(48, 40, noise=0.1)
    if kind == "grid":
        return meshio.grid(70, 53)
    return meshio.icosphere(14)


@pytest.mark.parametrize("kind,ps,n", [("torus", 128, 2), ("grid", 128, 3), ("ico", 128, 4), ("torus", 64, 2)])
def test_host_plan(kind, ps, n):
    import rxmesh_b200 as rx
    from rxmesh_b200 import distributed as D
    from rxmesh_b200.multi import RXMeshMulti
    V, F = _mesh(kind)
    g = rx.RXMeshStatic(F, device=False, patch_size=ps)
    fp = g.elem_patch(2).copy()
    for face_patch in (fp, None):  # caller's patching / the built-in patcher (must be the one RXMeshStatic runs)
        mm = RXMeshMulti(F, n, face_patch=face_patch, patch_size=ps, device=False)
        assert mm.info(0) == n and mm.info(1) == g.get_num_patches()
        assert mm.info(3) == V.shape[0] and mm.info(4) == F.shape[0]
        assert sum(mm.info(10, r) for r in range(n)) == g.get_num_patches()
        # every vertex belongs to exactly one shard; every ghost row has exactly one sender
        assert sum(mm.info(12, r) for r in range(n)) == V.shape[0]
        recv, send = [mm.info(13, r) for r in range(n)], [mm.info(14, r) for r in range(n)]
        assert sum(recv) == sum(send) == mm.halo_elements() > 0
        # the same plan as the one-process-per-GPU mode makes for this patching
        for r in range(n):
            sh = D.shard_faces(F, fp, r, n)
            sm = D.ShardedMesh(sh, r, n, patch_size=ps, device=False)
            assert mm.info(10, r) == sm.count and mm.info(11, r) == sh["fv"].shape[0]
            assert mm.info(12, r) == int(sm.real_owned_mask(0).sum())
            assert mm.info(13, r) == sm.halo_slots(0).shape[0]
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mm.laplacian_smooth(V, 0.01, 1)
    with pytest.raises(RuntimeError, match="fewer patches than shards"):
        RXMeshMulti(F[:4], 8, patch_size=4096, device=False)


def test_patch_ids_are_locality_ordered():
    """Lloyd numbers patches by seed position in the face list; the builder renumbers them breadth-first over the patch
    graph (role of the reference's Patcher::bfs, patcher/patcher.cu:583-638) so that a contiguous id range is a compact slab.
    On a face-shuffled sphere the ghost rows per exchange drop several times."""
    from rxmesh_b200 import meshio
    V, F = meshio.icosphere(40)
    rng = np.random.RandomState(5)
    F = F[rng.permutation(F.shape[0])]
    code = ("import sys, numpy as np; sys.path.insert(0, %r); from rxmesh_b200.multi import RXMeshMulti;"
            "F = np.load(sys.argv[1]); m = RXMeshMulti(F, 4, patch_size=256, device=False); print(m.info(1), m.halo_elements())" % ROOT)
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        np.save(os.path.join(td, "f.npy"), F)
        out = {}
        for tag, env in (("bfs", {}), ("seed", {"RXM_NO_PATCH_REORDER": "1"})):
            e = dict(os.environ, **env)
            r = subprocess.run([sys.executable, "-c", code, os.path.join(td, "f.npy")], env=e, capture_output=True, text=True,
                               timeout=300)
            assert r.returncode == 0, r.stderr
            out[tag] = [int(t) for t in r.stdout.split()]
    assert out["bfs"][0] == out["seed"][0]          # same patches, other ids
    assert out["bfs"][1] * 2 < out["seed"][1], out  # far fewer mirrored vertices


def test_reorder_keeps_the_patches():
    """Renumbering changes ids only: the face sets of the patches are the Lloyd patcher's."""
    import rxmesh_b200 as rx
    from rxmesh_b200._lib import lib  # noqa: F401
    V, F = _mesh("ico")
    ga = rx.RXMeshStatic(F, device=False, patch_size=128)
    a = ga.elem_patch(2).copy()
    code = ("import sys, numpy as np; sys.path.insert(0, %r); import rxmesh_b200 as rx; from rxmesh_b200 import meshio;"
            "V, F = meshio.icosphere(14); g = rx.RXMeshStatic(F, device=False, patch_size=128); np.save(sys.argv[1], g.elem_patch(2))" % ROOT)
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "fp.npy")
        r = subprocess.run([sys.executable, "-c", code, p], env=dict(os.environ, RXM_NO_PATCH_REORDER="1"), capture_output=True,
                           text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        b = np.load(p)
    assert not np.array_equal(a, b)
    # a bijection between the two labelings
    pairs = np.unique(np.stack([a, b], 1), axis=0)
    assert pairs.shape[0] == a.max() + 1 == b.max() + 1
    assert np.unique(pairs[:, 0]).shape[0] == np.unique(pairs[:, 1]).shape[0] == pairs.shape[0]


# ------------------------------------------------------------------------------------------ GPU
def _device_lists():
    import torch
    n = torch.cuda.device_count()
    lists = [[0, 0], [0, 0, 0]]
    if n >= 2:
        lists += [list(range(n)), [0, 1, 0, 1][: max(3, n)]]
    return lists


@pytest.mark.gpu
@pytest.mark.parametrize("kind,ps", [("torus", 64), ("ico", 128), ("grid", 128)])
def test_multi_laplacian_and_normals(kind, ps):
    import rxmesh_b200 as rx
    from oracle import oracle as O
    from rxmesh_b200.multi import RXMeshMulti
    rx.rx_init(0)
    V, F = _mesh(kind)
    V = V.astype(np.float32)
    g = rx.RXMeshStatic(F, patch_size=ps)
    lr, iters = 0.01, 25
    one = g.laplacian_smooth_host(V, lr, iters)
    T = O.Topology(F)
    ref, vv = V.astype(np.float64), T.query("VV")
    for _ in range(iters):
        ref = O.laplacian_step(vv, ref, lr, np.float64)
    assert np.abs(one - ref).max() < 1e-5 * iters
    refn = O.vertex_normals(F, V, np.float64)
    for devs in _device_lists():
        mm = RXMeshMulti(F, devs, patch_size=ps)
        assert mm.halo_elements() > 0
        # many short steps, twice (step counter and ping-pong parity carry over), against the unsharded mesh: bit-identical
        for rep in range(2):
            got = mm.laplacian_smooth(V, lr, iters)
            assert np.array_equal(got, one), (devs, rep, np.abs(got - one).max())
        got = mm.laplacian_smooth(V, lr, 0)
        assert np.array_equal(got, V)
        got = mm.laplacian_smooth(V, lr, 3)  # odd count: the parity flips between calls
        assert np.array_equal(got, g.laplacian_smooth_host(V, lr, 3))
        n = mm.vertex_normals(V)
        rel = np.linalg.norm(n - refn, axis=1) / np.linalg.norm(refn, axis=1)
        assert rel.max() < 1e-5, devs
        del mm
    # one shard: the plain kernel
    mm = RXMeshMulti(F, [0], patch_size=ps)
    assert mm.halo_elements() == 0
    assert np.array_equal(mm.laplacian_smooth(V, lr, iters), one)


# ------------------------------------------------------------------------- C++ drop-in (include/rxmesh/rxmesh_multi.h)
def _build_multi_user(tmp_path):
    import shutil
    if shutil.which("g++") is None:
        pytest.skip("needs g++")
    exe = str(tmp_path / "multi_user")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", exe,
                           os.path.join(ROOT, "tests", "cpp", "multi_user.cpp"), "-L", os.path.join(ROOT, "rxmesh_b200"),
                           "-lrxmesh_b200", "-Wl,-rpath," + os.path.join(ROOT, "rxmesh_b200")])
    return exe


def test_cpp_multi_header_plan(tmp_path):
    """RXMeshMultiGPU compiles as plain host C++ and plans shards without a device."""
    exe = _build_multi_user(tmp_path)
    r = subprocess.run([exe, "plan", "3"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    shards, patches, verts, mirrored = [int(t) for t in r.stdout.split()]
    assert shards == 3 and verts == 90 * 61 and patches >= 3 and 0 < mirrored < verts // 4


@pytest.mark.gpu
def test_cpp_multi_header_run(tmp_path):
    import torch
    exe = _build_multi_user(tmp_path)
    lists = ["0,0", "0,0,0"] + ([",".join(str(i) for i in range(torch.cuda.device_count()))] if torch.cuda.device_count() > 1 else [])
    for devs in lists:
        r = subprocess.run([exe, "run", devs], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and r.stdout.startswith("ok"), (devs, r.stdout, r.stderr)


@pytest.mark.gpu
def test_same_device_shards_with_a_long_cut_are_refused():
    """Shards on ONE device exist for tests: a fused step's boundary blocks spin on flags the neighbour's kernel raises, so a
    plan whose waiting blocks could fill the device is refused instead of risking a hang; on separate devices it is fine."""
    import torch
    import rxmesh_b200 as rx
    from rxmesh_b200 import meshio
    from rxmesh_b200.multi import RXMeshMulti
    rx.rx_init(0)
    nx, ny = 6001, 65
    V, F = meshio.grid(nx, ny)
    fp = meshio.grid_face_tiles(nx, ny, 8, 8)  # row-major 8 x 8-quad tiles: the cut between two shards runs along 750 tiles
    with pytest.raises(RuntimeError, match="share a device"):
        RXMeshMulti(F, [0, 0], face_patch=fp, patch_size=128)
    if torch.cuda.device_count() >= 2:
        mm = RXMeshMulti(F, [0, 1], face_patch=fp, patch_size=128)
        one = RXMeshMulti(F, [0], face_patch=fp, patch_size=128)
        Vf = V.astype(np.float32)
        assert np.array_equal(mm.laplacian_smooth(Vf, 0.01, 10), one.laplacian_smooth(Vf, 0.01, 10))
