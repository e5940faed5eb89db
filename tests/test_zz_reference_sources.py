"""The reference's OWN user-kernel source files, compiled UNMODIFIED against the drop-in headers (include/rxmesh,
include/glm) and run on the GPU: oracle/_ref/libshim_refsrc1.so holds apps/VertexNormal/vertex_normal_kernel.cuh,
apps/GaussianCurvature/gaussian_curvature_kernel.cuh, apps/MCF/mcf_kernels.cuh (matrix-free mat-vec),
tests/RXMesh_test/query_kernel.cuh and tests/RXMesh_test/higher_query.cuh; libshim_refsrc2.so holds
apps/Filtering/filtering_rxmesh_kernel.cuh (+ filtering_util.h).  They are built by `make -C oracle ref_user_kernels`
where /root/reference exists (sources are included from there, never copied) behind the drivers of tests/cpp/shim_apps.cu,
so the checks are those of tests/test_gpu_shim.py -- same inputs, same oracle comparisons -- with the reference's kernels
in place of the restated ones.  (The file name sorts last on purpose: these are the newest checks.)  The end of the file
also holds the other GPU checks written after the last GPU session of round 1: queries, vertex normals and Laplacian
smoothing on damaged meshes (conftest.damaged_mesh)."""
import ctypes as C
import os

import pytest

import rxmesh_b200 as rx
import test_gpu_shim as S
from conftest import ROOT

pytestmark = pytest.mark.gpu
REFDIR = os.path.join(ROOT, "oracle", "_ref")


def _load(name):
    path = os.path.join(REFDIR, name)
    if not os.path.exists(path):
        pytest.skip("%s not built (needs /root/reference at build time: make -C oracle ref_user_kernels)" % name)
    rx.lib()
    return C.CDLL(path)


@pytest.fixture(scope="module")
def refsrc1():
    return _load("libshim_refsrc1.so")


@pytest.fixture(scope="module")
def refsrc2():
    return _load("libshim_refsrc2.so")


@pytest.mark.parametrize("name", ["sphere3", "dragon", "bunnyhead"])
def test_reference_vertex_normal_kernel(refsrc1, name):
    """apps/VertexNormal/vertex_normal_kernel.cuh: compute_vertex_normal<float, 256>"""
    S.test_user_vertex_normal_kernel(refsrc1, name)


@pytest.mark.parametrize("op", ["VV", "VE", "VF", "EV", "EF", "FV", "FE", "FF"])
def test_reference_query_kernel(refsrc1, op):
    """tests/RXMesh_test/query_kernel.cuh: query_kernel<256, op, ...> for the eight static queries"""
    S.test_user_query_kernel(refsrc1, op)


@pytest.mark.parametrize("op", ["EVDiamond", "EE"])
def test_reference_query_kernel_edge4(refsrc1, op):
    S.test_user_edge4_query_kernel(refsrc1, op)


def test_reference_query_kernel_oriented_vv(refsrc1):
    S.test_oriented_vv(refsrc1)


@pytest.mark.parametrize("name", ["sphere3", "dragon"])
def test_reference_gaussian_curvature_kernel(refsrc1, name):
    """apps/GaussianCurvature/gaussian_curvature_kernel.cuh: compute_gaussian_curvature<float, 256>"""
    S.test_user_gaussian_curvature(refsrc1, name)


@pytest.mark.parametrize("name", ["sphere3", "torus", "dragon"])
def test_reference_mcf_matvec_kernel(refsrc1, name):
    """apps/MCF/mcf_kernels.cuh: matvec<float, 256> (cotan weights over ORIENTED VV, use_uniform_laplace = false)"""
    S.test_user_mcf_matvec(refsrc1, name)


@pytest.mark.parametrize("pcg", [0, 1])
@pytest.mark.parametrize("name,uniform", [("sphere3", 0), ("torus40x30", 0), ("dragon", 0), ("sphere3", 1), ("dragon", 1)])
def test_reference_mcf_cg(refsrc1, name, uniform, pcg):
    """apps/MCF/mcf_kernels.cuh: init_B<float, 256>, matvec<float, 256> and precond_matvec<float, 256> (both Laplacians),
    unmodified, under the drop-in CGMatFreeAttrSolver / PCGMatFreeAttrSolver -- the MCF app's two matrix-free solves --
    against the oracle and the fixed-function rxm_mcf_solve / rxm_mcf_solve_ex"""
    S.test_user_mcf_cg(refsrc1, name, uniform, pcg, kernels_have_uniform=True)


@pytest.mark.parametrize("name", ["sphere3", "dragon"])
def test_reference_higher_query_kernel(refsrc1, name):
    """tests/RXMesh_test/higher_query.cuh: higher_query<512, Op::VV> (2-ring through higher_query_block_dispatcher)"""
    S.test_user_higher_query_two_ring(refsrc1, name)


@pytest.mark.parametrize("name", ["sphere3", "dragon"])
def test_reference_filtering_kernels(refsrc2, name):
    """apps/Filtering/filtering_rxmesh_kernel.cuh: compute_vertex_normal<float, 512> + bilateral_filtering<float, 512, 80>
    (query_block_dispatcher / higher_query_block_dispatcher inside a per-vertex k-ring search)"""
    S.test_user_filtering_app(refsrc2, name)


_GTESTS = ["Attribute.Norm2", "Attribute.Dot", "Attribute.Reduce", "Attribute.ArgMax", "Attribute.CopyFrom",
           "Attribute.AddingAndRemoving", "Attribute.DefaultLayoutIsAoSoA", "Attribute.TrueSoAHostStorageIsColumnMajor",
           "Attribute.TrueSoADeviceWritesColumnMajor", "Attribute.ResetSetsAllComponents",
           "Attribute.ToAndFromMatrixPreserveLayouts", "RXMeshStatic.BoundaryVertex", "RXMeshStatic.EVDiamond",
           "RXMeshStatic.Export", "RXMeshStatic.ForEach", "RXMeshStatic.ForEachOnDevice"]


@pytest.fixture(scope="module")
def gtest_run(tmp_path_factory):
    """The reference's own gtest FILES -- tests/RXMesh_test/test_attribute.cu, test_boundary.cu, test_ev_diamond.cu,
    test_export.cu, test_for_each.cu -- compiled unmodified into oracle/_ref/ref_gtests, run ONCE in a directory holding the
    meshes they name, written from the committed fixtures (input/bumpy-cube.obj is not among them: dragon stands in, the
    ArgMax test does not depend on the geometry).  Own process: some of these tests call cudaDeviceReset()."""
    import subprocess

    import numpy as np

    from conftest import make_mesh
    exe = os.path.join(REFDIR, "ref_gtests")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_gtests not built (needs /root/reference at build time: make -C oracle ref_user_kernels)")
    work = tmp_path_factory.mktemp("ref_gtests")
    inp = work / "rxm_input"
    inp.mkdir()
    for obj, mesh in (("sphere3", "sphere3"), ("cube", "cube"), ("bunnyhead", "bunnyhead"), ("plane_5", "plane_5"),
                      ("bumpy-cube", "dragon")):
        V, F = make_mesh(mesh)
        S._write_obj(str(inp / (obj + ".obj")), np.asarray(V), np.asarray(F))
    r = subprocess.run([exe], cwd=str(work), capture_output=True, text=True, timeout=600)
    status = {}
    for line in r.stdout.split("\n"):
        if line.startswith("[       OK ] "):
            status[line[13:].strip()] = True
        elif line.startswith("[  FAILED  ] "):
            status[line[13:].strip()] = False
    return dict(rc=r.returncode, status=status, tail=(r.stdout + r.stderr)[-3000:], work=work)


@pytest.mark.parametrize("name", _GTESTS)
def test_reference_gtest(gtest_run, name):
    """one of the reference's own TEST()s (see gtest_run): ReduceHandle norm2 / dot / reduce / arg_max, copy_from, add /
    remove, layouts, tensor SoA on host and device, reset, to_matrix / from_matrix; bunnyhead's 98 boundary vertices through
    a VertexAttribute<bool>; plane_5's unit diamonds; export_obj / export_vtk; host and device for_each"""
    assert gtest_run["status"].get(name) is True, gtest_run["tail"]


def test_reference_gtest_binary_summary(gtest_run):
    assert gtest_run["rc"] == 0 and len(gtest_run["status"]) == len(_GTESTS), gtest_run["tail"]
    assert os.path.exists(gtest_run["work"] / "sphere3.vtk") and os.path.exists(gtest_run["work"] / "sphere3.obj")


# ---- not reference sources, but also added after the last GPU session: the eight queries on damaged inputs ----
@pytest.fixture(scope="module", params=["damaged0", "damaged1", "damaged3", "damaged8"])
def damaged(request):
    """conftest.damaged_mesh: holes, several components, flipped faces (no fans), a non-manifold edge with three faces
    (odd seeds: the generic FF / EF paths, no stored rows), or holes with the orientation kept (8: open fans)"""
    from conftest import make_mesh
    from oracle import oracle as O
    V, F = make_mesh(request.param)
    return request.param, V, F, rx.RXMeshStatic(F, patch_size=64), O.Topology(F)


@pytest.mark.parametrize("op", ["VV", "VE", "VF", "EV", "EF", "FV", "FE", "FF"])
def test_queries_on_damaged_meshes(damaged, op):
    import test_gpu_queries as Q
    Q.test_query(damaged, op)


@pytest.fixture(scope="module", params=["damaged5", "damaged8"])
def damaged_apps(request):
    """orientation kept, a few holes: 5 adds a non-manifold edge (no fans: the transposing fallbacks of the fixed-function
    kernels), 8 keeps the fans with open ones around the holes"""
    from conftest import make_mesh
    from oracle import oracle as O
    rx.rx_init(0)
    V, F = make_mesh(request.param)
    return request.param, V, F, rx.RXMeshStatic(F, patch_size=64), O.Topology(F)


def test_vertex_normals_on_damaged_meshes(damaged_apps):
    import test_gpu_apps as A
    A.test_vertex_normals(damaged_apps)


def test_laplacian_on_damaged_meshes(damaged_apps):
    import test_gpu_apps as A
    A.test_laplacian(damaged_apps)
