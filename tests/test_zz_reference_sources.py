"""The reference's OWN user-kernel source files, compiled UNMODIFIED against the drop-in headers (include/rxmesh,
include/glm) and run on the GPU: oracle/_ref/libshim_refsrc1.so holds apps/VertexNormal/vertex_normal_kernel.cuh,
apps/GaussianCurvature/gaussian_curvature_kernel.cuh, apps/MCF/mcf_kernels.cuh (matrix-free mat-vec),
tests/RXMesh_test/query_kernel.cuh and tests/RXMesh_test/higher_query.cuh; libshim_refsrc2.so holds
apps/Filtering/filtering_rxmesh_kernel.cuh (+ filtering_util.h).  They are built by `make -C oracle ref_user_kernels`
where /root/reference exists (sources are included from there, never copied) behind the drivers of tests/cpp/shim_apps.cu,
so the checks are those of tests/test_gpu_shim.py -- same inputs, same oracle comparisons -- with the reference's kernels
in place of the restated ones.  (The file name sorts last on purpose: these are the newest checks.)"""
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


@pytest.mark.parametrize("name", ["sphere3", "dragon"])
def test_reference_higher_query_kernel(refsrc1, name):
    """tests/RXMesh_test/higher_query.cuh: higher_query<512, Op::VV> (2-ring through higher_query_block_dispatcher)"""
    S.test_user_higher_query_two_ring(refsrc1, name)


@pytest.mark.parametrize("name", ["sphere3", "dragon"])
def test_reference_filtering_kernels(refsrc2, name):
    """apps/Filtering/filtering_rxmesh_kernel.cuh: compute_vertex_normal<float, 512> + bilateral_filtering<float, 512, 80>
    (query_block_dispatcher / higher_query_block_dispatcher inside a per-vertex k-ring search)"""
    S.test_user_filtering_app(refsrc2, name)


def test_reference_gtest_files(tmp_path):
    """The reference's own gtest FILES -- tests/RXMesh_test/test_attribute.cu (11 tests: ReduceHandle norm2 / dot / reduce /
    arg_max, copy_from, add / remove, layouts, tensor SoA on host and device, reset, to_matrix / from_matrix),
    test_boundary.cu (bunnyhead: 98 boundary vertices through a VertexAttribute<bool>), test_ev_diamond.cu (plane_5: the two
    triangles of every interior diamond have area 1), test_export.cu (export_obj / export_vtk) and test_for_each.cu (host
    and device for_each) -- compiled unmodified into oracle/_ref/ref_gtests and run in a directory holding the meshes they
    name, written from the committed fixtures (input/bumpy-cube.obj is not among them: dragon stands in, the ArgMax test
    does not depend on the geometry).  Runs in its own process: some of these tests call cudaDeviceReset()."""
    import subprocess

    import numpy as np

    from conftest import make_mesh
    exe = os.path.join(REFDIR, "ref_gtests")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_gtests not built (needs /root/reference at build time: make -C oracle ref_user_kernels)")
    inp = tmp_path / "rxm_input"
    inp.mkdir()
    for obj, mesh in (("sphere3", "sphere3"), ("cube", "cube"), ("bunnyhead", "bunnyhead"), ("plane_5", "plane_5"),
                      ("bumpy-cube", "dragon")):
        V, F = make_mesh(mesh)
        S._write_obj(str(inp / (obj + ".obj")), np.asarray(V), np.asarray(F))
    r = subprocess.run([exe], cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert "[==========] 16 tests ran, 0 failed" in r.stdout, tail
    assert os.path.exists(tmp_path / "sphere3.vtk") and os.path.exists(tmp_path / "sphere3.obj")
