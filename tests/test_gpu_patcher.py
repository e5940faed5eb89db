"""The Lloyd patcher's passes on the GPU (rxmesh_b200/csrc/rxm_patcher_gpu.cu) give THE SAME face -> patch array as the host
passes (mesh_builder.cpp: patcher_lloyd, itself tested against the serial FIFO definition in tests/test_host_build.py)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import make_mesh  # noqa: E402


def _patching(F, ps, gpu):
    import rxmesh_b200 as rx
    os.environ["RXM_PATCHER_GPU"] = "1" if gpu else "0"
    os.environ["RXM_VERBOSE"] = "1"
    try:
        m = rx.RXMeshStatic(F, device=False, patch_size=ps)
        return m.elem_patch(2).copy(), m.get_num_patches(), m.build_seconds(True)
    finally:
        del os.environ["RXM_PATCHER_GPU"], os.environ["RXM_VERBOSE"]


@pytest.mark.gpu
@pytest.mark.parametrize("name,ps", [("ico40", 256), ("torus120x90", 512), ("grid301x207", 1024), ("damaged1", 64), ("damaged5", 128),
                                     ("sphere3", 64), ("dragon", 512), ("ico120", 512)])
def test_gpu_patcher_equals_host(name, ps, capfd):
    import rxmesh_b200 as rx
    rx.rx_init(0)
    V, F = make_mesh(name)
    if name == "ico40":  # face order without locality: seeds land anywhere
        F = F[np.random.RandomState(7).permutation(F.shape[0])]
    a, na, _ = _patching(F, ps, False)
    capfd.readouterr()
    b, nb, _ = _patching(F, ps, True)
    err = capfd.readouterr().err
    assert "gpu)" in err, err  # the verbose line names the path that ran: the comparison is not host against host
    assert na == nb and np.array_equal(a, b)
    assert np.bincount(b).max() <= ps


@pytest.mark.gpu
def test_gpu_patcher_many_components_falls_back():
    """more pieces than the GPU loop seeds one by one (256): the host passes run, the result is the host's"""
    import rxmesh_b200 as rx
    rx.rx_init(0)
    V, F = make_mesh("ico4")
    F = np.concatenate([F + k * V.shape[0] for k in range(300)]).astype(np.uint32)
    a, na, _ = _patching(F, 4096, False)
    b, nb, _ = _patching(F, 4096, True)
    assert na == nb and np.array_equal(a, b)
