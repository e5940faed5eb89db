"""Mean-curvature flow by matrix-free CG (apps/MCF/mcf_cg_mat_free.h, matrix/cg_mat_free_attr_solver.h:45-125):
the oracle's solver on the CPU, and rxm_mcf_solve against it on the GPU.

The reference app has no correctness check for the solve (apps/MCF/mcf.cu), so the solve is pinned by
  * golden vectors from the reference's own kernels (init_B / matvec / precond_matvec, unmodified) run on a B200 under the
    drop-in solvers (tests/golden/ref_mcf.npz, test_oracle_vs_reference_kernels_golden), and by what it must satisfy:
  * the mat-vec it iterates is the one the reference's unmodified mcf_kernels.cuh computes (tests/test_gpu_shim.py,
    tests/test_oracle.py::test_mcf_matvec_constant_vector),
  * the result solves (M + dt L) X = M X0: the residual B - A X, evaluated by the float64 oracle, is at the tolerance,
  * known answers: one positive mass per vertex shared by x, y, z; a sphere stays a sphere and shrinks.
"""
import numpy as np
import pytest

from conftest import make_mesh
from oracle import oracle as O

CLOSED = ["sphere3", "ico10", "torus40x30", "dragon"]


def _mesh(name):
    V, F = make_mesh(name)
    return np.ascontiguousarray(V, np.float32), F


@pytest.mark.parametrize("name", CLOSED)
@pytest.mark.parametrize("uniform", [True, False])
def test_oracle_solve_satisfies_the_system(name, uniform):
    V, F = _mesh(name)
    rings = O.oriented_rings(F, V.shape[0])
    dt = 10.0 if uniform else 1e-2  # the cotangent system with the app's dt = 10 is too ill-conditioned for 100 iterations
    X, info, res = O.mcf_solve(rings, V, dt, uniform, 200, 1e-6, 0.0, with_residual=True)
    assert info["converged"] and 0 < info["iterations"] < 200
    assert info["final_residual"] < 1e-6 <= info["start_residual"]
    # the recursively updated residual is the true one (float64: no drift to speak of)
    assert abs((res ** 2).sum() - info["final_residual"]) < 1e-9
    # the independent residual routine agrees
    res2, bb = O.mcf_residual(rings, V, X, dt, uniform)
    assert np.allclose(res, res2, rtol=0, atol=1e-12) and bb > 0


def test_oracle_matvec_is_the_pinned_one():
    """The solver's operator (weights precomputed) equals rxo_mcf_matvec, the restatement the reference's own kernel is
    checked against: residual(X) = B - A X for X = X0 gives B - matvec(X0)."""
    V, F = _mesh("ico10")
    rings = O.oriented_rings(F, V.shape[0])
    res, _ = O.mcf_residual(rings, V, V.astype(np.float64), 0.5, uniform=False)
    AX = O.mcf_matvec(rings, V, V, 0.5)
    # B = X0 / v_weight = the mat-vec's own diagonal mass: take it from a solve-free identity, A X0 + res = B
    B = AX + res
    V64 = V.astype(np.float64)
    c = np.abs(V64).argmax(axis=1)
    k = B[np.arange(V.shape[0]), c] / V64[np.arange(V.shape[0]), c]
    assert np.all(k > 0) and np.allclose(B, k[:, None] * V64, rtol=0, atol=1e-12)  # one positive mass per vertex for x, y, z


def test_oracle_sphere_shrinks_uniformly():
    """Known answer: on a (nearly) uniform icosphere the uniform-Laplacian flow moves every vertex towards the centre."""
    V, F = _mesh("ico10")
    rings = O.oriented_rings(F, V.shape[0])
    X, info = O.mcf_solve(rings, V, 10.0, True, 100, 1e-10, 0.0)
    r0, r1 = np.linalg.norm(V, axis=1), np.linalg.norm(X, axis=1)
    assert info["converged"] and np.all(r1 < r0) and r1.std() < 0.05 * r1.mean()
    cosang = (X * V).sum(1) / (r0 * r1)
    assert cosang.min() > 0.999


@pytest.mark.parametrize("name", CLOSED)
@pytest.mark.parametrize("uniform", [True, False])
def test_oracle_jacobi_pcg_solves_the_same_system(name, uniform):
    """The preconditioned form (pcg_mat_free_attr_solver.h:40-140 with precond_matvec) reaches the same solution; its
    residual is measured in the M^-1 norm, so compare at tight relative tolerances."""
    V, F = _mesh(name)
    rings = O.oriented_rings(F, V.shape[0])
    dt = 10.0 if uniform else 1e-2
    a, ia = O.mcf_solve(rings, V, dt, uniform, 1000, 0.0, 1e-14)
    b, ib, res = O.mcf_solve(rings, V, dt, uniform, 1000, 0.0, 1e-14, with_residual=True, precond=True)
    assert ia["converged"] and ib["converged"] and ib["iterations"] <= ia["iterations"]
    assert np.abs(a - b).max() < 1e-7 * np.abs(V).max()
    assert (res ** 2).sum() < 1e-10 * ia["start_residual"]


def test_oracle_vs_reference_kernels_golden():
    """The pin of the oracle's solve: tests/golden/ref_mcf.npz holds what THE REFERENCE'S OWN kernels (init_B, matvec,
    precond_matvec of apps/MCF/mcf_kernels.cuh, compiled unmodified) produced on a B200 under the drop-in CG / PCG solvers
    (tests/golden/make_golden_mcf.py): 3 meshes x 2 Laplacians x 2 solvers.  The float64 oracle reaches the same solution in
    the same number of iterations (fp32 residual recursion: +-2, +-10 % beyond 20) from the same start residual."""
    import os
    from conftest import GOLDEN
    sys_path = os.path.join(GOLDEN, "make_golden_mcf.py")
    assert os.path.exists(sys_path)
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_mcf", sys_path)
    G = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(G)
    gold = np.load(os.path.join(GOLDEN, "ref_mcf.npz"))
    for name, uni, pcg in G.CASES:
        V, F = _mesh(name)
        rings = O.oriented_rings(F, V.shape[0])
        dt, ta, tr, mi = G.params(uni, pcg)
        ref, oinfo = O.mcf_solve(rings, V, dt, bool(uni), mi, ta, tr, precond=bool(pcg))
        key = "%s_%s_%s" % (name, "uniform" if uni else "cotangent", "pcg" if pcg else "cg")
        X, info = gold[key + "_X"], gold[key + "_info"]
        assert oinfo["converged"], key
        assert abs(int(info[0]) - oinfo["iterations"]) <= 2 + oinfo["iterations"] // 10, (key, info, oinfo)
        assert abs(info[1] - oinfo["start_residual"]) < 2e-3 * oinfo["start_residual"], (key, info, oinfo)
        assert np.abs(X - ref).max() < (1e-5 if uni else 5e-5) * np.abs(V).max(), (key, np.abs(X - ref).max())


def test_oracle_zero_iterations_and_max_iter():
    V, F = _mesh("sphere3")
    rings = O.oriented_rings(F, V.shape[0])
    X, info = O.mcf_solve(rings, V, 10.0, True, 0, 1e-6, 0.0)
    assert info["iterations"] == 0 and not info["converged"] and np.array_equal(X, V.astype(np.float64))
    X3, info3 = O.mcf_solve(rings, V, 10.0, True, 3, 1e-30, 0.0)
    assert info3["iterations"] == 3 and not info3["converged"]
    assert info3["final_residual"] < info3["start_residual"]


# ------------------------------------------------------------------------------------------------- GPU
def _gpu_solve(m, rx, V, **kw):
    x = m.add_vertex_attribute("mcf_x0", np.float32, 3, rx.LOCATION_ALL, kw.pop("layout", rx.AoS))
    y = m.add_vertex_attribute("mcf_x", np.float32, 3, rx.LOCATION_ALL, kw.pop("layout_out", rx.AoS))
    x.from_global(V)
    info = m.mcf_solve(x, y, **kw)
    got = y.to_global()
    m.remove_attribute("mcf_x0"), m.remove_attribute("mcf_x")
    return got, info


def _res_floor(V, dt, uniform, rings):
    """what fp32 rounding of A X alone leaves in |B - A X|^2: (eps * magnitude of the terms)^2 per component, summed"""
    if uniform:
        val = np.diff(rings[0].astype(np.int64)).astype(np.float64)
        mag = val * (1.0 + 2.0 * dt) * np.abs(V).max()
    else:
        _, mag = O.mcf_matvec(rings, V, V, dt, with_scale=True)
    return 3.0 * ((1.2e-7 * mag) ** 2).sum(), mag


@pytest.mark.gpu
@pytest.mark.parametrize("name", CLOSED + ["ico40"])
@pytest.mark.parametrize("uniform", [True, False])
def test_gpu_solve_vs_oracle(name, uniform):
    """Tolerances: fp32 against the float64 oracle.  Calibrated with a numpy fp32 emulation of the same algorithm (float64
    dot products, weights rounded to fp32 and perturbed by 3e-7 relative): three-iteration states agree to 4e-7 of the mesh
    size with the uniform Laplacian, converged solutions to 7e-6; the bounds below leave a factor 5-10."""
    import rxmesh_b200 as rx
    rx.rx_init(0)
    V, F = _mesh(name)
    m = rx.RXMeshStatic(F, patch_size=512 if F.shape[0] > 600 else 64)
    rings = O.oriented_rings(F, V.shape[0])
    dt = 10.0 if uniform else 1e-2  # cotangent: the app's dt = 10 on unit-size meshes does not converge in 100 iterations
    scale = np.abs(V).max()
    floor2, mag = _res_floor(V, dt, uniform, rings)
    # (1) the trajectory: three iterations, no convergence test in the way
    got3, info3 = _gpu_solve(m, rx, V, time_step=dt, use_uniform_laplace=uniform, max_iter=3, tol_abs=0.0, tol_rel=0.0)
    ref3, oinfo3 = O.mcf_solve(rings, V, dt, uniform, 3, 0.0, 0.0)
    assert info3["iterations"] == 3 and not info3["converged"]
    # R0 = B - A X0 = -dt L X0 is a difference of terms `mag` that cancel to |R0|: fp32 evaluates it (here and in the
    # reference's own mat-vec) with relative error rho = eps mag / |R0|, and the first iterations move X along R0
    rho = 3e-7 * mag.max() / np.sqrt(oinfo3["start_residual"] / (3 * V.shape[0]))
    tol3 = 1e-5 * scale + (0.0 if uniform else 10.0 * rho * np.abs(ref3 - V).max())
    assert np.abs(got3 - ref3).max() < tol3, (name, uniform, np.abs(got3 - ref3).max(), tol3)
    assert abs(info3["start_residual"] - oinfo3["start_residual"]) < (1e-4 + 20 * rho) * oinfo3["start_residual"]
    # (2) the converged solve.  uniform: the app's tolerances (mcf.cu:19-23); cotangent: a relative tolerance tight enough
    #     that the solution, not the stopping point, is compared
    ta, tr, mi = (1e-6, 0.0, 200) if uniform else (0.0, 1e-9, 500)
    got, info = _gpu_solve(m, rx, V, time_step=dt, use_uniform_laplace=uniform, max_iter=mi, tol_abs=ta, tol_rel=tr)
    ref, oinfo = O.mcf_solve(rings, V, dt, uniform, mi, ta, tr)
    assert info["converged"] and oinfo["converged"]
    assert abs(info["iterations"] - oinfo["iterations"]) <= 2 + oinfo["iterations"] // 10, (info, oinfo)
    res, bb = O.mcf_residual(rings, V, got, dt, uniform)
    true_r2 = (res ** 2).sum()
    assert true_r2 < 10.0 * (max(ta, tr * oinfo["start_residual"]) + floor2), (name, uniform, true_r2, floor2)
    assert np.abs(got - ref).max() < (1e-5 if uniform else 5e-5) * scale, (name, uniform, np.abs(got - ref).max())
    if not uniform:  # the app's absolute tolerance with the cotangent Laplacian: stops where the oracle stops
        got, info = _gpu_solve(m, rx, V, time_step=dt, use_uniform_laplace=False, max_iter=200, tol_abs=1e-6, tol_rel=0.0)
        ref, oinfo = O.mcf_solve(rings, V, dt, False, 200, 1e-6, 0.0)
        assert info["converged"] and abs(info["iterations"] - oinfo["iterations"]) <= 2 + oinfo["iterations"] // 10
        res, bb = O.mcf_residual(rings, V, got, dt, False)
        assert (res ** 2).sum() < 10.0 * (1e-6 + floor2)



@pytest.mark.gpu
@pytest.mark.parametrize("name", CLOSED + ["ico40"])
@pytest.mark.parametrize("uniform", [True, False])
def test_gpu_jacobi_pcg_vs_oracle(name, uniform):
    """rxm_mcf_solve_ex(jacobi = 1) against the oracle's preconditioned solve: the three-iteration state, then the converged
    solve (same iteration count, same solution, the system solved).  Tolerances as in test_gpu_solve_vs_oracle."""
    import rxmesh_b200 as rx
    rx.rx_init(0)
    V, F = _mesh(name)
    m = rx.RXMeshStatic(F, patch_size=512 if F.shape[0] > 600 else 64)
    rings = O.oriented_rings(F, V.shape[0])
    dt = 10.0 if uniform else 1e-2
    scale = np.abs(V).max()
    floor2, mag = _res_floor(V, dt, uniform, rings)
    got3, info3 = _gpu_solve(m, rx, V, time_step=dt, use_uniform_laplace=uniform, max_iter=3, tol_abs=0.0, tol_rel=0.0,
                             precondition=True)
    ref3, oinfo3 = O.mcf_solve(rings, V, dt, uniform, 3, 0.0, 0.0, precond=True)
    plain3, pinfo3 = O.mcf_solve(rings, V, dt, uniform, 3, 0.0, 0.0)
    assert info3["iterations"] == 3 and not info3["converged"]
    rho = 3e-7 * mag.max() / np.sqrt(pinfo3["start_residual"] / (3 * V.shape[0]))
    tol3 = 1e-5 * scale + (0.0 if uniform else 10.0 * rho * np.abs(ref3 - V).max())
    assert np.abs(got3 - ref3).max() < tol3, (name, uniform, np.abs(got3 - ref3).max(), tol3)
    assert abs(info3["start_residual"] - oinfo3["start_residual"]) < (1e-4 + 20 * rho) * oinfo3["start_residual"]
    ta, tr, mi = (0.0, 1e-9, 500)
    got, info = _gpu_solve(m, rx, V, time_step=dt, use_uniform_laplace=uniform, max_iter=mi, tol_abs=ta, tol_rel=tr,
                           precondition=True)
    ref, oinfo = O.mcf_solve(rings, V, dt, uniform, mi, ta, tr, precond=True)
    assert info["converged"] and oinfo["converged"]
    assert abs(info["iterations"] - oinfo["iterations"]) <= 2 + oinfo["iterations"] // 10, (info, oinfo)
    res, bb = O.mcf_residual(rings, V, got, dt, uniform)
    assert (res ** 2).sum() < 10.0 * (1e-7 * pinfo3["start_residual"] + floor2), (name, uniform, (res ** 2).sum(), floor2)
    assert np.abs(got - ref).max() < (1e-5 if uniform else 5e-5) * scale, (name, uniform, np.abs(got - ref).max())
    # fewer iterations than the plain solver needs for the same reduction of ITS residual is not guaranteed in general, but
    # the two must meet at the solution
    plain, _ = O.mcf_solve(rings, V, dt, uniform, 1000, 0.0, 1e-12)
    assert np.abs(got - plain).max() < (1e-5 if uniform else 5e-5) * scale


@pytest.mark.gpu
def test_gpu_solve_is_deterministic_and_layout_independent():
    import rxmesh_b200 as rx
    rx.rx_init(0)
    V, F = _mesh("ico40")
    m = rx.RXMeshStatic(F, patch_size=256)
    a, ia = _gpu_solve(m, rx, V, max_iter=100)
    b, ib = _gpu_solve(m, rx, V, max_iter=100)
    assert ia == ib and np.array_equal(a.view(np.uint32), b.view(np.uint32))  # block-ordered reductions: same bits
    c, ic = _gpu_solve(m, rx, V, max_iter=100, layout=rx.AoSoA, layout_out=rx.SoA)  # the reference's default layouts
    assert ic == ia and np.array_equal(a.view(np.uint32), c.view(np.uint32))
    # a patching with other patch sizes changes the summation order of the reductions only
    m2 = rx.RXMeshStatic(F, patch_size=1024)
    d, idd = _gpu_solve(m2, rx, V, max_iter=100)
    assert abs(idd["iterations"] - ia["iterations"]) <= 1 and np.abs(a - d).max() < 1e-5


@pytest.mark.gpu
def test_gpu_solve_stops_and_refuses_like_the_reference():
    import rxmesh_b200 as rx
    rx.rx_init(0)
    V, F = _mesh("sphere3")
    m = rx.RXMeshStatic(F, patch_size=64)
    got, info = _gpu_solve(m, rx, V, max_iter=0)
    assert info["iterations"] == 0 and not info["converged"] and np.array_equal(got, V)
    got, info = _gpu_solve(m, rx, V, max_iter=100, tol_abs=0.0, tol_rel=1e-4)  # the relative test alone
    assert info["converged"] and info["final_residual"] / info["start_residual"] < 1e-4
    # an open mesh: "mcf_rxmesh only takes watertight/closed mesh without boundaries" (mcf_cg_mat_free.h:39-43)
    Vg, Fg = make_mesh("grid8x8")
    mg = rx.RXMeshStatic(Fg, patch_size=64)
    with pytest.raises(rx.RXMeshError, match="closed"):
        _gpu_solve(mg, rx, np.ascontiguousarray(Vg, np.float32))
