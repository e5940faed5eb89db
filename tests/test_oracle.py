"""The oracle against the reference's golden vectors / known answers (CPU only)."""
import json
import os

import numpy as np

from conftest import GOLDEN, load_golden
from oracle import oracle as O

KNOWN = json.load(open(os.path.join(GOLDEN, "known_answers.json")))


def test_normals_bit_exact_vs_reference_loop():
    # vn_ref was produced by the reference's own vertex_normal_ref.h compiled unmodified
    # (oracle/_ref, tests/golden/make_golden.py)
    for name in ("sphere3", "dragon", "bunnyhead", "torus"):
        g = load_golden(name)
        n = O.vertex_normals(g["F"], g["V"], np.float32)
        assert np.array_equal(n.view(np.uint32), g["vn_ref"].view(np.uint32)), name


def test_normals_survey_checksums():
    g = load_golden("dragon")
    n = O.vertex_normals(g["F"], g["V"], np.float32)
    assert abs(n.sum(dtype=np.float64) - KNOWN["dragon_sum"]) < 2e-3
    assert abs(np.abs(n).sum(dtype=np.float64) - KNOWN["dragon_abs_sum"]) < 2e-2
    assert np.allclose(n[0], KNOWN["dragon_n0"], rtol=0, atol=1e-6)
    assert np.allclose(n[-1], KNOWN["dragon_nlast"], rtol=0, atol=1e-6)
    g = load_golden("sphere3")
    n = O.vertex_normals(g["F"], g["V"], np.float32)
    assert abs(np.abs(n).sum(dtype=np.float64) - KNOWN["sphere3_abs_sum"]) < 1e-3
    assert np.allclose(n[0], KNOWN["sphere3_n0"], rtol=0, atol=1e-6)
    assert np.allclose(n[-1], KNOWN["sphere3_nlast"], rtol=0, atol=1e-6)
    assert list(g["V"].shape[:1]) + list(g["F"].shape[:1]) == KNOWN["sphere3_VF"]


def test_normals_f32_close_to_f64():
    g = load_golden("dragon")
    a = O.vertex_normals(g["F"], g["V"], np.float32).astype(np.float64)
    b = O.vertex_normals(g["F"], g["V"], np.float64)
    rel = np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)
    assert rel.max() < 1e-5


def test_cube_counts_and_euler():
    g = load_golden("cube")
    T = O.Topology(g["F"])
    assert dict(V=T.nv, E=T.ne, F=T.nf) == KNOWN["cube_counts"]
    for name in ("sphere3", "dragon", "torus"):
        T = O.Topology(load_golden(name)["F"])
        chi = T.nv - T.ne + T.nf
        assert chi == {"sphere3": 2, "dragon": 0, "torus": 0}[name]  # the decimated dragon has genus 1
        assert T.stats()["is_closed"] and T.stats()["is_edge_manifold"]


def test_bunnyhead_boundary():
    T = O.Topology(load_golden("bunnyhead")["F"])
    n, flags = T.boundary_vertices()
    assert n == KNOWN["bunnyhead_boundary_vertices"] == int(flags.sum())
    assert not T.stats()["is_closed"]


def test_edge_numbering_first_seen():
    # rxmesh.cpp:589-611: ids in order of first appearance, (max, min) key
    fv = np.array([[0, 1, 2], [2, 1, 3]], dtype=np.uint32)
    T = O.Topology(fv)
    assert T.ev.tolist() == [[1, 0], [2, 1], [2, 0], [3, 1], [3, 2]]
    assert T.fe.tolist() == [[0, 1, 2], [1, 3, 4]]


def test_queries_are_mutually_consistent():
    g = load_golden("sphere3")
    T = O.Topology(g["F"])
    vv, ve, vf = T.query("VV"), T.query("VE"), T.query("VF")
    ev, ef, fv, fe, ff = T.query("EV"), T.query("EF"), T.query("FV"), T.query("FE"), T.query("FF")
    assert vv[1].shape[0] == ve[1].shape[0] == 2 * T.ne
    assert vf[1].shape[0] == ef[1].shape[0] == fv[1].shape[0] == fe[1].shape[0] == 3 * T.nf
    assert ff[1].shape[0] == 3 * T.nf  # closed manifold: 3 neighbours per face
    # v in VV[u] <=> u in VV[v]
    s = O.csr_to_sets(vv)
    for u in range(0, T.nv, 17):
        for v in s[u]:
            assert u in s[v]
    # every VE edge has the vertex as an endpoint
    se = O.csr_to_sets(ve)
    for v in range(0, T.nv, 13):
        for e in se[v]:
            assert v in T.ev[e]


def test_laplacian_f32_vs_f64():
    g = load_golden("sphere3")
    T = O.Topology(g["F"])
    vv = T.query("VV")
    a = O.laplacian_step(vv, g["V"], 0.01, np.float32)
    b = O.laplacian_step(vv, g["V"].astype(np.float64), 0.01, np.float64)
    assert np.abs(a - b).max() < 1e-6
    # a Jacobi step contracts a closed convex shape towards its centroid
    assert np.linalg.norm(a, axis=1).mean() < np.linalg.norm(g["V"], axis=1).mean()


def test_bilateral_denoises_noisy_sphere():
    from rxmesh_b200 import meshio
    V, F = meshio.icosphere(8)
    rng = np.random.RandomState(0)
    noisy = (V * (1 + 0.01 * rng.randn(V.shape[0], 1))).astype(np.float32)
    T = O.Topology(F)
    out, worst = O.bilateral_step(T.query("VV"), F, noisy)
    assert worst <= 80
    # the filter shrinks a convex shape slightly (all neighbours lie below the tangent plane) but
    # must reduce the roughness: spread of the radius
    assert np.linalg.norm(out, axis=1).std() < 0.75 * np.linalg.norm(noisy, axis=1).std()


def test_all_cores_ports_match_serial():
    """rxo_vertex_normals_f32_mt / rxo_consume_sum_f32_mt (the all-cores CPU baseline of bench.py) == the serial forms:
    per-thread accumulators summed in thread order differ from the serial sum only by fp32 reassociation."""
    g = load_golden("dragon")
    ref = O.vertex_normals(g["F"], g["V"], np.float32)
    for t in (1, 3, 8):
        n, _ = O.vertex_normals_mt(g["F"], g["V"], t)
        assert np.abs(n - ref).max() < 2e-6 * np.abs(ref).max()
    assert np.array_equal(O.vertex_normals_mt(g["F"], g["V"], 1)[0].view(np.uint32), ref.view(np.uint32))
    T = O.Topology(g["F"])
    x = np.random.RandomState(0).rand(T.nv).astype(np.float32)
    a = O.consume_sum_mt(T.query("VV"), x, 4)
    assert np.allclose(a, O.consume_sum(T.query("VV"), x), rtol=1e-6)


def test_gaussian_curvature_gauss_bonnet():
    """Known-answer property of the restated GaussianCurvature accumulators (apps/GaussianCurvature): on a closed surface
    sum_v (2 pi - sum of incident angles) = 2 pi chi (Gauss-Bonnet), and the mixed areas add up to the surface area."""
    for name, chi in (("sphere3", 2), ("torus", 0), ("dragon", 0)):
        g = load_golden(name)
        V, F = g["V"], g["F"]
        gcs, amix = O.gaussian_curvature(F, V)
        assert abs((2 * np.pi + gcs).sum() - 2 * np.pi * chi) < 1e-6 * V.shape[0], name
        X = V.astype(np.float64)
        area = 0.5 * np.linalg.norm(np.cross(X[F[:, 1]] - X[F[:, 0]], X[F[:, 2]] - X[F[:, 0]]), axis=1).sum()
        assert abs(amix.sum() - area) < 1e-6 * area, name


def test_ev_diamond_and_ee_against_brute_force():
    """rxo-side EVDiamond / EE restatements (oracle.Topology.ev_diamond / ee) against an independent per-edge search."""
    for name in ("sphere3", "bunnyhead", "plane_5"):
        F = load_golden(name)["F"]
        T = O.Topology(F)
        evd, ee = T.ev_diamond(), T.ee()
        faces_of = {}
        for f in range(T.nf):
            for j in range(3):
                faces_of.setdefault(int(T.fe[f, j]), []).append((f, j))
        for e in range(T.ne):
            v0, v1 = int(T.ev[e, 0]), int(T.ev[e, 1])
            assert (evd[e, 0], evd[e, 2]) == (v0, v1)
            want_v = {0: 0xFFFFFFFF, 1: 0xFFFFFFFF}
            want_e = {0: (0xFFFFFFFF, 0xFFFFFFFF), 1: (0xFFFFFFFF, 0xFFFFFFFF)}
            for f, j in faces_of[e]:
                a, b, c = (int(F[f, j]), int(F[f, (j + 1) % 3]), int(F[f, (j + 2) % 3]))
                d = 0 if (a, b) == (v0, v1) else 1
                want_v[d] = c                                   # the vertex opposite to the edge in that face
                want_e[d] = (int(T.fe[f, (j + 1) % 3]), int(T.fe[f, (j + 2) % 3]))
            assert (evd[e, 1], evd[e, 3]) == (want_v[0], want_v[1]), (name, e)
            assert tuple(ee[e]) == want_e[0] + want_e[1], (name, e)


def test_mcf_matvec_constant_vector():
    """Known answer for the restated MCF mat-vec (apps/MCF/mcf_kernels.cuh:117-205): the cotangent part annihilates a
    constant vector whatever the time step, so out(p) = in / vw(p) with one positive factor per vertex shared by the three
    components (vw = 0.5 / sum of the partial Voronoi areas in the reference's own scaling, geometry_util.cuh:120-172)."""
    g = load_golden("sphere3")
    V, F = g["V"], g["F"]
    rings = O.oriented_rings(F, V.shape[0])
    c = np.array([0.3, -1.2, 2.0], np.float32)
    vin = np.tile(c, (V.shape[0], 1))
    out10, out1 = O.mcf_matvec(rings, V, vin, 10.0), O.mcf_matvec(rings, V, vin, 1.0)
    k = out10[:, 0] / np.float64(c[0])
    assert np.all(k > 0)
    assert np.allclose(out10, k[:, None] * vin.astype(np.float64), rtol=1e-7, atol=1e-9)
    assert np.allclose(out10, out1, rtol=1e-6, atol=1e-8)  # independent of the time step


def test_mt_ports_equal_the_serial_oracle():
    """The OpenMP forms used as CPU baselines (bilateral filter split over vertices like filtering_openmesh.h:112-116, the
    manual-smoothing step) give the serial oracle's result bit for bit, for any thread count."""
    from rxmesh_b200 import meshio
    V, F = meshio.torus(60, 44, noise=0.2)
    vv = O.Topology(F).query("VV")
    ref, _ = O.bilateral_step(vv, F, V, 80, 2)
    lap = O.laplacian_step(vv, V, 0.01, np.float32)
    for threads in (1, 3, 8):
        got, _ = O.bilateral_step_mt(vv, F, V, threads)
        assert np.array_equal(got, ref)
        got, _ = O.laplacian_step_mt(vv, V, 0.01, threads)
        assert np.array_equal(got, lap)
