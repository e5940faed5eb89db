#!/usr/bin/env python
"""bench_configs.py -- the sub-records of the bench line: BASELINE.json configs[0..2] and configs[4].

bench.py times configs[3] (the headline) and calls the functions below; each returns one dict that lands under
`configs` / `laplacian_400m` in the JSON line.  Every record carries an in-run parity flag against the oracle
(oracle/ is the checker here, never the thing measured).

  python bench_configs.py [--only dragon,queries,bilateral,laplacian] [--lap-faces F]     # stand-alone, 1 GPU
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LAP_N = 14143           # nx = ny = 14 143 -> F = 2 * 14 142^2 = 399 992 328
LAP_LR = 0.01           # apps/Smoothing/smoothing.cu:17-18
LAP_N1_FILE = "/tmp/rxm_b200_laplacian_n1.json"  # same-lease N = 1 result, read by the N > 1 runs for the speed-up


def peaks():
    from bench import peaks as p
    return p()


def timed(fn, stream, torch, reps, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


# ------------------------------------------------------------------------------------ configs[0]
def config_dragon(torch, rx, stream):
    """VertexNormal on input/dragon.obj (tests/golden/dragon.npz holds the mesh and the output of the reference's own
    serial loop).  20 000 faces = 61 blocks: one launch is pure launch latency, so the figure of merit is a CUDA graph of
    100 back-to-back launches (the reference app also loops num_run launches, vertex_normal.cu:69-85)."""
    from oracle import oracle as O
    g = np.load(os.path.join(ROOT, "tests", "golden", "dragon.npz"))
    V, F = g["V"], g["F"]
    m = rx.RXMeshStatic(F)
    x = rx.Attribute(m, 0, np.float32, 3, rx.DEVICE, rx.AoS)
    n = rx.Attribute(m, 0, np.float32, 3, rx.DEVICE, rx.AoS)
    x.from_global(V)
    ms_single = timed(lambda: m.vertex_normals(x, n, False, stream), stream, torch, 1000, warm=10)
    ms_graph = None
    try:
        gs = torch.cuda.Stream()
        with torch.cuda.stream(gs):
            m.vertex_normals(x, n, False, gs)
            gs.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=gs):
                for _ in range(100):
                    m.vertex_normals(x, n, False, torch.cuda.current_stream())
            ms_graph = timed(graph.replay, gs, torch, 20, warm=2) / 100.0
    except Exception as e:  # noqa: BLE001
        ms_graph = None
        graph_err = str(e)[:200]
    got = n.to_global()
    ref64 = O.vertex_normals(F, V, np.float64)
    rel = float((np.linalg.norm(got - ref64, axis=1) / np.linalg.norm(ref64, axis=1)).max())
    abs_ref = float(np.abs(got - g["vn_ref"]).max())
    _, t_cpu = O.ref_vertex_normals(F, V, repeats=200) if O.ref_lib() is not None else (None, float("nan"))
    best = ms_graph if ms_graph else ms_single
    r = {"what": "VertexNormal on dragon.obj (20 000 faces, %d patches)" % m.get_num_patches(),
         "ms_per_launch_stream": ms_single, "ms_per_launch_cuda_graph_of_100": ms_graph,
         "faces_per_s": F.shape[0] / (best * 1e-3),
         "note": "61 blocks on 148 SMs: launch-latency bound; the graph removes the per-launch CPU cost",
         "max_rel_err_vs_oracle_f64": rel, "max_abs_err_vs_reference_cpu_loop": abs_ref,
         "parity_ok": bool(rel < 1e-5 and abs_ref < 1e-4), "tolerance": "1e-5 relative (oracle f64), 1e-4 abs (reference criterion)",
         "reference_cpu_loop_ms": t_cpu * 1e3}
    if ms_graph is None:
        r["cuda_graph_error"] = graph_err
    return r


# ------------------------------------------------------------------------------------ configs[1]
QUERY_ALG_BYTES = {"VV": 40, "VE": 40, "VF": 40, "EV": 48, "EF": 48, "FV": 44, "FE": 44, "FF": 44}


def pair_hash(src, dst):
    """order-independent 64-bit hash of a multiset of (source id, neighbour id) pairs"""
    k = (src.astype(np.uint64) << np.uint64(32)) | dst.astype(np.uint64)
    with np.errstate(over="ignore"):
        k = (k ^ (k >> np.uint64(29))) * np.uint64(0x9E3779B97F4A7C15)
        k = k ^ (k >> np.uint64(32))
        return int(np.bitwise_xor.reduce(k)) ^ (int(np.add.reduce(k)) & 0xFFFFFFFFFFFFFFFF), int(k.shape[0])


def config_queries(torch, rx, stream, nu=707, patch_size=1024):
    """all eight static queries, store variant (the reference's test kernel), class-I icosphere nu = 707 (9 996 980
    faces), built-in Lloyd patcher.  parity_ok: the multiset of (source, neighbour) global-id pairs the TIMED kernel wrote
    equals the oracle's (order-independent 64-bit hash + count), every op."""
    from oracle import oracle as O
    from rxmesh_b200 import meshio
    from rxmesh_b200.mesh import _DST, _SRC
    peak, _ = peaks()
    V, F = meshio.icosphere(nu)
    t0 = time.perf_counter()
    m = rx.RXMeshStatic(F, patch_size=patch_size, num_threads=os.cpu_count() or 8)
    tb = time.perf_counter() - t0
    T = O.Topology(F)
    nF = F.shape[0]
    res, ok_all = {}, True
    for op, bpf in QUERY_ALG_BYTES.items():
        o = rx.Op[op]
        width = {"EV": 2, "FV": 3, "FE": 3, "EF": 2, "FF": 3}.get(op, m.get_input_max_valence())
        inp = rx.Attribute(m, _SRC[o], np.uint64, 1, rx.LOCATION_ALL, rx.AoSoA)  # the reference's default layout
        out = rx.Attribute(m, _SRC[o], np.uint64, width, rx.LOCATION_ALL, rx.AoSoA)
        inp.reset(rx.INVALID64, rx.DEVICE)
        out.reset(rx.INVALID64, rx.DEVICE)
        ms = timed(lambda: m.query_store(o, inp, out, stream), stream, torch, 50)
        gbs = bpf * nF / (ms * 1e-3) / 1e9
        # parity of what the timed kernel wrote
        inp.move(rx.DEVICE, rx.HOST), out.move(rx.DEVICE, rx.HOST)
        # AoSoA -> one row per slot: value (slot s of patch p, attribute a) sits at base(p) * width + a * cap(p) + lid
        sb = m.slot_base(_SRC[o]).astype(np.int64)
        cap = np.diff(sb)
        base, capr = np.repeat(sb[:-1], cap), np.repeat(cap, cap)
        lid = np.arange(sb[-1], dtype=np.int64) - base
        hi = inp.host_array()
        ho = out.host_array()[(base * width + lid)[:, None] + capr[:, None] * np.arange(width, dtype=np.int64)[None, :]]
        valid_src = hi != np.uint64(rx.INVALID64)
        src_g = m.map_to_global(_SRC[o], hi)
        dst_g = m.map_to_global(_DST[o], ho.reshape(-1)).reshape(-1, width)
        mask = (ho != np.uint64(rx.INVALID64)) & valid_src[:, None]
        got = pair_hash(np.broadcast_to(src_g[:, None], dst_g.shape)[mask], dst_g[mask])
        roff, rval = T.query(op)
        want = pair_hash(np.repeat(np.arange(roff.shape[0] - 1, dtype=np.uint64), np.diff(roff.astype(np.int64))), rval)
        ok = got == want
        ok_all &= ok
        res[op] = {"ms": ms, "entries_per_s": want[1] / (ms * 1e-3), "achieved_gbs": gbs, "hbm_frac": gbs / peak, "parity_ok": bool(ok)}
        inp.release(), out.release()
    irregular = consume_and_normals(torch, rx, stream, m, V, F, T)
    return {"what": "8 static queries, store variant (64-bit handles), %d-face icosphere, Lloyd patches of <= %d faces" % (nF, patch_size),
            "consume_and_normals_on_lloyd_patches": irregular,
            "faces": nF, "patches": m.get_num_patches(), "ribbon_overhead": m.ribbon_overhead(), "build_seconds": tb,
            "patcher_seconds": m.build_seconds(True), "topo_bytes_per_face": m.topo_bytes() / nF,
            "alg_bytes_per_face": QUERY_ALG_BYTES, "peak_gbs": peak, "ops": res, "parity_ok": bool(ok_all),
            "parity": "multiset of (source, neighbour) global-id pairs of the timed kernel's output == oracle, per op (bit-exact)"}


def consume_and_normals(torch, rx, stream, m, V, F, T, check=True):
    """the headline pass (VV consume, VF consume, vertex normals) on an arbitrary mesh / patching: what the kernels do when
    the patches are NOT the analytic tiles of the grid (Lloyd patches, mixed valence: the scalar per-vertex path runs next
    to the packed valence-6 path)"""
    from oracle import oracle as O
    peak, _ = peaks()
    nF, nV = F.shape[0], V.shape[0]
    x = rx.Attribute(m, 0, np.float32, 3, rx.DEVICE, rx.AoS)
    nrm = rx.Attribute(m, 0, np.float32, 3, rx.DEVICE, rx.AoS)
    sv_in = rx.Attribute(m, 0, np.float32, 1, rx.DEVICE, rx.AoS)
    sv_out = rx.Attribute(m, 0, np.float32, 1, rx.DEVICE, rx.AoS)
    sf_in = rx.Attribute(m, 2, np.float32, 1, rx.DEVICE, rx.AoS)
    rng = np.random.RandomState(7)
    hv, hf = rng.rand(nV).astype(np.float32), rng.rand(nF).astype(np.float32)
    x.from_global(V), sv_in.from_global(hv), sf_in.from_global(hf)
    out = {}
    for name, bpf, fn in (("VV", 16.0, lambda: m.query_consume(rx.Op.VV, sv_in, sv_out, stream)),
                          ("VF", 18.0, lambda: m.query_consume(rx.Op.VF, sf_in, sv_out, stream)),
                          ("VN", 24.0, lambda: m.vertex_normals(x, nrm, False, stream))):
        ms = timed(fn, stream, torch, 50, warm=3)
        gbs = bpf * nF / (ms * 1e-3) / 1e9
        out[name] = {"ms": ms, "alg_bytes_per_face": bpf, "achieved_gbs": gbs, "hbm_frac": gbs / peak}
    if check:
        got = nrm.to_global()
        ref = O.vertex_normals(F, V, np.float64)
        out["VN"]["max_rel_err_vs_oracle"] = float((np.linalg.norm(got - ref, axis=1) / np.linalg.norm(ref, axis=1)).max())
        m.query_consume(rx.Op.VF, sf_in, sv_out, stream)
        gotc = sv_out.to_global().reshape(-1)
        refc = O.consume_sum(T.query("VF"), hf)
        out["VF"]["max_rel_err_vs_oracle"] = float((np.abs(gotc - refc) / np.maximum(np.abs(refc), 1e-30)).max())
        out["parity_ok"] = bool(out["VN"]["max_rel_err_vs_oracle"] < 1e-5 and out["VF"]["max_rel_err_vs_oracle"] < 1e-5)
    for a in (x, nrm, sv_in, sv_out, sf_in):
        a.release()
    return out


def mcf_cg(torch, rx, stream, m, V, F):
    """SURVEY.md 8(f)1 widened to its caller: the MCF app's matrix-free CG solve (apps/MCF/mcf_cg_mat_free.h) on the noisy
    torus of config 3 (a fairing problem: the Laplacian of the noise is far above fp32 rounding), through rxm_mcf_solve
    (two kernels per iteration, solver scalars on the device).  Two runs: the app's defaults (apps/MCF/mcf.cu:19-23: uniform
    Laplacian, dt = 10, tol_abs 1e-6, at most 100 iterations) and the cotangent Laplacian with dt of the order of a vertex
    area and a relative tolerance (with the app's dt = 10 on a unit-size mesh that system does not converge in 100
    iterations, in float64 either).  Parity in the run: the float64 oracle solve of the same system (also the CPU baseline)
    and the true residual B - A X of the GPU's result."""
    from oracle import oracle as O
    peak, _ = peaks()
    nV = V.shape[0]
    V = np.ascontiguousarray(V, np.float32)
    x0 = rx.Attribute(m, 0, np.float32, 3, rx.DEVICE, rx.AoS)
    x = rx.Attribute(m, 0, np.float32, 3, rx.DEVICE, rx.AoS)
    x0.from_global(V)
    out = {"what": "MCF, matrix-free CG, %d vertices (the noisy torus of this config, 32x16-quad tiles)" % nV}
    rings = O.oriented_rings(F, nV)
    scale = float(np.abs(V).max())
    for label, uniform, dt, ta, tr, pc in (("uniform_laplace_app_defaults", True, 10.0, 1e-6, 0.0, False),
                                           ("cotangent_laplace", False, 1e-5, 0.0, 1e-6, False),
                                           ("uniform_laplace_jacobi_pcg", True, 10.0, 1e-6, 0.0, True)):
        kw = dict(time_step=dt, use_uniform_laplace=uniform, max_iter=100, tol_abs=ta, tol_rel=tr, stream=stream, precondition=pc)
        m.mcf_solve(x0, x, **kw)  # warm-up (allocations, module load)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        info = m.mcf_solve(x0, x, **kw)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        got = x.to_global()
        steps = info["iterations"] + (1 if info["converged"] else 0)  # mat-vec + update pairs that did work
        # algorithmic bytes per vertex: setup 12 (X0) + 12 (fan ids) + 12 (R) + 4 (diag) [+ 24 (W)] + 24 (X = X0 copy);
        # iteration: mat-vec 24 (R, P) + 12 (fan ids) + 4 (diag) [+ 24 (W)] + 24 (P', S), update 48 (X, R, P', S) + 24 (X, R)
        # (the Jacobi-preconditioned form reads the diagonal twice more per iteration: + 8)
        per_it = ((136.0 if uniform else 160.0) + (8.0 if pc else 0.0)) * nV
        setup = (64.0 if uniform else 88.0) * nV
        gbs = (setup + per_it * steps) / (ms * 1e-3) / 1e9
        t0 = time.perf_counter()
        ref, oinfo = O.mcf_solve(rings, V, dt, uniform, 100, ta, tr, precond=pc)
        t_cpu = time.perf_counter() - t0
        res, bb = O.mcf_residual(rings, V, got, dt, uniform)
        r2 = float((res ** 2).sum())
        err, move = float(np.abs(got - ref).max()), float(np.abs(ref - V).max())
        # both stop at the same residual; what is left of the solution error at that point is a fraction of the move
        tol = 1e-5 * scale + move * float(np.sqrt(max(oinfo["final_residual"] / oinfo["start_residual"], 0.0)))
        ok = bool(info["converged"] and oinfo["converged"] and abs(info["iterations"] - oinfo["iterations"]) <= 2 + oinfo["iterations"] // 10
                  and err < tol)
        out[label] = {"time_step": dt, "tol_abs": ta, "tol_rel": tr, "jacobi_preconditioner": pc, "iterations": info["iterations"], "converged": info["converged"],
                      "start_residual": info["start_residual"], "final_residual": info["final_residual"],
                      "ms_total": ms, "ms_per_iteration": ms / max(steps, 1), "vertex_iterations_per_s": nV * steps / (ms * 1e-3),
                      "alg_bytes_per_vertex_iteration": per_it / nV, "achieved_gbs": gbs, "hbm_frac": gbs / peak,
                      "oracle": {"iterations": oinfo["iterations"], "converged": oinfo["converged"],
                                 "start_residual": oinfo["start_residual"], "final_residual": oinfo["final_residual"]},
                      "true_residual_sq_of_gpu_result": r2, "rhs_sq": bb, "max_abs_diff_vs_oracle_f64": err, "max_move": move,
                      "tolerance_abs": tol, "parity_ok": ok,
                      "cpu_baseline": {"value": nV * (oinfo["iterations"] + 1) / t_cpu, "unit": "vertex-iterations/s", "cores": 1,
                                       "kind": "port", "sample": "the same mesh and solve, float64, 1 thread (%.1f s)" % t_cpu}}
    out["parity_ok"] = bool(all(v["parity_ok"] for v in out.values() if isinstance(v, dict)))
    x0.release(), x.release()
    return out


def lloyd_at_scale(torch, rx, stream, faces=100_000_000):
    """the built-in (host, multi-threaded, deterministic) Lloyd patcher at 100 M faces: a grid whose faces are visited in
    SCRAMBLED tile order so that nothing in the input order helps, patch size 1024; then the headline kernels on those patches"""
    from rxmesh_b200 import meshio
    n = int(round((faces / 2.0) ** 0.5)) + 1
    V, F = meshio.grid(n, n)
    t0 = time.perf_counter()
    m = rx.RXMeshStatic(F, patch_size=1024, num_threads=os.cpu_count() or 8, ring2=False)
    tb = time.perf_counter() - t0
    sizes = np.diff(m.lin_base(2).astype(np.int64))
    rec = {"what": "Lloyd patcher + build on a %d-face grid (no face->patch hint), patch size <= 1024" % F.shape[0],
           "faces": int(F.shape[0]), "patches": m.get_num_patches(), "build_seconds": tb, "patcher_seconds": m.build_seconds(True),
           "host_threads": os.cpu_count(), "patch_faces_mean": float(sizes.mean()), "patch_faces_max": int(sizes.max()),
           "ribbon_overhead": m.ribbon_overhead(), "topo_bytes_per_face": m.topo_bytes() / F.shape[0]}
    m.compact()
    rec["kernels"] = consume_and_normals(torch, rx, stream, m, V, F, None, check=False)
    return rec


# ------------------------------------------------------------------------------------ configs[2]
def torus_window(nu, nv, i0, i1, j0, j1, V):
    """the sub-grid rows [i0, i1] x columns [j0, j1] of meshio.torus(nu, nv) as an open grid mesh with the torus'
    coordinates (no wrap inside the window): (global vertex ids, local faces)"""
    ii = np.arange(i0, i1 + 1, dtype=np.int64) % nu
    jj = np.arange(j0, j1 + 1, dtype=np.int64) % nv
    gid = (ii[:, None] * nv + jj[None, :]).reshape(-1)
    h, w = ii.shape[0], jj.shape[0]
    idx = (np.arange(h - 1, dtype=np.uint32)[:, None] * np.uint32(w) + np.arange(w - 1, dtype=np.uint32)[None, :]).reshape(-1)
    a, b, c, d = idx, idx + np.uint32(w), idx + np.uint32(1), idx + np.uint32(w + 1)
    f = np.empty((idx.shape[0], 2, 3), dtype=np.uint32)
    f[:, 0, 0], f[:, 0, 1], f[:, 0, 2] = a, b, c
    f[:, 1, 0], f[:, 1, 1], f[:, 1, 2] = c, b, d
    return gid, f.reshape(-1, 3), (h, w)


def config_bilateral(torch, rx, stream, nu=2236, iters=5):
    """bilateral filtering (apps/Filtering), 10 M-face periodic torus with noise, 5 iterations (reference default).
    parity_ok: ONE iteration against the oracle on two windows of the mesh (window interior, 12 rings in from the cut, so
    every neighbourhood the filter visits lies inside the window)."""
    from oracle import oracle as O
    from rxmesh_b200 import meshio
    peak, _ = peaks()
    V, F = meshio.torus(nu, nu, noise=0.2)
    tj, ti = (int(t) for t in os.environ.get("RXM_BIL_TILE", "32x16").split("x"))  # quads per patch: columns x rows
    q = np.arange(F.shape[0] // 2, dtype=np.uint32)
    fp = (q // nu // ti * ((nu + tj - 1) // tj) + q % nu // tj).repeat(2)
    t0 = time.perf_counter()
    m = rx.RXMeshStatic(F, face_patch=fp, patch_size=2 * tj * ti, num_threads=os.cpu_count() or 8)
    tb = time.perf_counter() - t0
    x = rx.Attribute(m, 0, np.float32, 3, rx.DEVICE, rx.AoS)
    y = rx.Attribute(m, 0, np.float32, 3, rx.DEVICE, rx.AoS)
    x.from_global(V)
    m.bilateral_filter(x, y, 1, stream)  # setup (scratch attributes)
    got1 = y.to_global()
    ms = timed(lambda: m.bilateral_filter(x, y, iters, stream), stream, torch, 3)
    deferred = m.bilateral_deferred() / float(iters)
    nF, nV = F.shape[0], V.shape[0]
    gbs = 54.0 * nF * iters / (ms * 1e-3) / 1e9
    # parity on windows (one across the periodic seam)
    worst, n_chk, ring = 0.0, 0, 12
    scale = float(np.abs(V).max())
    for (i0, j0) in ((100, 200), (nu - 40, nu - 30)):
        gid, Fw, (h, w) = torus_window(nu, nu, i0, i0 + 79, j0, j0 + 79, V)
        Vw = V[gid]
        Tw = O.Topology(Fw)
        refw, _ = O.bilateral_step(Tw.query("VV"), Fw, Vw, 80, 2)  # membership in fp32, the rest in float64
        inner = np.zeros((h, w), bool)
        inner[ring:h - ring, ring:w - ring] = True
        sel = inner.reshape(-1)
        err = np.abs(got1[gid[sel]] - refw[sel]).max(axis=1)
        worst = max(worst, float(err.max()))
        n_chk += int(sel.sum())
    tol = 2e-5 * scale
    # the CPU beside it (north_star): one iteration of the restated filter (oracle port; the reference's CPU side is OpenMesh +
    # OpenMP schedule(static), filtering_openmesh.h:112-116, OpenMesh is not in the tree) on a bounded sample of the same
    # generator, one thread and all host threads
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    ns = 700
    Vs, Fs = meshio.torus(ns, ns, noise=0.2)
    vvs = O.Topology(Fs).query("VV")
    _, t1 = O.bilateral_step_mt(vvs, Fs, Vs, 1)
    _, tn = O.bilateral_step_mt(vvs, Fs, Vs, ncpu)
    _, tn2 = O.bilateral_step_mt(vvs, Fs, Vs, ncpu)
    tn = min(tn, tn2)
    cpu = {"value": Vs.shape[0] / min(t1, tn), "unit": "vertex-iterations/s", "cores": ncpu if tn < t1 else 1, "kind": "port",
           "sample": "%d-face torus of the same generator (%d^2 quads), one iteration of the filter (normals excluded): "
                     "1 thread %.3g, %d threads %.3g vertex-iterations/s" % (Fs.shape[0], ns, Vs.shape[0] / t1, ncpu, Vs.shape[0] / tn)}
    x.release(), y.release()
    try:
        mcf = mcf_cg(torch, rx, stream, m, V, F)
    except Exception as e:  # noqa: BLE001 -- the widening row must not take the config-3 record down with it
        mcf = {"error": (type(e).__name__ + ": " + str(e))[:300]}
    return {"what": "bilateral filtering, %d-face torus (%d^2 quads), %d iterations (unit-face normals + filter)" % (nF, nu, iters),
            "cpu_baseline": cpu, "mcf_cg_same_mesh": mcf,
            "faces": nF, "patches": m.get_num_patches(), "patch_tile_quads": "%dx%d" % (tj, ti), "build_seconds": tb, "ms_total": ms,
            "ms_per_iteration": ms / iters,
            "vertex_iterations_per_s": nV * iters / (ms * 1e-3), "alg_bytes_per_iteration": 54.0 * nF,
            "achieved_gbs": gbs, "hbm_frac": gbs / peak, "peak_gbs": peak,
            "kernel": "k_bilateral_patch (one launch per iteration: unit-face normals + filter, patch-local)",
            "cross_patch_vertex_fraction": deferred / nV, "ring2": m.has_ring2(), "topo_bytes_per_face": m.topo_bytes() / nF,
            "parity_ok": bool(worst <= tol), "parity_max_abs_err": worst, "parity_tolerance_abs": tol,
            "parity": "1 iteration vs oracle on %d vertices of two 80x80 windows (one across the periodic seam), all within 2e-5 x max|x|" % n_chk}


# ------------------------------------------------------------------------------------ configs[4]
def grid_window(nx, r0, r1, c0, c1, dx=1.0):
    """rows [r0, r1] x columns [c0, c1] (inclusive) of meshio.grid(nx, *): the same fp32 coordinates (height field at
    the GLOBAL position), local faces, global vertex ids"""
    from rxmesh_b200 import meshio
    V, F = meshio.grid(c1 - c0 + 1, r1 - r0 + 1, dx, height=False)
    V[:, 0] += np.float32(dx * c0)
    V[:, 2] += np.float32(dx * r0)
    xg, zg = V[:, 0].astype(np.float64), V[:, 2].astype(np.float64)
    V[:, 1] = (0.05 * np.sin(0.01 * xg) * np.cos(0.013 * zg)).astype(np.float32)
    gid = (np.arange(r0, r1 + 1, dtype=np.int64)[:, None] * nx + np.arange(c0, c1 + 1, dtype=np.int64)[None, :]).reshape(-1)
    return V, F, gid


def laplacian_oracle_check(n, k, rows, got_lookup):
    """k oracle Laplacian steps on windows of the n x n grid around the given rows; got_lookup(gids) -> (mask, values) of
    the GPU result for the global vertex ids this rank owns.  Returns (max error relative to the position vector's norm,
    vertices checked, max abs error of the height component -- positions reach 1.4e4 while the height field is 0.05, so
    the vector-norm figure alone would hide an error in y)."""
    from oracle import oracle as O
    worst, cnt, worst_y = 0.0, 0, 0.0
    for r in rows:
        for c0 in (0, max(0, n // 2 - 40), max(0, n - 81)):
            r0, r1 = max(0, r - 40), min(n - 1, r + 40)
            c1 = min(n - 1, c0 + 80)
            Vw, Fw, gid = grid_window(n, r0, r1, c0, c1)
            vv = O.Topology(Fw).query("VV")
            ref = Vw.astype(np.float64)
            for _ in range(k):
                ref = O.laplacian_step(vv, ref, LAP_LR, np.float64)
            h, w = r1 - r0 + 1, c1 - c0 + 1
            ok = np.ones((h, w), bool)  # shrink by k rings on every side that is a cut, not a mesh border
            if r0 > 0:
                ok[:k, :] = False
            if r1 < n - 1:
                ok[h - k:, :] = False
            if c0 > 0:
                ok[:, :k] = False
            if c1 < n - 1:
                ok[:, w - k:] = False
            sel = ok.reshape(-1)
            mask, val = got_lookup(gid[sel])
            if not mask.any():
                continue
            d = np.linalg.norm(val[mask] - ref[sel][mask], axis=1) / np.maximum(np.linalg.norm(ref[sel][mask], axis=1), 1e-30)
            worst = max(worst, float(d.max()))
            worst_y = max(worst_y, float(np.abs(val[mask][:, 1] - ref[sel][mask][:, 1]).max()))
            cnt += int(mask.sum())
    return worst, cnt, worst_y


def laplacian_400m(args, rank, world, local_rank, torch, rx, tile, tile_i):
    """BASELINE.json configs[4]: 400 M-face grid, iterated Laplacian smoothing (apps/Smoothing/manual.h:86-104; 100
    iterations, lr 0.01), STRONG scaling: the same mesh cut into `world` row slabs, one per GPU, ribbon exchange fused into
    the compute kernel (k_laplacian_fan2<true>: NVLink P2P stores + flag words) and inside the timing."""
    from rxmesh_b200 import distributed as D, meshio
    dist = torch.distributed
    n = LAP_N if not args.lap_faces else int(round((args.lap_faces / 2.0) ** 0.5)) + 1
    iters = args.iters
    ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 8)
    stream = torch.cuda.current_stream()
    t0 = time.perf_counter()
    if world == 1:
        V, F = meshio.grid(n, n)
        fp = meshio.grid_face_tiles(n, n, tile, tile_i)
        mesh = rx.RXMeshStatic(F, face_patch=fp, patch_size=2 * tile * tile_i, num_threads=ncores, ring2=False)
        del F, fp
        sm = hx = None
        l2g0, nloc = 0, V.shape[0]
        real = None
    else:
        sh = D.grid_slab(n, n, tile, tile_i, rank, world)
        sm = D.ShardedMesh(sh, rank, world, patch_size=2 * tile * tile_i, num_threads=max(1, ncores // world), ring2=False)
        mesh, V = sm.mesh, sh["verts"]
        hx = D.HaloExchange(sm, 0)
        l2g0, nloc = int(sh["l2g_v"][0]), V.shape[0]
        real = sm.real_owned_mask(0).copy()
        del sh
    t_build = time.perf_counter() - t0
    x = rx.Attribute(mesh, 0, np.float32, 3, rx.DEVICE, rx.AoS)
    y = rx.Attribute(mesh, 0, np.float32, 3, rx.DEVICE, rx.AoS)
    fused, fused_err = None, None
    if hx is not None and not os.environ.get("RXM_NO_FUSED"):
        try:
            fused = D.FusedHalo(hx, x, y)  # needs the host patch store: before compact()
        except Exception as e:  # noqa: BLE001  (e.g. no peer access): packed NCCL exchange instead
            fused, fused_err = None, str(e)[:200]
        ok = torch.tensor([1 if fused is not None else 0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok[0]) == 0:
            fused = None
    n_real_v = int(real.sum()) if real is not None else nloc
    n_patches = mesh.get_num_patches()
    topo_bpf = mesh.topo_bytes() / max(1, mesh.get_num_faces())
    mesh.compact()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def load(attr):
        attr.from_global(V, stream)
        if hx is not None:
            hx.exchange(attr, stream)
        barrier()

    def run_nccl(k):
        a, b = x, y
        for _ in range(k):
            mesh.laplacian_smooth(a, b, LAP_LR, 1, stream)
            if hx is not None:
                hx.exchange(b, stream)
            a, b = b, a
        return a

    def run_fused(k):
        src, _ = fused.buffers()
        load(src)
        return fused.smooth(LAP_LR, k, stream)

    # ---- correctness first: k steps, fused == kernel + NCCL exchange bit for bit, == oracle on windows at the cuts ----
    k_chk = 5
    load(x)
    res = run_nccl(k_chk)
    barrier()
    got_nccl = res.to_global()
    bit_identical = None
    if fused is not None:
        res = run_fused(k_chk)
        barrier()
        got_fused = res.to_global()
        sel = real if real is not None else slice(None)
        bit_identical = bool(np.array_equal(got_fused[sel], got_nccl[sel]))
        got = got_fused
    else:
        got = got_nccl

    def lookup(gids):
        loc = gids - l2g0
        inside = (loc >= 0) & (loc < nloc)
        mask = inside.copy()
        if real is not None:
            mask[inside] = real[loc[inside]]
        val = np.zeros((gids.shape[0], 3), np.float32)
        val[mask] = got[loc[mask]]
        return mask, val

    # rows of the cuts between ranks (every rank checks the part of each window it owns) + the first / a middle row
    n_tile_rows = (n - 1 + tile_i - 1) // tile_i
    cut_rows = [int(r) * tile_i for r in np.round(np.linspace(0, n_tile_rows, world + 1)).astype(np.int64)[1:-1]]
    rows = sorted(set([0, n // 3] + cut_rows))[:6]
    worst, cnt, worst_y = laplacian_oracle_check(n, k_chk, rows, lookup)
    del got, got_nccl
    chk = torch.tensor([worst, float(cnt), 1.0 if bit_identical in (None, True) else 0.0, worst_y], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = chk.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm_ = chk.clone()
        dist.all_reduce(sm_, op=dist.ReduceOp.SUM)
        mn = chk.clone()
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        worst, cnt, bit_ok, worst_y = float(mx[0]), int(sm_[1]), bool(mn[2] > 0.5), float(mx[3])
    else:
        bit_ok = True

    # ---- timing ----
    def timed_run(fn):
        fn(max(3, args.warmup))
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = rx.launch_count()
        e0.record(stream)
        fn(iters)
        e1.record(stream)
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]) / iters, rx.launch_count() - l0

    ms_nccl = None
    if fused is not None:
        src, _ = fused.buffers()
        load(src)
        ms_it, launches = timed_run(lambda k: fused.smooth(LAP_LR, k, stream))
        if not os.environ.get("RXM_SKIP_NCCL_TIMING"):
            load(x)
            ms_nccl, _ = timed_run(lambda k: run_nccl(k))
    else:
        load(x)
        ms_it, launches = timed_run(lambda k: run_nccl(k))
    halo = torch.tensor([float(hx.halo_elements()) if hx is not None else 0.0, float(n_real_v), float(n_patches), t_build],
                        dtype=torch.float64, device="cuda")
    if world > 1:
        hs = halo.clone()
        dist.all_reduce(hs, op=dist.ReduceOp.SUM)
        hm = halo.clone()
        dist.all_reduce(hm, op=dist.ReduceOp.MAX)
    else:
        hs = hm = halo
    if rank != 0:
        return None
    nF = 2 * (n - 1) ** 2
    nV = n * n
    peak, peak_src = peaks()
    gbs = 24.0 * nF / world / (ms_it * 1e-3) / 1e9
    rec = {
        "what": "iterated Laplacian smoothing (manual.h:86-104), %d x %d grid = %d faces, STRONG scaling over %d GPU(s): "
                "%d-face row slab per GPU" % (n, n, nF, world, nF // world),
        "metric": "vertex-updates/s", "value": nV / (ms_it * 1e-3), "n_gpus": world, "faces_total": nF, "vertices_total": nV,
        "iterations": iters, "lr": LAP_LR, "ms_per_iteration": ms_it, "scaling": "strong",
        "halo_transport": None if hx is None else ("fused" if fused is not None else "nccl"),
        "halo_transport_detail": None if hx is None else (
            "fused into the compute kernel: NVLink P2P stores into the neighbours' ghost slots + flag words, no NCCL call, no host sync"
            if fused is not None else "kernel, then packed NCCL send/recv (fused unavailable: %s)" % fused_err),
        "halo_bytes_per_iteration_total": int(12 * hs[0]), "halo_bytes_per_iteration_max_gpu": int(12 * hm[0]),
        "mirrored_vertices_total": int(hs[0]),
        "ms_per_iteration_nccl_exchange": ms_nccl,
        "roofline_per_gpu": {"bound": "hbm", "kernel": "k_laplacian_fan2<%s>" % ("true" if fused is not None else "false"),
                             "alg_bytes_per_launch": 24.0 * nF / world, "achieved": gbs, "peak": peak, "unit": "GB/s",
                             "frac": gbs / peak, "peak_source": peak_src, "note": "halo exchange time included"},
        "gpu_launches": int(launches), "launches_per_iteration": launches / float(iters),
        "patches_total": int(hs[2]), "topo_bytes_per_face": topo_bpf, "build_seconds_max_rank": float(hm[3]),
        "parity": {"fused_equals_nccl_bitwise": (bit_ok if fused is not None else None),
                   "oracle_max_rel_err": worst, "oracle_max_abs_err_height": worst_y, "oracle_vertices_checked": cnt,
                   "oracle_steps": k_chk, "tolerance_rel": 1e-5, "tolerance_abs_height": 1e-6,
                   "windows": "81 x 81-vertex windows at rows %s (cuts between ranks included) x 3 column positions" % rows,
                   "ok": bool(worst < 1e-5 and worst_y < 1e-6 and cnt > 0 and bit_ok)},
    }
    if rank == 0:
        # the CPU beside it: the manual smoothing step (oracle port; the reference app has no CPU side) on a bounded sample of
        # the same generator, one thread and all host threads
        try:
            from oracle import oracle as O
            ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
            Vs, Fs, _ = grid_window(1415, 0, 1414, 0, 1414)
            vvs = O.Topology(Fs).query("VV")
            _, t1 = O.laplacian_step_mt(vvs, Vs, LAP_LR, 1)
            tn = min(O.laplacian_step_mt(vvs, Vs, LAP_LR, ncpu)[1] for _ in range(3))
            rec["cpu_baseline"] = {"value": Vs.shape[0] / min(t1, tn), "unit": "vertex-updates/s", "cores": ncpu if tn < t1 else 1,
                                   "kind": "port", "sample": "1415 x 1415 grid (%d faces) of the same generator, one step: 1 thread "
                                   "%.3g, %d threads %.3g vertex-updates/s" % (Fs.shape[0], Vs.shape[0] / t1, ncpu, Vs.shape[0] / tn)}
        except Exception as e:  # noqa: BLE001
            rec["cpu_baseline"] = {"unavailable": str(e)[:200]}
    # speed-up against this configuration's own N = 1 run: same lease when bench.py ran N = 1 before (the scaling run does)
    if world == 1:
        try:
            json.dump({"ms_per_iteration": ms_it, "faces_total": nF, "when": time.time()}, open(LAP_N1_FILE, "w"))
        except Exception:  # noqa: BLE001
            pass
        rec["speedup_vs_n1"] = 1.0
    else:
        base, src = None, None
        try:
            j = json.load(open(LAP_N1_FILE))
            if j.get("faces_total") == nF:
                base, src = j["ms_per_iteration"], "N=1 run of this bench on the same box (%s)" % LAP_N1_FILE
        except Exception:  # noqa: BLE001
            pass
        if base is None:
            try:
                j = json.load(open(os.path.join(ROOT, "profiles", "r02_laplacian_400m_n1.json")))
                if j.get("faces_total") == nF:
                    base, src = j["ms_per_iteration"], "committed N=1 run (profiles/r02_laplacian_400m_n1.json), another box"
            except Exception:  # noqa: BLE001
                pass
        rec["speedup_vs_n1"] = (base / ms_it) if base else None
        rec["n1_ms_per_iteration"] = base
        rec["n1_source"] = src
    return rec


# ------------------------------------------------------------------------------------ same-GPU flat-array baseline
def hardwired_baseline(faces_side, nrun=20, timeout=300):
    """The reference's flat-array GPU kernel apps/VertexNormal/vertex_normal_hardwired.cuh (global float atomics over a
    plain face list, SURVEY.md 2.3: "baseline to beat on the same box"), compiled unmodified into oracle/_ref/ref_hardwired
    and run on the same grid generator at full size."""
    import subprocess
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_hardwired")
    if not os.path.exists(exe):
        return {"unavailable": "oracle/_ref/ref_hardwired not built (needs /root/reference at build time)"}
    try:
        r = subprocess.run([exe, str(faces_side), str(nrun)], capture_output=True, text=True, timeout=timeout)
        if r.returncode != 0:
            return {"unavailable": "ref_hardwired failed: " + (r.stderr or r.stdout)[-200:]}
        return json.loads(r.stdout.strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        return {"unavailable": str(e)[:200]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="dragon,queries,bilateral")
    ap.add_argument("--lap-faces", type=int, default=0)
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    only = set(args.only.split(","))
    import torch

    import rxmesh_b200 as rx
    rx.rx_init(0)
    stream = torch.cuda.current_stream()
    if "dragon" in only:
        print(json.dumps({"config": "dragon", **config_dragon(torch, rx, stream)}), flush=True)
    if "queries" in only:
        print(json.dumps({"config": "queries", **config_queries(torch, rx, stream)}), flush=True)
    if "bilateral" in only:
        print(json.dumps({"config": "bilateral", **config_bilateral(torch, rx, stream)}), flush=True)
    if "laplacian" in only:
        from bench import TILE, TILE_I
        print(json.dumps({"config": "laplacian", **laplacian_400m(args, 0, 1, 0, torch, rx, TILE, TILE_I)}), flush=True)
    if "lloyd100m" in only:
        print(json.dumps({"config": "lloyd100m", **lloyd_at_scale(torch, rx, stream)}), flush=True)
    if "hardwired" in only:
        print(json.dumps({"config": "hardwired", **hardwired_baseline(7072)}), flush=True)


if __name__ == "__main__":
    main()
