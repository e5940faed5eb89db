#!/usr/bin/env python
"""bench_configs.py -- the other BASELINE.json configs (bench.py runs configs[3], the headline).

  python bench_configs.py [--only 0,1,2]     # 1 GPU: dragon normals, 10M-sphere queries, 10M-torus bilateral
  torchrun ... bench.py --workload laplacian --faces 50000000 --gpus 8    # configs[4] (400 M faces total)

One JSON line per config; achieved GB/s use the ALGORITHMIC bytes of SURVEY.md 8(d) and the measured HBM peak.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from bench import peaks  # noqa: E402


def timed(fn, stream, torch, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="0,1,2")
    args = ap.parse_args()
    only = {int(x) for x in args.only.split(",")}
    import torch

    import rxmesh_b200 as rx
    from oracle import oracle as O
    from rxmesh_b200 import meshio
    from rxmesh_b200.mesh import _SRC

    rx.rx_init(0)
    stream = torch.cuda.current_stream()
    peak, peak_src = peaks()

    if 0 in only:  # VertexNormal on input/dragon.obj
        g = np.load(os.path.join(ROOT, "tests", "golden", "dragon.npz"))
        V, F = g["V"], g["F"]
        m = rx.RXMeshStatic(F)
        x = rx.Attribute(m, 0, np.float32, 3, rx.DEVICE, rx.AoS)
        n = rx.Attribute(m, 0, np.float32, 3, rx.DEVICE, rx.AoS)
        x.from_global(V)
        ms = timed(lambda: m.vertex_normals(x, n, False, stream), stream, torch, 1000)
        got = n.to_global()
        ref32 = g["vn_ref"]
        _, t_cpu = O.ref_vertex_normals(F, V, repeats=200)
        print(json.dumps({"config": 0, "what": "VertexNormal on dragon.obj (20 000 faces, 61 patches)", "ms": ms,
                          "faces_per_s": F.shape[0] / (ms * 1e-3), "note": "launch-latency bound at this size",
                          "max_abs_err_vs_reference_cpu_loop": float(np.abs(np.abs(got) - np.abs(ref32)).max()),
                          "reference_cpu_loop_ms (oracle/_ref, 1 thread)": t_cpu * 1e3}), flush=True)

    if 1 in only:  # all eight queries, store variant, 10 M-face sphere
        V, F = meshio.icosphere(707)
        t0 = time.perf_counter()
        m = rx.RXMeshStatic(F, patch_size=1024)
        tb = time.perf_counter() - t0
        nF = F.shape[0]
        alg = {"VV": 40, "VE": 40, "VF": 40, "EV": 48, "EF": 48, "FV": 44, "FE": 44, "FF": 44}
        res = {}
        for op, bpf in alg.items():
            o = rx.Op[op]
            width = {"EV": 2, "FV": 3, "FE": 3, "EF": 2, "FF": 5}.get(op, m.get_input_max_valence())
            inp = rx.Attribute(m, _SRC[o], np.uint64, 1, rx.DEVICE, rx.AoSoA)
            out = rx.Attribute(m, _SRC[o], np.uint64, width, rx.DEVICE, rx.AoSoA)
            inp.reset(rx.INVALID64, rx.DEVICE)
            out.reset(rx.INVALID64, rx.DEVICE)
            ms = timed(lambda: m.query_store(o, inp, out, stream), stream, torch, 50)
            gbs = bpf * nF / (ms * 1e-3) / 1e9
            res[op] = {"ms": ms, "entries_per_s": 3.0 * nF / (ms * 1e-3), "achieved_gbs": gbs, "hbm_frac": gbs / peak}
            inp.release(), out.release()
        print(json.dumps({"config": 1, "what": "8 static queries, store variant (64-bit handles), 10M-face icosphere, Lloyd patches of <= 1024 faces",
                          "faces": nF, "patches": m.get_num_patches(), "ribbon_overhead": m.ribbon_overhead(),
                          "build_seconds": tb, "packed": m.is_packed(), "fans": m.has_fans(), "peak_gbs": peak, "ops": res}), flush=True)
        del m

    if 2 in only:  # bilateral filtering, 10 M-face torus, 5 iterations
        nu = 2236
        V, F = meshio.torus(nu, nu, noise=0.2)
        fp = meshio.torus_face_tiles(nu, nu, 32)
        fp = (np.arange(F.shape[0] // 2, dtype=np.uint32) // nu // 16 * ((nu + 31) // 32) +
              np.arange(F.shape[0] // 2, dtype=np.uint32) % nu // 32).repeat(2)
        t0 = time.perf_counter()
        m = rx.RXMeshStatic(F, face_patch=fp, patch_size=1024)
        tb = time.perf_counter() - t0
        x = rx.Attribute(m, 0, np.float32, 3, rx.DEVICE, rx.AoS)
        y = rx.Attribute(m, 0, np.float32, 3, rx.DEVICE, rx.AoS)
        x.from_global(V)
        m.bilateral_filter(x, y, 1, stream)  # builds the VV CSR (setup)
        iters = 5
        ms = timed(lambda: m.bilateral_filter(x, y, iters, stream), stream, torch, 3)
        nF, nV = F.shape[0], V.shape[0]
        gbs = 54.0 * nF * iters / (ms * 1e-3) / 1e9
        print(json.dumps({"config": 2, "what": "bilateral filtering, 10M-face torus (2236^2 quads), 5 iterations (normals + filter)",
                          "faces": nF, "patches": m.get_num_patches(), "build_seconds": tb, "ms_total": ms, "ms_per_iteration": ms / iters,
                          "vertex_iterations_per_s": nV * iters / (ms * 1e-3), "alg_bytes_per_iteration": 54.0 * nF,
                          "achieved_gbs": gbs, "hbm_frac": gbs / peak, "peak_gbs": peak}), flush=True)


if __name__ == "__main__":
    main()
