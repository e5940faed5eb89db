#!/usr/bin/env python
"""bench.py -- BASELINE.json's headline workload on B200.

Workload (BASELINE.json configs[3]): a 100 M-face procedurally generated
grid (create_plane semantics, nx = ny = 7072, height field) on 1 GPU.  One "step" = one pass of the
hot path over the mesh: VV query (consume), VF query (consume), vertex normals -- three of our
sm_100a kernels.  `value` = faces/s through the whole pass (F / step time); the per-kernel rates
(VV / VF neighbour entries/s, vertex-normal faces/s) and their HBM-roofline fractions are in
`kernels`; `roofline` describes the dominant (longest) kernel.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--faces F]

N > 1 (launched with torch.distributed.run): weak scaling, every rank owns a mesh of `--faces`
faces (see DESIGN.md "Multi-GPU").
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

ALG_BYTES_PER_FACE = {"VV": 16.0, "VF": 18.0, "VN": 24.0}  # SURVEY.md 8(d), BASELINE.md 2
# 32 x 16 quads = 1024 owned faces per patch: the B200-tuned patch size (profiles/r01_tile_sweep.txt;
# the reference's default of 512 = 16 x 16 runs the same kernels at 44-53 % instead of 62-85 % of roofline)
TILE = 32
TILE_I = 16
if os.environ.get("RXM_TILE"):  # experiment knob: "<columns>x<rows>" quads per patch
    TILE, TILE_I = (int(v) for v in os.environ["RXM_TILE"].split("x"))


def grid_side(faces):
    return int(round(math.sqrt(faces / 2.0))) + 1


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t_begin=None, t_end=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = self.lines
        if t_begin is not None:
            win = [r for r in rows if t_begin - 0.02 <= r[0] <= t_end + 0.12]
            if not win and rows:  # region shorter than the sampling period: nearest sample
                win = [min(rows, key=lambda r: abs(r[0] - 0.5 * (t_begin + t_end)))]
            rows = win
        for _, ln in rows:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------ CPU arm
def cpu_sample_mesh(faces):
    from rxmesh_b200 import meshio
    n = grid_side(faces)
    V, F = meshio.grid(n, n)
    return V, F, n


def cpu_pass_seconds(V, F, T, vv, vf, repeats, threads=1):
    """One CPU pass = VV consume + VF consume over cached CSR adjacency (oracle port) + vertex normals.
    threads == 1: the reference's own serial vertex-normal loop (oracle/_ref when present, else the oracle port), as the
    reference app runs it.  threads > 1: the all-cores OpenMP ports (rxo_*_mt, per-thread accumulators)."""
    from oracle import oracle as O
    rng = np.random.RandomState(1)
    fvals = rng.rand(T.nv).astype(np.float32)
    ffvals = rng.rand(T.nf).astype(np.float32)
    have_ref = O.ref_lib() is not None
    if threads > 1:
        t_c = time.perf_counter()
        for _ in range(repeats):
            O.consume_sum_mt(vv, fvals, threads)
            O.consume_sum_mt(vf, ffvals, threads)
        t_c = (time.perf_counter() - t_c) / repeats
        _, t_n = O.vertex_normals_mt(F, V, threads, repeats)
        return t_c + t_n, t_n, have_ref
    t_c = time.perf_counter()
    for _ in range(repeats):
        O.consume_sum(vv, fvals)
        O.consume_sum(vf, ffvals)
    t_c = (time.perf_counter() - t_c) / repeats
    if have_ref:
        _, t_n = O.ref_vertex_normals(F, V, repeats=repeats)
    else:
        t_n = time.perf_counter()
        for _ in range(repeats):
            O.vertex_normals(F, V, np.float32)
        t_n = (time.perf_counter() - t_n) / repeats
    return t_c + t_n, t_n, have_ref


def host_threads():
    """cores this process may run on.  NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1 to every rank, which
    would silently turn the all-cores CPU arm into a 1-thread run at N > 1 (the oracle's *_mt entry points take the thread
    count explicitly)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:  # noqa: BLE001
        return max(1, os.cpu_count() or 1)


def cpu_baseline(sample_faces, repeats):
    from oracle import oracle as O
    V, F, n = cpu_sample_mesh(sample_faces)
    T = O.Topology(F)
    vv, vf = T.query("VV"), T.query("VF")
    t1, t_n1, have_ref = cpu_pass_seconds(V, F, T, vv, vf, repeats, 1)
    nt = host_threads()
    tm, t_nm, _ = cpu_pass_seconds(V, F, T, vv, vf, repeats, nt) if nt > 1 else (t1, t_n1, have_ref)
    serial = ("1 thread as the reference runs it: VV+VF consume over cached CSR adjacency = oracle port; vertex-normal leg = %s "
              "(%.3g faces/s for that leg alone): %.3g faces/s") % (
                  "the reference's own vertex_normal_ref.h compiled unmodified (oracle/_ref)" if have_ref
                  else "oracle port of vertex_normal_ref.h", F.shape[0] / t_n1, F.shape[0] / t1)
    allc = "%d threads, OpenMP ports with per-thread accumulators (oracle rxo_*_mt): %.3g faces/s" % (nt, F.shape[0] / tm)
    best_mt = tm < t1
    return {
        "value": F.shape[0] / min(t1, tm), "unit": "faces/s", "cores": nt if best_mt else 1,
        "kind": "port",
        "sample": "%d x %d grid (%d faces) of the same generator, %d passes; the faster of [%s] and [%s]" %
                  (n, n, F.shape[0], repeats, serial, allc),
        "serial_faces_per_s": F.shape[0] / t1, "all_cores_faces_per_s": F.shape[0] / tm, "host_threads": nt,
    }, (V, F, T, vv, vf)


def reference_gpu_sample(sample_faces, nrun=20, timeout=240):
    """The reference's OWN GPU kernels (oracle/_ref/ref_gpu_queries: rxmesh.cpp + patcher + Query<256>::dispatch
    compiled unmodified for sm_100a, see oracle/ref_gpu_queries.cu) on a bounded sample of the workload, this GPU.
    Reported next to our numbers; never on the product path."""
    import tempfile
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_gpu_queries")
    if not os.path.exists(exe):
        return {"unavailable": "oracle/_ref/ref_gpu_queries not built (needs /root/reference at build time)"}
    V, F, n = cpu_sample_mesh(sample_faces)
    try:
        with tempfile.TemporaryDirectory() as td:
            mesh = os.path.join(td, "mesh.bin")
            with open(mesh, "wb") as fh:
                np.asarray([V.shape[0], F.shape[0]], np.uint32).tofile(fh)
                np.ascontiguousarray(F, np.uint32).tofile(fh)
                np.ascontiguousarray(V, np.float32).tofile(fh)
            r = subprocess.run([exe, mesh, td, "512", "0", str(nrun)], capture_output=True, text=True, timeout=timeout)
            if r.returncode != 0:
                return {"unavailable": "reference binary failed: " + (r.stderr or r.stdout)[-200:]}
            meta = json.load(open(os.path.join(td, "meta.json")))
    except Exception as e:  # noqa: BLE001
        return {"unavailable": str(e)[:200]}
    ms = {"VV": meta["consume"]["VV"], "VF": meta["consume"]["VF"], "VN": meta["vertex_normals"]["ms"]}
    step = sum(ms.values())
    return {"value": F.shape[0] / (step * 1e-3), "unit": "faces/s", "ms": ms, "ms_per_step": step,
            "patches": meta["patches"], "patch_size": 512, "build_seconds": meta["build_ms"] / 1e3,
            "blocks_per_sm": meta["ops"]["VV"]["blocks_per_sm"], "regs": meta["ops"]["VV"]["regs"],
            "sample": ("%d x %d grid (%d faces) of the same generator; the reference's unmodified Query<256>::dispatch "
                       "VV / VF consume lambdas + its FV vertex-normal lambda (global atomics) over raw per-patch arrays, "
                       "its own Lloyd patching at its default patch size; its host build is O(patches x elements), so "
                       "the full 100M-face mesh is out of its reach (111 s at 20M faces); kernel time scales linearly "
                       "with faces (profiles/r01_reference_gpu_vs_ours.jsonl)") % (n, n, F.shape[0])}


def run_reference(args, rank):
    if rank != 0:
        return
    from oracle import oracle as O  # noqa: F401
    sample = min(args.faces, args.cpu_sample_faces)
    V, F, n = cpu_sample_mesh(sample)
    T = O.Topology(F)
    vv, vf = T.query("VV"), T.query("VF")
    nt = host_threads()
    # warm-up also decides the configuration: the reference's serial form, or the all-cores ports when they are faster
    tw = {1: 0.0, nt: 0.0}
    for _ in range(max(1, args.warmup)):
        for k in tw:
            tw[k] = cpu_pass_seconds(V, F, T, vv, vf, 1, k)[0]
    use = min(tw, key=tw.get)
    dt = 0.0
    for _ in range(args.steps):  # timed: the passes only (vector<vector<>> conversion is setup, as in the app)
        t, _, have_ref = cpu_pass_seconds(V, F, T, vv, vf, 1, use)
        dt += t
    dt /= args.steps
    val = F.shape[0] / dt
    how = ("%d OpenMP threads (oracle ports rxo_*_mt; faster here than the reference's serial loop)" % use) if use > 1 else \
          ("1 thread, vertex-normal leg = %s" % ("oracle/_ref (reference's vertex_normal_ref.h, unmodified)" if have_ref
                                                 else "oracle port"))
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "faces/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(config_dict(args, grid_side(args.faces), 2 * (grid_side(args.faces) - 1) ** 2, None),
                       sample="%d x %d grid, %d faces per step" % (n, n, F.shape[0])),
        "cpu_baseline": {"value": val, "unit": "faces/s", "cores": use, "kind": "port" if use > 1 or not have_ref else "reference",
                         "sample": ("each step = one pass over a %d x %d grid (%d faces), a bounded sample of the "
                                    "workload; %s; VV/VF consume legs = oracle port over cached CSR (the reference has no "
                                    "CPU query engine); host threads available: %d") % (n, n, F.shape[0], how, nt)},
        "e2e": {"value": val, "unit": "faces/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


METRIC = "faces/s through one VV-query + VF-query + vertex-normal pass (100M-face grid); per-kernel entries/s, faces/s and HBM fraction in `kernels`"


def headline_oracle_check(n, mesh, rx, sv_in, sv_out, sf_in, nrm, in_v, in_f, stream, torch):
    """What the timed kernels left in their output attributes (vertex normals, VF consume) plus one more VV consume, against
    the oracle on 81 x 81-vertex windows of the n x n grid (corners, edges, centre); vertices on a window side that is a cut
    through the mesh, not a mesh border, lack neighbours inside the window and are left out."""
    from bench_configs import grid_window
    from oracle import oracle as O
    got_vn = nrm.to_global()
    got_vf = sv_out.to_global().reshape(-1)  # the step's last query
    mesh.query_consume(rx.Op.VV, sv_in, sv_out, stream)
    torch.cuda.synchronize()
    got_vv = sv_out.to_global().reshape(-1)
    worst = {"VV": 0.0, "VF": 0.0, "VN": 0.0}
    cnt = 0
    for r0 in (0, n // 2 - 40, n - 81):
        for c0 in (0, n // 2 - 40, n - 81):
            r1, c1 = r0 + 80, c0 + 80
            Vw, Fw, gid = grid_window(n, r0, r1, c0, c1)
            h, w = r1 - r0 + 1, c1 - c0 + 1
            ok = np.ones((h, w), bool)
            if r0 > 0:
                ok[0, :] = False
            if r1 < n - 1:
                ok[-1, :] = False
            if c0 > 0:
                ok[:, 0] = False
            if c1 < n - 1:
                ok[:, -1] = False
            sel = ok.reshape(-1)
            T = O.Topology(Fw)
            lq = np.arange((h - 1) * (w - 1), dtype=np.int64)
            gq = (r0 + lq // (w - 1)) * (n - 1) + c0 + lq % (w - 1)
            gf = np.stack([2 * gq, 2 * gq + 1], 1).reshape(-1)  # global ids of the window's faces, in its face order
            for name, op, src, got in (("VV", "VV", in_v[gid].astype(np.float64), got_vv), ("VF", "VF", in_f[gf].astype(np.float64), got_vf)):
                off, val = T.query(op)
                ref = np.add.reduceat(np.concatenate([src[val], [0.0]]), np.minimum(off[:-1], val.shape[0]))
                ref[np.diff(off) == 0] = 0.0
                d = np.abs(got[gid][sel] - ref[sel]) / np.maximum(np.abs(ref[sel]), 1e-30)
                worst[name] = max(worst[name], float(d.max()))
            refn = O.vertex_normals(Fw, Vw, np.float64)
            d = np.linalg.norm(got_vn[gid][sel] - refn[sel], axis=1) / np.maximum(np.linalg.norm(refn[sel], axis=1), 1e-30)
            worst["VN"] = max(worst["VN"], float(d.max()))
            cnt += int(sel.sum())
    tol = {"VV": 1e-6, "VF": 1e-6, "VN": 1e-5}
    return {"what": "results of the timed kernels on the full mesh against the oracle, nine 81 x 81-vertex windows (corners, "
                    "edges, centre), relative error", "vertices_checked": cnt, "max_rel_err": worst, "tolerance_rel": tol,
            "ok": bool(all(worst[k] <= tol[k] for k in tol))}


def config_dict(args, n, faces, mesh):
    c = {"workload": "VV + VF query (consume) + vertex normals on a %d x %d procedurally generated grid "
                     "(create_plane semantics + height field), %d faces per GPU" % (n, n, faces),
         "faces_per_gpu": int(faces), "patch_size": 2 * TILE * TILE_I, "patcher": "analytic %dx%d-quad tiles" % (TILE, TILE_I),
         "l2": "inputs larger than L2 (topology + attributes stream > 2 GB per step; no explicit flush)",
         "parallelism": "patches sharded per GPU" if args.gpus > 1 else "1 GPU"}
    if mesh is not None:
        c.update(patches=mesh.get_num_patches(), ribbon_overhead=round(mesh.ribbon_overhead(), 4),
                 topo_bytes_per_face=round(mesh.topo_bytes() / faces, 2))
    return c


# ------------------------------------------------------------------------------------ GPU arm
def run_ours(args, rank, world, local_rank):
    import torch

    import rxmesh_b200 as rx
    from rxmesh_b200 import meshio

    torch.cuda.set_device(local_rank)
    rx.rx_init(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n = grid_side(args.faces)
    ncores = host_threads()
    t0 = time.perf_counter()
    hx_v = hx_f = None
    if world == 1:
        V, F = meshio.grid(n, n)
        fp = meshio.grid_face_tiles(n, n, TILE, TILE_I)
        # ring2=False: the ring extension serves k-ring consumers (bilateral filtering) only; this mesh never runs one
        mesh = rx.RXMeshStatic(F, face_patch=fp, patch_size=2 * TILE * TILE_I, num_threads=ncores, ring2=False)
        del fp
        nF, nV = mesh.get_num_faces(), mesh.get_num_vertices()
        nE_real = mesh.get_num_edges()
        halo_bytes = 0
    else:
        # weak scaling: a (world * (n-1) + 1)-row grid, every rank owns a slab of ~args.faces faces and
        # mirrors one ghost tile row per neighbour (rxmesh_b200/distributed.py)
        from rxmesh_b200 import distributed as D
        sh = D.grid_slab(n, world * (n - 1) + 1, TILE, TILE_I, rank, world)
        sm = D.ShardedMesh(sh, rank, world, patch_size=2 * TILE * TILE_I, num_threads=max(1, ncores // world), ring2=False)
        mesh, V = sm.mesh, sh["verts"]
        hx_v, hx_f = D.HaloExchange(sm, 0), D.HaloExchange(sm, 2)
        _ = mesh.ribbon_overhead()
        lbf, lbv, lbe = mesh.lin_base(2), mesh.lin_base(0), mesh.lin_base(1)
        a, b = sm.first, sm.first + sm.count
        nF, nV = int(lbf[b] - lbf[a]), mesh.get_num_vertices()  # real faces; host arrays cover the whole slab
        nE_real = int(lbe[b] - lbe[a])
        halo_bytes = 12 * hx_v.halo_elements()
    t_build = time.perf_counter() - t0
    stream = torch.cuda.current_stream()
    x = rx.Attribute(mesh, 0, np.float32, 3, rx.DEVICE, rx.AoS)
    nrm = rx.Attribute(mesh, 0, np.float32, 3, rx.DEVICE, rx.AoS)
    fused = None
    if hx_v is not None and args.workload == "laplacian" and not os.environ.get("RXM_NO_FUSED"):
        from rxmesh_b200 import distributed as D
        try:
            fused = D.FusedHalo(hx_v, x, nrm)  # needs the host patch store: before compact()
        except Exception as e:  # noqa: BLE001  (e.g. no peer access between the devices): packed NCCL exchange instead
            print("rank %d: fused halo unavailable (%s), using NCCL send/recv" % (rank, str(e)[:200]), file=sys.stderr)
            fused = None
        # every rank must take the same path
        ok = torch.tensor([1 if fused is not None else 0], device="cuda")
        torch.distributed.all_reduce(ok, op=torch.distributed.ReduceOp.MIN)
        if int(ok[0]) == 0:
            fused = None
    mesh.compact()  # the 100 M-face mesh lives on the device; drop host helper arrays (halo plans are built)
    sv_in = rx.Attribute(mesh, 0, np.float32, 1, rx.DEVICE, rx.AoS)
    sv_out = rx.Attribute(mesh, 0, np.float32, 1, rx.DEVICE, rx.AoS)
    sf_in = rx.Attribute(mesh, 2, np.float32, 1, rx.DEVICE, rx.AoS)
    rng = np.random.RandomState(1 + rank)
    h_x = torch.from_numpy(V).pin_memory()
    h_sv = torch.from_numpy(rng.rand(nV).astype(np.float32)).pin_memory()
    h_sf = torch.from_numpy(rng.rand(mesh.get_num_faces()).astype(np.float32)).pin_memory()
    x.from_global(h_x.numpy(), stream)
    sv_in.from_global(h_sv.numpy(), stream)
    sf_in.from_global(h_sf.numpy(), stream)
    torch.cuda.synchronize()
    if hx_v is not None:  # static inputs: mirrored once; the coordinates are re-mirrored every step
        hx_v.exchange(sv_in, stream)
        hx_f.exchange(sf_in, stream)
        hx_v.exchange(x, stream)
        torch.cuda.synchronize()

    if args.workload == "laplacian":
        return run_laplacian(args, rank, world, local_rank, mesh, x, nrm, hx_v, nV if world == 1 else None,
                             n, t_build, torch, rx, stream, fused)

    # high priority: the exchange's small kernels (gather, NCCL send/recv, scatter) are scheduled as soon as a block slot frees
    # up instead of behind the ~100 000 blocks the query kernel of the same step still has to dispatch
    comm = torch.cuda.Stream(priority=-1) if hx_v is not None else None

    def step(evs=None):
        if evs:
            evs[0].record(stream)
        if hx_v is not None:
            # ribbon (halo) exchange of the coordinates, inside the timed step, on its own stream: only the vertex
            # normals read the mirrored coordinates, so the NVLink transfer overlaps the two query kernels
            comm.wait_stream(stream)
            with torch.cuda.stream(comm):
                hx_v.exchange(x, comm)
        mesh.query_consume(rx.Op.VV, sv_in, sv_out, stream)
        if evs:
            evs[1].record(stream)
        mesh.query_consume(rx.Op.VF, sf_in, sv_out, stream)
        if evs:
            evs[2].record(stream)
        if hx_v is not None:
            stream.wait_stream(comm)
        mesh.vertex_normals(x, nrm, False, stream)
        if evs:
            evs[3].record(stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    t_w = time.perf_counter()
    n_warm = 0
    while True:  # >= W steps and long enough for the clocks to settle / the sampler to start
        for _ in range(16):
            step()
        n_warm += 16
        torch.cuda.synchronize()
        more = n_warm < args.warmup or time.perf_counter() - t_w < 0.6
        if world > 1:
            # every step holds a send/recv pair with the neighbours: all ranks must run the SAME number of steps, so the
            # time-based decision is taken collectively (a rank-local clock would leave unmatched exchanges behind)
            flag = torch.tensor([1 if more else 0], device="cuda")
            torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MAX)
            more = bool(int(flag[0]))
        if not more:
            break
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    barrier()
    l0 = rx.launch_count()
    t_begin = time.perf_counter()
    for k in range(args.steps):
        step(evs[k])
    barrier()
    t_end = time.perf_counter()
    t_wall = t_end - t_begin
    launches = rx.launch_count() - l0
    clocks = sampler.stop(t_begin, t_end)
    total_ms = evs[0][0].elapsed_time(evs[-1][3])
    k_ms = [sum(e[i].elapsed_time(e[i + 1]) for e in evs) / args.steps for i in range(3)]

    # ---- the timed kernels' results against the oracle, at full size, on windows of the grid (N = 1) ----
    parity = None
    if world == 1 and not args.no_cpu:
        parity = headline_oracle_check(n, mesh, rx, sv_in, sv_out, sf_in, nrm, h_sv.numpy(), h_sf.numpy(), stream, torch)

    # ---- end to end through the C ABI with pinned HOST buffers (H2D + kernel + D2H per call) ----
    h_n = torch.empty((nV, 3), dtype=torch.float32).pin_memory()
    h_o1 = torch.empty(nV, dtype=torch.float32).pin_memory()
    h_o2 = torch.empty(nV, dtype=torch.float32).pin_memory()

    # three host calls on three streams (rxm_set_async): their PCIe copies (H2D and D2H are full duplex) and
    # kernels overlap; the step ends when all three results are back in the pinned host buffers
    s3 = [torch.cuda.Stream() for _ in range(3)]
    rx.set_async(True)

    def e2e_step():
        # largest transfer first: its D2H (normals, 12 B/vertex) then overlaps the H2D of the two query inputs
        mesh.vertex_normals_host_ptr(h_x.data_ptr(), h_n.data_ptr(), s3[2])
        mesh.query_consume_host_ptr(rx.Op.VF, h_sf.data_ptr(), h_o2.data_ptr(), s3[1])
        mesh.query_consume_host_ptr(rx.Op.VV, h_sv.data_ptr(), h_o1.data_ptr(), s3[0])
        for st in s3:
            st.synchronize()

    e2e_steps = max(1, min(args.steps, 5))
    e2e_step()
    barrier()
    te = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    te = (time.perf_counter() - te) / e2e_steps
    rx.set_async(False)
    # what the box's host <-> device path allows for exactly these bytes: plain pinned copies, H2D on one stream and D2H on
    # another (full duplex), every rank at the same time, no kernel.  e2e can approach this floor, not beat it; at N > 1 the
    # ranks share the host's memory and PCIe root complexes, so the floor itself grows with N
    d_in = [torch.empty_like(t, device="cuda") for t in (h_x, h_sv, h_sf)]
    d_out = [torch.empty_like(t, device="cuda") for t in (h_n, h_o1, h_o2)]

    def copy_only():
        with torch.cuda.stream(s3[0]):
            for d_, h_ in zip(d_in, (h_x, h_sv, h_sf)):
                d_.copy_(h_, non_blocking=True)
        with torch.cuda.stream(s3[1]):
            for d_, h_ in zip(d_out, (h_n, h_o1, h_o2)):
                h_.copy_(d_, non_blocking=True)
        s3[0].synchronize(), s3[1].synchronize()

    keep = [t.clone() for t in (h_n[:1000], h_o1[:1000], h_o2[:1000])]
    for d_, h_ in zip(d_out, (h_n, h_o1, h_o2)):
        d_.copy_(h_)
    torch.cuda.synchronize()
    copy_only()
    barrier()
    tc = time.perf_counter()
    for _ in range(3):
        copy_only()
    barrier()
    tc = (time.perf_counter() - tc) / 3
    assert all(torch.equal(k_, h_[:1000]) for k_, h_ in zip(keep, (h_n, h_o1, h_o2)))
    del d_in, d_out
    # the overlapped, pipelined calls must give exactly what a synchronous call gives
    assert np.array_equal(mesh.vertex_normals_host(h_x.numpy())[:1000], h_n.numpy()[:1000])
    h2d = 12 * nV + 4 * nV + 4 * mesh.get_num_faces()
    d2h = 12 * nV + 4 * nV + 4 * nV

    # ---- max over ranks ----
    tm = torch.tensor([total_ms, te] + k_ms + [tc], dtype=torch.float64, device="cuda")
    if world > 1:
        torch.distributed.all_reduce(tm, op=torch.distributed.ReduceOp.MAX)
    total_ms, te = float(tm[0]), float(tm[1])
    k_ms = [float(v) for v in tm[2:5]]
    tc = float(tm[5])
    if rank != 0:
        return None
    ms_step = total_ms / args.steps
    peak, peak_src = peaks()
    kern = {}
    for name, ms, unit, units in (("VV", k_ms[0], "neighbour entries/s", 2.0 * nE_real),
                                  ("VF", k_ms[1], "neighbour entries/s", 3.0 * nF),
                                  ("VN", k_ms[2], "faces/s", float(nF))):
        gbs = ALG_BYTES_PER_FACE[name] * nF / (ms * 1e-3) / 1e9
        kern[name] = {"ms": ms, "rate": units / (ms * 1e-3) * world, "unit": unit,
                      "alg_bytes": ALG_BYTES_PER_FACE[name] * nF, "achieved_gbs": gbs, "hbm_frac": gbs / peak}
    dom = max(kern, key=lambda k: kern[k]["ms"])
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            traffic = tj.get(dom)
            # fraction by ACTUAL DRAM bytes (ncu, per launch, 100M-face mesh): the algorithmic-byte fraction can exceed 1
            # because 16-bit local ids and owned-only attributes move fewer bytes than the canonical count
            for k in kern:
                if k in tj and world == 1 and abs(nF - 99998082) < 1000:
                    kern[k]["dram_bytes_ncu"] = tj[k]
                    kern[k]["dram_frac"] = tj[k] / (kern[k]["ms"] * 1e-3) / 1e9 / peak
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": nF * world / (ms_step * 1e-3), "unit": "faces/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "warmup_steps_run": n_warm, "ms_per_step": ms_step,
        "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, n, nF, mesh),
        "kernels": kern,
        "roofline": {"bound": "hbm", "kernel": {"VV": "k_vv_consume_fan<128>", "VF": "k_vf_consume_fan<128>",
                                                "VN": "k_vertex_normals_fan2<0>"}[dom],
                     "achieved": kern[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                     "frac": kern[dom]["hbm_frac"], "traffic": traffic, "peak_source": peak_src,
                     "alg_bytes_per_launch": kern[dom]["alg_bytes"]},
        "e2e": {"value": nF * world / te, "unit": "faces/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": te * 1e3,
                "copy_only_ms_per_step": tc * 1e3, "frac_of_copy_only": tc / te,
                "copy_only": "the same H2D and D2H bytes as plain pinned copies on two streams, all ranks at once, no kernel: "
                             "the floor this box's host memory / PCIe path sets for the step (%.1f GB/s aggregate)"
                             % ((h2d + d2h) * world / tc / 1e9),
                "api": "rxm_vertex_normals_host, rxm_query_consume_host(VF), rxm_query_consume_host(VV) on 3 streams (rxm_set_async), "
                       "pinned host buffers; each call is a chunked H2D / kernel / D2H pipeline"},
        "gpu_launches": int(launches), "clocks": clocks, "halo_bytes_per_step_per_gpu": int(halo_bytes),
        "wall_ms_per_step": t_wall / args.steps * 1e3, "build_seconds": t_build,
        "host_peak_rss_gb": __import__("resource").getrusage(__import__("resource").RUSAGE_SELF).ru_maxrss / 1048576.0,
    }
    if parity is not None:
        line["parity"] = parity
    if world == 1 and not args.no_cpu:
        cb, _ = cpu_baseline(min(args.faces, args.cpu_sample_faces), 3)
        line["cpu_baseline"] = cb
        line["reference_gpu"] = reference_gpu_sample(min(args.faces, args.cpu_sample_faces))
    return line


def run_laplacian(args, rank, world, local_rank, mesh, x, y, hx_v, nv_single, n, t_build, torch, rx, stream, fused=None):
    """BASELINE.json configs[4]: iterated Laplacian smoothing (apps/Smoothing/manual.h:86-104), patches
    sharded across the ranks, ribbon exchange after every iteration inside the timed region."""
    lbv = mesh.lin_base(0)
    if hx_v is None:
        n_upd = mesh.get_num_vertices()
    else:
        n_upd = None
    # Exchange after every iteration on the compute stream.  Splitting the step into boundary patches -> exchange on a
    # second stream -> interior patches was measured at N = 2 (0.266 vs 0.262 ms per iteration): the 60 KB transfer is
    # already cheap, the extra launches cost more than the overlap returns, so the simple order stays.
    # Default at N > 1: ONE kernel per iteration that also pushes the mirrored rows into the neighbours' ghost slots over
    # NVLink and synchronises with them through flag words (rxmesh_b200/distributed.py: FusedHalo).  RXM_NO_FUSED=1:
    # kernel, then packed NCCL send/recv.
    def iterate(k):
        if fused is not None:
            fused.smooth(0.01, k, stream)
            return
        a, b = x, y
        for _ in range(k):
            mesh.laplacian_smooth(a, b, 0.01, 1, stream)
            if hx_v is not None:
                hx_v.exchange(b, stream)
            a, b = b, a

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
            torch.cuda.synchronize()
    iterate(max(3, args.warmup))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = rx.launch_count()
    e0.record(stream)
    for _ in range(args.steps):
        iterate(args.iters)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    tm = torch.tensor([ms], dtype=torch.float64, device="cuda")
    real_v = torch.tensor([float(n_upd) if n_upd is not None else float((mesh.elem_patch(0) >= 0).sum())],
                          dtype=torch.float64, device="cuda")
    if world > 1:
        from rxmesh_b200 import distributed as D  # noqa: F401
        torch.distributed.all_reduce(tm, op=torch.distributed.ReduceOp.MAX)
    if rank != 0:
        return
    ms = float(tm[0])
    its = args.steps * args.iters
    nF = 2 * (n - 1) ** 2 * world
    peak, peak_src = peaks()
    per_it = ms / its
    print(json.dumps({
        "metric": "vertex-updates/s, iterated Laplacian smoothing, patches sharded across GPUs, ribbon exchange in the timing",
        "value": (nF / 2.0) / (per_it * 1e-3), "unit": "vertex-updates/s", "n_gpus": world, "steps": args.steps,
        "iters_per_step": args.iters, "ms_per_iteration": per_it, "higher_is_better": True, "scaling": "weak",
        "dtype": "f32", "data": "synthetic", "faces_total": nF,
        "roofline": {"bound": "hbm", "kernel": "k_laplacian_fan2", "achieved": 24.0 * nF / world / (per_it * 1e-3) / 1e9,
                     "peak": peak, "unit": "GB/s", "frac": 24.0 * nF / world / (per_it * 1e-3) / 1e9 / peak,
                     "peak_source": peak_src, "note": "per GPU, includes the halo exchange time"},
        "halo_bytes_per_iteration_per_gpu": 0 if hx_v is None else 12 * hx_v.halo_elements(),
        "halo_transport": None if hx_v is None else ("fused into the compute kernel: NVLink P2P stores + flag words" if fused is not None
                                                      else "packed NCCL send/recv after the kernel"),
        "gpu_launches": int(rx.launch_count() - l0), "build_seconds": t_build}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--faces", type=int, default=100_000_000)
    ap.add_argument("--cpu-sample-faces", type=int, default=4_000_000)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--workload", default="queries", choices=["queries", "laplacian"],
                    help="queries = the headline VV+VF+normals pass; laplacian = iterated smoothing (configs[4])")
    ap.add_argument("--iters", type=int, default=100, help="Laplacian iterations per step")
    ap.add_argument("--sub", default="all",
                    help="sub-records next to the headline: all | none | comma list of dragon,queries,bilateral,hardwired,laplacian")
    ap.add_argument("--lap-faces", type=int, default=0, help="faces of the strong-scaled Laplacian mesh (default: 400M, configs[4])")
    ap.add_argument("--sub-timeout", type=int, default=600, help="seconds after which the sub-records are abandoned")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    import gc
    line, done = None, threading.Event()

    def emit():
        if rank == 0 and line is not None and not done.is_set():
            done.set()
            print(json.dumps(line), flush=True)

    def watchdog():
        # a hung sub-record (one rank failing inside a collective) must not cost the headline: print what we have and leave
        if rank == 0 and line is not None:
            line.setdefault("sub_records_error", "watchdog: sub-records exceeded %d s" % args.sub_timeout)
        emit()
        os._exit(0 if line is not None or rank != 0 else 1)

    try:
        if args.workload == "laplacian":
            run_ours(args, rank, world, local_rank)
            return
        line = run_ours(args, rank, world, local_rank)
        gc.collect()
        import torch
        torch.cuda.empty_cache()
        subs = set() if args.sub == "none" else set(args.sub.split(","))
        timer = threading.Timer(args.sub_timeout, watchdog)
        timer.daemon = True
        timer.start()
        try:
            run_sub_records(args, rank, world, local_rank, line, subs)
        finally:
            timer.cancel()
        emit()
    finally:
        if world > 1:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.destroy_process_group()


def run_sub_records(args, rank, world, local_rank, line, subs):
    """BASELINE.json configs[0..2] (1 GPU only) and configs[4] (every N) next to the headline, each with an in-run parity
    flag; bench_configs.py holds the workloads.  A failing sub-record reports its error and leaves the rest alone."""
    import traceback

    import torch

    import bench_configs as BC
    import rxmesh_b200 as rx
    stream = torch.cuda.current_stream()
    want = lambda k: "all" in subs or k in subs  # noqa: E731

    def guarded(fn):
        t0 = time.perf_counter()
        try:
            r = fn()
        except Exception as e:  # noqa: BLE001
            r = {"error": (type(e).__name__ + ": " + str(e))[:300], "trace": traceback.format_exc()[-600:]}
        if isinstance(r, dict):
            r["wall_seconds"] = round(time.perf_counter() - t0, 2)
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        return r

    if world == 1:
        cfg = {}
        if want("dragon"):
            cfg["1_vertex_normal_dragon"] = guarded(lambda: BC.config_dragon(torch, rx, stream))
        if want("queries"):
            cfg["2_eight_queries_10m_sphere"] = guarded(lambda: BC.config_queries(torch, rx, stream))
        if want("bilateral"):
            cfg["3_bilateral_10m_torus"] = guarded(lambda: BC.config_bilateral(torch, rx, stream))
            # the MCF solve on the same mesh (SURVEY.md 8(f)1 taken to its caller) as a record of its own
            c3 = cfg["3_bilateral_10m_torus"]
            if isinstance(c3, dict) and isinstance(c3.get("mcf_cg_same_mesh"), dict):
                cfg["f1_mcf_solve_10m_torus"] = c3.pop("mcf_cg_same_mesh")
        if cfg and line is not None:
            line["configs"] = cfg
        if want("hardwired") and line is not None and isinstance(line.get("reference_gpu"), dict):
            hw = guarded(lambda: BC.hardwired_baseline(grid_side(args.faces)))
            line["reference_gpu"]["hardwired"] = hw
            if "hardwired_ms" in hw:
                line["reference_gpu"]["hardwired_ms"] = hw["hardwired_ms"]
                line["reference_gpu"]["ours_vn_ms"] = line["kernels"]["VN"]["ms"]
    if want("laplacian"):
        rec = guarded(lambda: BC.laplacian_400m(args, rank, world, local_rank, torch, rx, TILE, TILE_I))
        if line is not None:
            line["laplacian_400m"] = rec


if __name__ == "__main__":
    main()
