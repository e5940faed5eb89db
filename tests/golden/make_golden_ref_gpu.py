"""Generate tests/golden/ref_gpu_<mesh>.npz from the REFERENCE'S OWN GPU implementation.

Runs ON THE B200 BOX (the reference's kernels need a GPU; the build container has none):

  gpurun -- 'python tests/golden/make_golden_ref_gpu.py'     # writes gpurun_out/ref_gpu/*.npz + timings.jsonl
  cp gpurun_out/ref_gpu/ref_gpu_*.npz tests/golden/          # then commit

oracle/_ref/ref_gpu_queries is the reference's unmodified rxmesh.cpp / patcher / LP hash table / patch stash /
Query<256>::dispatch compiled from /root/reference by `make -C oracle ref_gpu` in the build container (see
oracle/ref_gpu_queries.cu); it travels to the box with the snapshot.  For every fixture mesh this records what the
reference itself produced on this GPU: its patching (face -> patch), per-patch local->global maps and owned counts,
the eight query results as global ids in the reference's iteration order, its vertex normals, and the positions after 1 and
5 iterations of its manual smoothing (apps/Smoothing/manual.h:86-104, lap_1 / lap_5).
tests/test_gpu_ref_golden.py replays the patching through our builder and compares handle for handle.

With --timing it also times the reference's kernels on larger procedural meshes next to ours (same mesh, same GPU)
and appends the lines to gpurun_out/ref_gpu/timings.jsonl (copied to profiles/ by hand).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
BIN = os.path.join(ROOT, "oracle", "_ref", "ref_gpu_queries")
OUT = os.path.join(ROOT, "gpurun_out", "ref_gpu")
OPS = ["VV", "VE", "VF", "EV", "EF", "FV", "FE", "FF"]


def run_ref(V, F, patch_size, dump, nrun, timeout=1500):
    """-> (meta dict, dict of arrays) from one run of the reference binary."""
    with tempfile.TemporaryDirectory() as td:
        mesh = os.path.join(td, "mesh.bin")
        with open(mesh, "wb") as fh:
            np.asarray([V.shape[0], F.shape[0]], np.uint32).tofile(fh)
            np.ascontiguousarray(F, np.uint32).tofile(fh)
            np.ascontiguousarray(V, np.float32).tofile(fh)
        t0 = time.perf_counter()
        r = subprocess.run([BIN, mesh, td, str(patch_size), "1" if dump else "0", str(nrun)], capture_output=True,
                           text=True, timeout=timeout)
        if r.returncode != 0:
            raise RuntimeError("reference binary failed (%d): %s %s" % (r.returncode, r.stdout[-2000:], r.stderr[-2000:]))
        meta = json.load(open(os.path.join(td, "meta.json")))
        meta["wall_s"] = time.perf_counter() - t0
        arrs = {}
        if dump:
            for fn in os.listdir(td):
                if fn.endswith(".u32"):
                    arrs[fn[:-4]] = np.fromfile(os.path.join(td, fn), dtype=np.uint32)
                elif fn.endswith(".f32"):
                    arrs[fn[:-4]] = np.fromfile(os.path.join(td, fn), dtype=np.float32)
            for op in OPS + ["EVDiamond", "EE"]:
                if "q_" + op in arrs:
                    arrs["q_" + op] = arrs["q_" + op].reshape(-1, meta["ops"][op]["width"])
            arrs["vn"] = arrs["vn"].reshape(-1, 3)
            for k in ("lap_1", "lap_5"):
                if k in arrs:
                    arrs[k] = arrs[k].reshape(-1, 3)
        return meta, arrs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--meshes", default="sphere3,dragon,bunnyhead,torus,cube,plane_5")
    ap.add_argument("--timing", default="", help="comma list of grid sizes n (n x n vertices) to time, e.g. 708,1415")
    ap.add_argument("--patch-size", type=int, default=512)
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    from conftest import make_mesh

    for name in [m for m in args.meshes.split(",") if m]:
        V, F = make_mesh(name)
        meta, arrs = run_ref(V, F, args.patch_size, True, 10)
        arrs["meta_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        # widths > what is needed waste space: keep as is (0xFFFFFFFF compresses well)
        np.savez_compressed(os.path.join(OUT, "ref_gpu_%s.npz" % name), **arrs)
        print(name, {k: meta[k] for k in ("nv", "ne", "nf", "patches")}, "ops ms", {o: round(meta["ops"][o]["ms"], 4) for o in OPS},
              flush=True)

    if args.timing:
        import torch

        import rxmesh_b200 as rx
        from rxmesh_b200 import meshio
        from rxmesh_b200.mesh import _SRC
        sys.path.insert(0, ROOT)
        from bench_configs import timed
        rx.rx_init(0)
        stream = torch.cuda.current_stream()
        for n in [int(s) for s in args.timing.split(",")]:
            V, F = meshio.grid(n, n)
            line = {"mesh": "grid %d x %d (create_plane semantics + height field)" % (n, n), "faces": int(F.shape[0])}
            try:
                meta, arrs = run_ref(V, F, args.patch_size, True, 100)
            except Exception as e:  # the reference's O(P (V+E)) host build does not scale; record and go on
                line["reference"] = {"failed": str(e)[:300]}
                print(json.dumps(line), flush=True)
                continue
            line["reference"] = {"what": "reference Query<256>::dispatch store kernels + FV normals lambda, recompiled "
                                         "unmodified for sm_100a, patch_size %d, its own GPU Lloyd patching" % args.patch_size,
                                 "patches": meta["patches"], "build_s": meta["build_ms"] / 1e3,
                                 "ms": {o: meta["ops"][o]["ms"] for o in OPS}, "vertex_normals_ms": meta["vertex_normals"]["ms"],
                                 "blocks_per_sm": meta["ops"]["VV"]["blocks_per_sm"]}
            ours = {}
            # (a) the reference's own patching replayed; (b) our default Lloyd patches of <= 1024 faces
            for tag, kw in (("same_patching", dict(face_patch=arrs["face_patch"], patch_size=args.patch_size)),
                            ("lloyd_1024", dict(patch_size=1024))):
                t0 = time.perf_counter()
                m = rx.RXMeshStatic(F, **kw)
                tb = time.perf_counter() - t0
                res = {}
                for op in OPS:
                    o = rx.Op[op]
                    width = meta["ops"][op]["width"]
                    inp = rx.Attribute(m, _SRC[o], np.uint64, 1, rx.DEVICE, rx.AoSoA)
                    out = rx.Attribute(m, _SRC[o], np.uint64, width, rx.DEVICE, rx.AoSoA)
                    res[op] = timed(lambda: m.query_store(o, inp, out, stream), stream, torch, 100)
                    inp.release(), out.release()
                x = rx.Attribute(m, 0, np.float32, 3, rx.DEVICE, rx.AoS)
                nrm = rx.Attribute(m, 0, np.float32, 3, rx.DEVICE, rx.AoS)
                x.from_global(V)
                vn = timed(lambda: m.vertex_normals(x, nrm, False, stream), stream, torch, 100)
                ours[tag] = {"patches": m.get_num_patches(), "build_s": tb, "ms": res, "vertex_normals_ms": vn}
                del m, x, nrm
            line["ours"] = ours
            # the reference app's own kernel source (Query::dispatch<Op::FV> lambda with global atomics,
            # tests/cpp/shim_apps.cu:user_vertex_normal) compiled against OUR drop-in headers: what a user gets
            # without touching their code
            import ctypes as C
            shim = C.CDLL(os.path.join(ROOT, "tests", "cpp", "libshim_apps.so"))
            msf = C.c_float()
            fpr = np.ascontiguousarray(arrs["face_patch"], np.uint32)
            if shim.shim_time_vertex_normals(F.ctypes.data_as(C.c_void_p), F.shape[0], V.ctypes.data_as(C.c_void_p), V.shape[0],
                                             fpr.ctypes.data_as(C.c_void_p), args.patch_size, 100, C.byref(msf)) == 0:
                line["user_kernel_on_our_headers_vertex_normals_ms"] = msf.value
                line["speedup_user_kernel_unchanged"] = meta["vertex_normals"]["ms"] / msf.value
            line["speedup_same_patching"] = {o: meta["ops"][o]["ms"] / ours["same_patching"]["ms"][o] for o in OPS}
            line["speedup_same_patching"]["vertex_normals"] = meta["vertex_normals"]["ms"] / ours["same_patching"]["vertex_normals_ms"]
            line["speedup_lloyd_1024"] = {o: meta["ops"][o]["ms"] / ours["lloyd_1024"]["ms"][o] for o in OPS}
            line["speedup_lloyd_1024"]["vertex_normals"] = meta["vertex_normals"]["ms"] / ours["lloyd_1024"]["vertex_normals_ms"]
            with open(os.path.join(OUT, "timings.jsonl"), "a") as fh:
                fh.write(json.dumps(line) + "\n")
            print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
