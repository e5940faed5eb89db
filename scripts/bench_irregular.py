"""Consume + normals + Laplacian on a mesh of MIXED valence (4..8): an n x n grid whose quads are split along a random
diagonal (meshio.grid_random_diagonals), Lloyd patches -- most vertices leave the valence-6 fast paths.
python scripts/bench_irregular.py [n] [patch_size]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import rxmesh_b200 as rx  # noqa: E402
from oracle import oracle as O  # noqa: E402
from rxmesh_b200 import meshio  # noqa: E402


def timed(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


n = int(sys.argv[1]) if len(sys.argv) > 1 else 2237
ps = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
rx.rx_init(0)
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
peak = json.load(open(pk)).get("hbm_gbs", 6539.9) if os.path.exists(pk) else 6539.9
V, F = meshio.grid_random_diagonals(n)
t = time.time()
m = rx.RXMeshStatic(F, patch_size=ps)
tb = time.time() - t
nF, nV = F.shape[0], V.shape[0]
val = np.bincount(F.reshape(-1), minlength=nV)
x = rx.Attribute(m, 0, np.float32, 3, rx.DEVICE, rx.AoS)
y = rx.Attribute(m, 0, np.float32, 3, rx.DEVICE, rx.AoS)
sv, so = rx.Attribute(m, 0, np.float32, 1, rx.DEVICE, rx.AoS), rx.Attribute(m, 0, np.float32, 1, rx.DEVICE, rx.AoS)
sf = rx.Attribute(m, 2, np.float32, 1, rx.DEVICE, rx.AoS)
x.from_global(V)
rng = np.random.RandomState(3)
hv, hf = rng.rand(nV).astype(np.float32), rng.rand(nF).astype(np.float32)
sv.from_global(hv), sf.from_global(hf)
rec = {"mesh": "%d x %d grid, random diagonals, %d faces, Lloyd patches <= %d" % (n, n, nF, ps), "patches": m.get_num_patches(),
       "build_s": tb, "has_fans": bool(m.has_fans()),
       "faces_per_vertex_histogram": {int(k): int(c) for k, c in enumerate(np.bincount(val)) if c}}
for key, fn, bpf in (("VV", lambda: m.query_consume(rx.Op.VV, sv, so), 16.0), ("VF", lambda: m.query_consume(rx.Op.VF, sf, so), 18.0),
                     ("VN", lambda: m.vertex_normals(x, y), 24.0), ("LAP", lambda: m.laplacian_smooth(x, y, 0.01, 1), 24.0)):
    ms = timed(fn)
    rec[key] = {"ms": ms, "hbm_frac": bpf * nF / (ms * 1e-3) / 1e9 / peak}
if nF <= 12_000_000:  # parity on the whole mesh
    T = O.Topology(F)
    m.vertex_normals(x, y)
    ref = O.vertex_normals(F, V, np.float64)
    got = y.to_global()
    rec["VN"]["max_rel_err"] = float((np.linalg.norm(got - ref, axis=1) / np.maximum(np.linalg.norm(ref, axis=1), 1e-30)).max())
    m.query_consume(rx.Op.VV, sv, so)
    off, vals = T.query("VV")
    refv = np.add.reduceat(np.concatenate([hv.astype(np.float64)[vals], [0.0]]), np.minimum(off[:-1], vals.shape[0]))
    rec["VV"]["max_rel_err"] = float((np.abs(so.to_global().reshape(-1) - refv) / np.maximum(np.abs(refv), 1e-30)).max())
    m.laplacian_smooth(x, y, 0.01, 1)
    refl = O.laplacian_step(T.query("VV"), V.astype(np.float64), 0.01, np.float64)
    rec["LAP"]["max_abs_err"] = float(np.abs(y.to_global() - refl).max())
print(json.dumps(rec))
