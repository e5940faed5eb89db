"""Single-process multi-GPU mode (rxm_multi_*, include/rxmesh/rxmesh_multi.h) on every visible device: iterated Laplacian on a
Lloyd-patched icosphere, time per iteration = difference of two calls with different iteration counts (upload / download
cancel), result against the one-device run (bit-identical) -- python scripts/bench_multi.py [nu] [patch_size]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import rxmesh_b200 as rx  # noqa: E402
from rxmesh_b200 import meshio  # noqa: E402
from rxmesh_b200.multi import RXMeshMulti  # noqa: E402

nu = int(sys.argv[1]) if len(sys.argv) > 1 else 707
ps = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
ndev = torch.cuda.device_count()
rx.rx_init(0)
V, F = meshio.icosphere(nu)
V = V.astype(np.float32)
out = {"mesh": "icosphere nu=%d, %d faces, Lloyd patches <= %d" % (nu, F.shape[0], ps), "devices": ndev, "runs": {}}
ref = None
for devs in ([0], list(range(ndev))) if ndev > 1 else ([0], [0, 0]):
    t = time.time()
    mm = RXMeshMulti(F, devs, patch_size=ps)
    t_build = time.time() - t
    mm.laplacian_smooth(V, 0.01, 5)
    k0, k1 = 20, 220
    t = time.time(); a = mm.laplacian_smooth(V, 0.01, k0); t0 = time.time() - t
    t = time.time(); b = mm.laplacian_smooth(V, 0.01, k1); t1 = time.time() - t
    rec = {"shards": len(devs), "devices": devs, "build_s": t_build, "mirrored_vertices": mm.halo_elements(),
           "ms_per_iteration": (t1 - t0) / (k1 - k0) * 1e3, "call_s_20_iters": t0, "call_s_220_iters": t1}
    if ref is None:
        ref = b
    else:
        rec["equals_one_device_bitwise"] = bool(np.array_equal(ref, b))
    out["runs"]["x".join(str(d) for d in devs) if len(devs) < 3 else "%d_devices" % len(devs)] = rec
    del mm
r = list(out["runs"].values())
out["speedup"] = r[0]["ms_per_iteration"] / r[1]["ms_per_iteration"]
print(json.dumps(out))
