#!/bin/bash
# r02m: GPU Lloyd patcher (equality with the host passes, timing at 10 M / 100 M faces), ncu of the bilateral kernel and of
# vertex normals on Lloyd patches
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_patcher.py -m gpu -x -q > gpurun_out/r02m_patcher.log 2>&1; echo "patcher tests rc=$?"; tail -15 gpurun_out/r02m_patcher.log
timeout 900 python - > gpurun_out/r02m_patcher_timing.txt 2>&1 <<PY
import os, time, json
import numpy as np
import rxmesh_b200 as rx
from rxmesh_b200 import meshio
rx.rx_init(0)
os.environ["RXM_VERBOSE"] = "1"
out = {}
for name, mk in (("icosphere_10m", lambda: meshio.icosphere(707)), ("grid_100m", lambda: meshio.grid(7072, 7072))):
    V, F = mk()
    del V
    res = {}
    for tag, env in (("gpu", "1"), ("host", "0")):
        if tag == "host" and name == "grid_100m":
            continue  # 50 s, measured in profiles/r02i_lloyd_100m.json
        os.environ["RXM_PATCHER_GPU"] = env
        t = time.time()
        m = rx.RXMeshStatic(F, device=False, patch_size=1024)
        res[tag] = {"build_s": time.time() - t, "patcher_s": m.build_seconds(True), "patches": m.get_num_patches()}
        fp = m.elem_patch(2).copy()
        if tag == "gpu":
            keep = fp
        else:
            res["equal"] = bool(np.array_equal(keep, fp))
        del m
    out[name] = res
    print(json.dumps({name: res}), flush=True)
PY
tail -12 gpurun_out/r02m_patcher_timing.txt
ncu --set full --clock-control none --import-source on -k regex:k_bilateral_patch -c 2 -o gpurun_out/r02m_bil -f \
    python bench_configs.py --only bilateral > gpurun_out/r02m_ncu_bil.log 2>&1; echo "ncu bil rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_vertex_normals_fan2 -c 2 -o gpurun_out/r02m_vn_lloyd -f \
    python bench_configs.py --only queries > gpurun_out/r02m_ncu_vn.log 2>&1; echo "ncu vn rc=$?"
ls -la gpurun_out/r02m*
