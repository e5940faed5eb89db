#!/bin/bash
# r02k: single-process multi-GPU mode on one device + bitmap-free bilateral kernel (block size / register sweep)
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_multi.py -m gpu -x -q > gpurun_out/r02k_multi.log 2>&1; echo "multi rc=$?"; tail -3 gpurun_out/r02k_multi.log
timeout 600 python -m pytest tests/test_gpu_apps.py tests/test_gpu_shim.py -m gpu -x -q -k "bilateral" > gpurun_out/r02k_bil_tests.log 2>&1; echo "bil tests rc=$?"; tail -3 gpurun_out/r02k_bil_tests.log
run() { # tag, env...
  tag=$1; shift
  env "$@" timeout 200 python bench_configs.py --only bilateral > gpurun_out/r02k_bil_$tag.json 2> gpurun_out/r02k_bil_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02k_bil_$tag.json").read().strip().splitlines()[-1])
    print("$tag", "ms/iter", round(d["ms_per_iteration"], 4), "parity", d.get("parity_ok"), "deferred", d.get("deferred"))
except Exception as e:
    print("$tag", "failed", e)
PY
}
run default X=1
run relaxed RXM_BILATERAL_RELAXED=1
run bt576 RXM_BILATERAL_BT=576
run bt576r RXM_BILATERAL_BT=576 RXM_BILATERAL_RELAXED=1
run bt384 RXM_BILATERAL_BT=384
run bt288 RXM_BILATERAL_BT=288
run bt256 RXM_BILATERAL_BT=256
run bt192 RXM_BILATERAL_BT=192
