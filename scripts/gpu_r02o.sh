#!/bin/bash
# r02o: L2 bulk prefetch of the patch that takes over a block's slot (k_vertex_normals_fan2 / k_laplacian_fan2), distance sweep
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_apps.py tests/test_multi.py tests/test_gpu_large.py -m gpu -x -q > gpurun_out/r02o_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02o_pytest.log
for d in 0 default 740 2960 5920; do
  if [ $d = default ]; then unset RXM_PREFETCH_DIST; else export RXM_PREFETCH_DIST=$d; fi
  timeout 400 python bench.py --sub none --steps 50 --no-cpu > gpurun_out/r02o_bench_$d.json 2> gpurun_out/r02o_bench_$d.err
  timeout 300 python bench_configs.py --only queries > gpurun_out/r02o_q_$d.json 2> gpurun_out/r02o_q_$d.err
  timeout 300 python bench_configs.py --only laplacian --lap-faces 100000000 > gpurun_out/r02o_lap_$d.json 2> gpurun_out/r02o_lap_$d.err
  python - <<PY
import json
def last(p):
    try:
        return json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e:
        return None
b, q, l = last("gpurun_out/r02o_bench_$d.json"), last("gpurun_out/r02o_q_$d.json"), last("gpurun_out/r02o_lap_$d.json")
print("DIST=$d", "100M grid VN ms", b and round(b["kernels"]["VN"]["ms"], 4), "| Lloyd icosphere VN ms", q and round(q["consume_and_normals_on_lloyd_patches"]["VN"]["ms"], 4),
      "| Laplacian 100M ms/iter", l and round(l.get("ms_per_iteration", -1), 4))
PY
done
