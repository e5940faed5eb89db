#!/bin/bash
# r02u: block size of the two-vertices-per-thread kernels (561-vertex tiles = 281 pairs: 128 threads need 3 rounds, 160 need 2)
set -u
mkdir -p gpurun_out
for v in default bt160 bt192 bt96; do
  if [ $v = default ]; then unset RXM_LIB; else export RXM_LIB=$PWD/rxmesh_b200/librxmesh_b200_$v.so; fi
  timeout 400 python bench.py --sub none --steps 50 --no-cpu > gpurun_out/r02u_bench_$v.json 2> gpurun_out/r02u_bench_$v.err
  timeout 300 python bench_configs.py --only queries > gpurun_out/r02u_q_$v.json 2> gpurun_out/r02u_q_$v.err
  timeout 300 python bench_configs.py --only laplacian --lap-faces 100000000 > gpurun_out/r02u_lap_$v.json 2> gpurun_out/r02u_lap_$v.err
  python - <<PY
import json
def last(p):
    try:
        return json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e:
        return None
b, q, l = last("gpurun_out/r02u_bench_$v.json"), last("gpurun_out/r02u_q_$v.json"), last("gpurun_out/r02u_lap_$v.json")
print("$v", "100M grid VN ms", b and round(b["kernels"]["VN"]["ms"], 4), "| Lloyd icosphere VN ms", q and round(q["consume_and_normals_on_lloyd_patches"]["VN"]["ms"], 4),
      "| Laplacian 100M ms/iter", l and round(l.get("ms_per_iteration", -1), 4), "| clocks", b and b["clocks"]["sm_mhz"])
PY
done
timeout 200 python -m pytest tests/test_multi.py -m gpu -x -q -k "long_cut" 2>&1 | tail -2
