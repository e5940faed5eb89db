#!/bin/bash
# r02n: resident blocks (register allocation) of k_vertex_normals_fan2 / k_laplacian_fan2: 8 (64 regs) / 10 (48) / 12 (40)
set -u
mkdir -p gpurun_out
for m in 8 10 12; do
  RXM_VN_MINB=$m timeout 400 python bench.py --sub none --steps 50 --no-cpu > gpurun_out/r02n_bench_$m.json 2> gpurun_out/r02n_bench_$m.err
  RXM_VN_MINB=$m timeout 300 python bench_configs.py --only queries > gpurun_out/r02n_q_$m.json 2> gpurun_out/r02n_q_$m.err
  RXM_VN_MINB=$m timeout 300 python bench_configs.py --only laplacian --lap-faces 100000000 > gpurun_out/r02n_lap_$m.json 2> gpurun_out/r02n_lap_$m.err
  python - <<PY
import json
def last(p):
    try:
        return json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e:
        return None
b, q, l = last("gpurun_out/r02n_bench_$m.json"), last("gpurun_out/r02n_q_$m.json"), last("gpurun_out/r02n_lap_$m.json")
print("MINB=$m", "100M grid VN ms", b and round(b["kernels"]["VN"]["ms"], 4), "| Lloyd icosphere VN ms", q and round(q["consume_and_normals_on_lloyd_patches"]["VN"]["ms"], 4),
      "| Laplacian 100M ms/iter", l and round(l.get("ms_per_iteration", -1), 4))
PY
done
