#!/bin/bash
# r02F (final): final evidence on one GPU after the consume-kernel change: bench (default arguments), launch list, full ncu of the headline kernels
set -u
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02F_bench.json 2> gpurun_out/r02F_bench.err; echo "bench rc=$?"; tail -c 200 gpurun_out/r02F_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02F_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --sub none > gpurun_out/r02F_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_vertex_normals_fan2|k_vv_consume_fan|k_vf_consume_fan" -s 30 -c 6 -o gpurun_out/r02F_hot -f \
    python bench.py --steps 2 --warmup 3 --no-cpu --sub none > gpurun_out/r02F_ncu_hot.log 2>&1; echo "ncu hot rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r02F_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["parity"]["ok"], {k:round(v["ms"],4) for k,v in d["kernels"].items()}, d["e2e"]["ms_per_step"], d["clocks"])
PY
