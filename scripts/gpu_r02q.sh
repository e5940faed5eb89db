#!/bin/bash
# r02q: evidence pass on one GPU: full GPU suite, smoke, bench (default arguments), launch list + full ncu capture of the hot kernels
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02q_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02q_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02q_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02q_bench_ref.json 2> gpurun_out/r02q_bench_ref.err; echo "ref rc=$?"; tail -c 400 gpurun_out/r02q_bench_ref.json
RXM_VERBOSE=1 timeout 300 python bench_configs.py --only lloyd100m > gpurun_out/r02q_lloyd100m.json 2> gpurun_out/r02q_lloyd100m.err; echo "lloyd100m rc=$?"; grep "lloyd:\|build:" gpurun_out/r02q_lloyd100m.err | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02q_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --sub none > gpurun_out/r02q_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_vertex_normals_fan2|k_vv_consume_fan|k_vf_consume_fan" -s 30 -c 6 -o gpurun_out/r02q_hot -f \
    python bench.py --steps 2 --warmup 3 --no-cpu --sub none > gpurun_out/r02q_ncu_hot.log 2>&1; echo "ncu hot rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_laplacian_fan2 -s 3 -c 2 -o gpurun_out/r02q_lap -f \
    python bench_configs.py --only laplacian --lap-faces 100000000 > gpurun_out/r02q_ncu_lap.log 2>&1; echo "ncu lap rc=$?"
ls -la gpurun_out/r02q*
