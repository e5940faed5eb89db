#!/bin/bash
# First GPU call of the next session: everything changed after the last GPU-verified commit.
#   gpurun --timeout 900 -- 'bash scripts/next_gpu_session.sh'            # 1 GPU
#   gpurun --gpus 2 --timeout 600 -- 'bash scripts/next_gpu_session.sh n2' # the 2-GPU warm-up fix (20M and 100M faces)
set -u
mkdir -p gpurun_out
if [ "${1:-}" = "n2" ]; then
    T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
    timeout 200 $T bench.py --gpus 2 --steps 20 --warmup 3 --faces 20000000 --no-cpu > gpurun_out/n2_q20.json 2> gpurun_out/n2_q20.err; echo "n2 20M rc=$?"
    timeout 300 $T bench.py --gpus 2 --steps 50 --warmup 3 --no-cpu > gpurun_out/n2_q100.json 2> gpurun_out/n2_q100.err; echo "n2 100M rc=$?"
    timeout 200 $T bench.py --gpus 2 --workload laplacian --faces 50000000 --steps 3 --warmup 3 --no-cpu > gpurun_out/n2_lap.json 2> gpurun_out/n2_lap.err; echo "n2 laplacian rc=$?"
    tail -c 400 gpurun_out/n2_q20.err gpurun_out/n2_q100.err gpurun_out/n2_lap.err
    exit 0
fi
python -m pytest tests -m gpu -q 2>&1 | tail -15
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py > gpurun_out/bench_next.json 2> gpurun_out/bench_next.err; tail -c 300 gpurun_out/bench_next.err
python bench_configs.py --only 1 > gpurun_out/cfg1_next.json 2> gpurun_out/cfg1_next.err
# launch list + full capture of the query-store kernels whose paths changed in round 1 (stored fans / FF rows / EF pairs)
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_next.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_next.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_query_store -c 8 -o gpurun_out/prof_next_store \
    python bench_configs.py --only 1 > gpurun_out/ncu_store.log 2>&1
ls -la gpurun_out | tail -8
