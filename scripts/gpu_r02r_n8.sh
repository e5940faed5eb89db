#!/bin/bash
# r02r: 8 GPUs -- single-process multi-GPU mode (tests + bench_multi), then bench.py as the driver launches it
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_multi.py tests/test_distributed.py -m gpu -x -q > gpurun_out/r02r_n8_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r02r_n8_tests.log
timeout 400 python scripts/bench_multi.py 1000 1024 > gpurun_out/r02r_multi_n8.json 2> gpurun_out/r02r_multi_n8.err; echo "multi rc=$?"; cut -c1-1200 gpurun_out/r02r_multi_n8.json; tail -2 gpurun_out/r02r_multi_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 100 --warmup 3 > gpurun_out/r02r_bench_n8.json 2> gpurun_out/r02r_bench_n8.err; echo "n8 rc=$?"; tail -c 400 gpurun_out/r02r_bench_n8.err
python - <<PY
import json
for f in ("gpurun_out/r02r_bench_n8.json",):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        l = d.get("laplacian_400m", {})
        print(f, "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 1), "copy-only", round(d["e2e"].get("copy_only_ms_per_step", 0), 1),
              "| lap ms/iter", l.get("ms_per_iteration"), "speedup", l.get("speedup_vs_n1"), "transport", l.get("halo_transport"), "parity", l.get("parity", {}).get("ok"), l.get("parity", {}).get("fused_equals_nccl_bitwise"))
    except Exception as e:
        print(f, "failed", e)
PY
