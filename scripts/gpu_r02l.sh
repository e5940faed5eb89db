#!/bin/bash
# r02l: trimmed bilateral fast path (register sweep), (t, t + half) pairing in VN / Laplacian, full GPU suite, bench
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02l_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02l_pytest.log
run() { # tag, env...
  tag=$1; shift
  env "$@" timeout 200 python bench_configs.py --only bilateral > gpurun_out/r02l_bil_$tag.json 2> gpurun_out/r02l_bil_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02l_bil_$tag.json").read().strip().splitlines()[-1])
    print("$tag", "ms/iter", round(d["ms_per_iteration"], 4), "parity", d.get("parity_ok"))
except Exception as e:
    print("$tag", "failed", e)
PY
}
run default X=1
run minb4 RXM_BILATERAL_MINB=4
run minb2 RXM_BILATERAL_MINB=2
run bt576 RXM_BILATERAL_BT=576
run bt384 RXM_BILATERAL_BT=384
run bt256 RXM_BILATERAL_BT=256
timeout 300 python bench_configs.py --only queries > gpurun_out/r02l_queries.json 2> gpurun_out/r02l_queries.err; echo "queries rc=$?"
timeout 600 python bench.py > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02l_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r02l_bench.json").read().strip().splitlines()[-1])
print({k: round(v["ms"], 4) for k, v in d["kernels"].items()}, "e2e", round(d["e2e"]["ms_per_step"], 2), "copy-only", round(d["e2e"]["copy_only_ms_per_step"], 2))
q = json.loads(open("gpurun_out/r02l_queries.json").read().strip().splitlines()[-1])
print({k: (round(v["ms"], 4), round(v["hbm_frac"], 3)) for k, v in q["consume_and_normals_on_lloyd_patches"].items() if isinstance(v, dict)})
PY
