#!/bin/bash
# r02J: first GPU run of the MCF solver: parity tests, sanitizer, bench sub-record on the 10 M-face noisy torus
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mcf.py -m gpu -q --tb=short > gpurun_out/r02J_mcf_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r02J_mcf_pytest.log
timeout 300 compute-sanitizer --tool memcheck python scripts/sanitize_mcf.py > gpurun_out/r02J_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ok$|ERROR SUMMARY|Invalid|uniform|cotangent" gpurun_out/r02J_memcheck.log | tail -8
timeout 300 compute-sanitizer --tool racecheck python scripts/sanitize_mcf.py > gpurun_out/r02J_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "ok$|RACECHECK SUMMARY|hazard" gpurun_out/r02J_racecheck.log | tail -6
timeout 600 python bench_configs.py --only bilateral > gpurun_out/r02J_bilateral_mcf.json 2> gpurun_out/r02J_bilateral_mcf.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r02J_bilateral_mcf.err
python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/r02J_bilateral_mcf.json") if l.startswith("{")][-1]
    b=d.get("bilateral", d)
    m=b.get("mcf_cg_same_mesh", {})
    print("bilateral ms/iter", b.get("ms_per_iteration"), b.get("parity_ok"))
    for k,v in m.items():
        if isinstance(v, dict): print(k, {q: v.get(q) for q in ("iterations","converged","ms_total","ms_per_iteration","achieved_gbs","hbm_frac","max_abs_diff_vs_oracle_f64","tolerance_abs","parity_ok","true_residual_sq_of_gpu_result")}, v.get("oracle"), v.get("cpu_baseline"))
        else: print(k, v)
except Exception as e:
    print("summary failed", e)
PY
