#!/bin/bash
# r02U: weights-from-global is the default of the cotangent mat-vec: MCF tests (all three files), sanitizers, bench sub-record
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mcf.py tests/test_gpu_shim.py tests/test_zz_reference_sources.py -m gpu -q --tb=short -k "mcf" > gpurun_out/r02U_mcf_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02U_mcf_pytest.log | cut -c1-200
timeout 300 compute-sanitizer --tool racecheck python scripts/sanitize_mcf.py > gpurun_out/r02U_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "ok$|RACECHECK SUMMARY" gpurun_out/r02U_racecheck.log | tail -3
timeout 300 compute-sanitizer --tool memcheck python scripts/sanitize_mcf.py > gpurun_out/r02U_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ok$|ERROR SUMMARY" gpurun_out/r02U_memcheck.log | tail -3
timeout 600 python bench_configs.py --only bilateral > gpurun_out/r02U_bilateral_mcf.json 2> gpurun_out/r02U_bilateral_mcf.err; echo "bench rc=$?"
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/r02U_bilateral_mcf.json") if l.startswith("{")][-1]
m=d.get("mcf_cg_same_mesh", {})
for k,v in m.items():
    if isinstance(v, dict): print(k, {q: v.get(q) for q in ("iterations","ms_total","ms_per_iteration","hbm_frac","max_abs_diff_vs_oracle_f64","parity_ok")})
    else: print(k, v)
PY
