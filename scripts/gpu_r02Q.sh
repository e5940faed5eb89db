#!/bin/bash
# r02Q: the Jacobi-preconditioned MCF solve (rxm_mcf_solve_ex, PCGMatFreeAttrSolver header): every MCF test, sanitizer, bench sub-record
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mcf.py tests/test_gpu_shim.py tests/test_zz_reference_sources.py -m gpu -q --tb=short -k "mcf" > gpurun_out/r02Q_mcf_pytest.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r02Q_mcf_pytest.log
timeout 600 python bench_configs.py --only bilateral > gpurun_out/r02Q_bilateral_mcf.json 2> gpurun_out/r02Q_bilateral_mcf.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r02Q_bilateral_mcf.err
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/r02Q_bilateral_mcf.json") if l.startswith("{")][-1]
m=d.get("mcf_cg_same_mesh", {})
for k,v in m.items():
    if isinstance(v, dict): print(k, {q: v.get(q) for q in ("iterations","ms_total","ms_per_iteration","hbm_frac","max_abs_diff_vs_oracle_f64","tolerance_abs","parity_ok")}, v.get("oracle"))
    else: print(k, v)
PY
