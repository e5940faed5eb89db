#!/bin/bash
# r02L: MCF after the kernel / host-loop changes: parity tests (fixed-function + the drop-in CG solver over user kernels and
# over the reference's own mcf_kernels.cuh), bench sub-record, ncu launch list
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mcf.py tests/test_gpu_shim.py tests/test_zz_reference_sources.py -m gpu -q --tb=short -k "mcf" > gpurun_out/r02L_mcf_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r02L_mcf_pytest.log
timeout 600 python bench_configs.py --only bilateral > gpurun_out/r02L_bilateral_mcf.json 2> gpurun_out/r02L_bilateral_mcf.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r02L_bilateral_mcf.err
python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/r02L_bilateral_mcf.json") if l.startswith("{")][-1]
    m=d.get("mcf_cg_same_mesh", {})
    print("bilateral ms/iter", d.get("ms_per_iteration"), d.get("parity_ok"))
    for k,v in m.items():
        if isinstance(v, dict): print(k, {q: v.get(q) for q in ("iterations","converged","ms_total","ms_per_iteration","achieved_gbs","hbm_frac","max_abs_diff_vs_oracle_f64","tolerance_abs","parity_ok")}, v.get("oracle"), v.get("cpu_baseline"))
        else: print(k, v)
except Exception as e:
    print("summary failed", e)
PY
timeout 300 python scripts/mcf_profile.py > gpurun_out/r02L_mcf_wall.log 2>&1; echo "wall rc=$?"; tail -4 gpurun_out/r02L_mcf_wall.log
for bt in 128 256; do RXM_MCF_BT=$bt timeout 300 python scripts/mcf_profile.py 2>&1 | tail -4 | sed "s/^/BT=$bt /"; done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_mcf -c 200 --csv --log-file gpurun_out/r02L_mcf_launches.csv python scripts/mcf_profile.py > gpurun_out/r02L_ncu.log 2>&1; echo "ncu rc=$?"
python profiles/summarize.py launches gpurun_out/r02L_mcf_launches.csv
