#!/bin/bash
# r02O: MCF mat-vec, batched loads with 40 registers (8 resident blocks of 192): tests, block sizes, launch list, bench sub-record
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mcf.py tests/test_gpu_shim.py tests/test_zz_reference_sources.py -m gpu -q --tb=short -k "mcf" > gpurun_out/r02O_mcf_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02O_mcf_pytest.log
for bt in 192 128 256; do RXM_MCF_BT=$bt timeout 300 python scripts/mcf_profile.py 2>&1 | grep " 1 " | sed "s/^/BT=$bt /" | sed "s/{.*}//"; done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_mcf -c 200 --csv --log-file gpurun_out/r02O_mcf_launches.csv python scripts/mcf_profile.py > gpurun_out/r02O_ncu.log 2>&1; echo "ncu rc=$?"
python profiles/summarize.py launches gpurun_out/r02O_mcf_launches.csv
timeout 600 python bench_configs.py --only bilateral > gpurun_out/r02O_bilateral_mcf.json 2> gpurun_out/r02O_bilateral_mcf.err; echo "bench rc=$?"
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/r02O_bilateral_mcf.json") if l.startswith("{")][-1]
m=d.get("mcf_cg_same_mesh", {})
for k,v in m.items():
    if isinstance(v, dict): print(k, {q: v.get(q) for q in ("iterations","ms_total","ms_per_iteration","hbm_frac","max_abs_diff_vs_oracle_f64","parity_ok")})
    else: print(k, v)
PY
