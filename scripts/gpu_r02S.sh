#!/bin/bash
# r02S: final verification of the round-2 tree (MCF with the Jacobi form, race fix): all GPU tests, smoke, the default bench, MCF launch list
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02S_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02S_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02S_bench.json 2> gpurun_out/r02S_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02S_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_mcf -c 200 --csv --log-file gpurun_out/r02S_mcf_launches.csv python scripts/mcf_profile.py > gpurun_out/r02S_ncu.log 2>&1; echo "ncu rc=$?"
python profiles/summarize.py launches gpurun_out/r02S_mcf_launches.csv
python - <<PY
import json
d=json.loads(open("gpurun_out/r02S_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["parity"]["ok"], {k:round(v["ms"],4) for k,v in d["kernels"].items()}, d["e2e"]["ms_per_step"], d["clocks"])
c=d["configs"]; print({k:(v.get("parity_ok"), v.get("wall_seconds"), v.get("error")) for k,v in c.items()})
m=c["3_bilateral_10m_torus"].get("mcf_cg_same_mesh",{})
for k,v in m.items():
    if isinstance(v, dict): print(k, {q: v.get(q) for q in ("iterations","ms_total","ms_per_iteration","hbm_frac","max_abs_diff_vs_oracle_f64","parity_ok")})
    else: print(k, v)
l=d["laplacian_400m"]; print(l.get("ms_per_iteration"), l.get("parity",{}).get("ok"), l.get("wall_seconds"))
PY
