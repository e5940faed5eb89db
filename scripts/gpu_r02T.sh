#!/bin/bash
# r02T: cotangent mat-vec with the weights read from global memory (RXM_MCF_WSTAGE=0) against the staged form
set -u
mkdir -p gpurun_out
RXM_MCF_WSTAGE=0 timeout 300 python -m pytest tests/test_mcf.py -m gpu -q --tb=short > gpurun_out/r02T_mcf_pytest_w0.log 2>&1; echo "pytest WSTAGE=0 rc=$?"; tail -3 gpurun_out/r02T_mcf_pytest_w0.log | cut -c1-200
timeout 300 python -m pytest tests/test_mcf.py -m gpu -q --tb=short > gpurun_out/r02T_mcf_pytest_w1.log 2>&1; echo "pytest default rc=$?"; tail -2 gpurun_out/r02T_mcf_pytest_w1.log | cut -c1-200
for w in 1 0; do RXM_MCF_WSTAGE=$w timeout 300 python scripts/mcf_profile.py 2>&1 | grep "cotangent" | sed "s/^/WSTAGE=$w /" | sed "s/{.*}//"; done
RXM_MCF_WSTAGE=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_mcf_matvec -c 80 --csv --log-file gpurun_out/r02T_mcf_launches_w0.csv python scripts/mcf_profile.py > gpurun_out/r02T_ncu.log 2>&1; echo "ncu rc=$?"
python profiles/summarize.py launches gpurun_out/r02T_mcf_launches_w0.csv
