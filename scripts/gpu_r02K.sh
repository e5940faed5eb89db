#!/bin/bash
# r02K: where the time of an MCF solve goes: wall times, then the ncu launch list of the k_mcf kernels
set -u
mkdir -p gpurun_out
timeout 300 python scripts/mcf_profile.py > gpurun_out/r02K_mcf_wall.log 2>&1; echo "wall rc=$?"; cat gpurun_out/r02K_mcf_wall.log | tail -6
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_mcf -c 200 --csv --log-file gpurun_out/r02K_mcf_launches.csv python scripts/mcf_profile.py > gpurun_out/r02K_ncu.log 2>&1; echo "ncu rc=$?"
python profiles/summarize.py launches gpurun_out/r02K_mcf_launches.csv
