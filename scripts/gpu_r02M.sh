#!/bin/bash
# r02M: full ncu capture of the MCF kernels (one launch each of the mat-vec in both forms, the update, the setup)
set -u
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_mcf_matvec|k_mcf_update" -s 6 -c 4 -o gpurun_out/r02M_mcf -f python scripts/mcf_profile.py > gpurun_out/r02M_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r02M_ncu.log
