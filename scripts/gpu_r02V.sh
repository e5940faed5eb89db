#!/bin/bash
# r02V: full ncu capture of the final MCF kernels (mat-vec uniform + cotangent, update)
set -u
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_mcf_matvec|k_mcf_update" -s 6 -c 2 -o gpurun_out/r02V_mcf_uniform -f python scripts/mcf_profile.py > gpurun_out/r02V_ncu1.log 2>&1; echo "ncu rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_mcf_matvec<0" -s 4 -c 1 -o gpurun_out/r02V_mcf_cot -f python scripts/mcf_profile.py > gpurun_out/r02V_ncu2.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r02V_ncu2.log
