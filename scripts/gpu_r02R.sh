#!/bin/bash
# r02R: after the fix of the pad-zeros / ribbon-rows race in k_mcf_matvec: every MCF test (twice), racecheck + memcheck with all four solver forms
set -u
mkdir -p gpurun_out
for rep in 1 2; do timeout 600 python -m pytest tests/test_mcf.py tests/test_gpu_shim.py tests/test_zz_reference_sources.py -m gpu -q --tb=short -k "mcf" > gpurun_out/r02R_mcf_pytest_$rep.log 2>&1; echo "pytest $rep rc=$?"; tail -12 gpurun_out/r02R_mcf_pytest_$rep.log | cut -c1-200; done
timeout 300 compute-sanitizer --tool racecheck python scripts/sanitize_mcf.py > gpurun_out/r02R_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "ok$|RACECHECK SUMMARY|hazard" gpurun_out/r02R_racecheck.log | tail -6
timeout 300 compute-sanitizer --tool memcheck python scripts/sanitize_mcf.py > gpurun_out/r02R_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ok$|ERROR SUMMARY|Invalid" gpurun_out/r02R_memcheck.log | tail -4
