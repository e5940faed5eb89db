#!/bin/bash
# r02W: the rebuilt shim libraries (solve time in info[3]) pass their MCF tests; the MCF solve three ways on one mesh
set -u
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_shim.py tests/test_zz_reference_sources.py -m gpu -q --tb=short -k "mcf" > gpurun_out/r02W_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02W_pytest.log | cut -c1-160
timeout 80 python scripts/mcf_generic_vs_fixed.py > gpurun_out/r02W_mcf_three_ways.log 2>&1; echo "compare rc=$?"; tail -6 gpurun_out/r02W_mcf_three_ways.log | cut -c1-400
