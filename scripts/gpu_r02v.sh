#!/bin/bash
# r02v: Report JSON test on both shim builds, compute-sanitizer on the round-2 paths
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_shim.py tests/test_zz_reference_sources.py -m gpu -x -q > gpurun_out/r02v_shim.log 2>&1; echo "shim tests rc=$?"; tail -3 gpurun_out/r02v_shim.log
SANITIZE_SKIP_SHIM=1 timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize.py > gpurun_out/r02v_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ok$|ERROR SUMMARY" gpurun_out/r02v_memcheck.log | tail -9
SANITIZE_MULTI=1 timeout -s KILL 300 compute-sanitizer --tool memcheck python scripts/sanitize.py > gpurun_out/r02v_memcheck_multi.log 2>&1; echo "memcheck multi rc=$?"; grep -E "ok$|ERROR SUMMARY" gpurun_out/r02v_memcheck_multi.log | tail -4
nvidia-smi --query-gpu=utilization.gpu,memory.used --format=csv,noheader
SANITIZE_SKIP_SHIM=1 timeout 1200 compute-sanitizer --tool racecheck python scripts/sanitize.py > gpurun_out/r02v_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "ok$|RACECHECK SUMMARY|hazard" gpurun_out/r02v_racecheck.log | tail -9
