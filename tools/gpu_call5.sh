#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/repro_porous.py 8192 8192 0 > gpurun_out/c5_repro_8192.log 2>&1; tail -2 gpurun_out/c5_repro_8192.log
python tools/repro_porous.py 8192 8192 1 > gpurun_out/c5_repro_8192_macro.log 2>&1; tail -2 gpurun_out/c5_repro_8192_macro.log
python tools/repro_porous.py 1024 512 0 > gpurun_out/c5_repro_1024.log 2>&1; tail -2 gpurun_out/c5_repro_1024.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tools/repro_porous.py 1024 512 0 > gpurun_out/c5_sanitizer_1024.log 2>&1; grep -E "Invalid|at |by |ERROR SUMMARY|Address" gpurun_out/c5_sanitizer_1024.log | head -30
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python tools/repro_porous.py 8192 1024 0 > gpurun_out/c5_sanitizer_8192x1024.log 2>&1; grep -E "Invalid|at |by |ERROR SUMMARY|Address" gpurun_out/c5_sanitizer_8192x1024.log | head -30
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/c5_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c5_pytest.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/c5_pytest.log | tail -30
