#!/usr/bin/env bash
# compute-sanitizer memcheck over the paths added in session 3 (single-slab cases: slabs in one process need concurrent kernels)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_fused.py tests/test_gpu_frames.py -m gpu -q -x -k "(packed and 1-0) or (packed and 1-3) or present or curl or particle_frames or rust" > gpurun_out/r02_s3_compute_sanitizer_memcheck.log 2>&1
grep -E "passed|failed|ERROR SUMMARY" gpurun_out/r02_s3_compute_sanitizer_memcheck.log | tail -3
