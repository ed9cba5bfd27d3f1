#!/usr/bin/env bash
# session 3: the Rust-golden click/drag test on a handle, and ncu of the derived-field passes (k_present, k_curl, k_mass)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_frames.py -m gpu -q -x -k "rust or present or curl or particle_frames" > gpurun_out/s3_pytest_rust.log 2>&1; tail -2 gpurun_out/s3_pytest_rust.log
python tools/present_mass_probe.py
ncu --set full --clock-control none --import-source on -k regex:"k_present|k_mass|k_curl" -c 12 -f -o gpurun_out/r02_s3_derived python tools/present_mass_probe.py > /dev/null 2>&1
ls -la gpurun_out/r02_s3_derived.ncu-rep
