#!/usr/bin/env bash
# session 3: ncu --set full of the FINAL sweep kernel (16384^2, 4096^2)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_frame2 -s 4 -c 2 -f -o gpurun_out/r02d_frame2_16384 python bench.py --config 3 --steps 6 --warmup 4 --e2e-steps 0 --cpu-seconds 0 --no-secondary --no-graph > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_frame2 -s 4 -c 2 -f -o gpurun_out/r02d_frame2_4096 python bench.py --config 2 --steps 6 --warmup 4 --e2e-steps 0 --cpu-seconds 0 --no-secondary --no-graph > /dev/null 2>&1
ls -la gpurun_out/r02d_*
