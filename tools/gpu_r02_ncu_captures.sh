#!/usr/bin/env bash
# ncu captures: launch list of the default bench command; --set full of the plain 16384^2 sweep, the 4096^2 sweep,
# the masked 8192^2 porous sweep (with tracers), and the multi-slab instance on a 16384x2048 slab (8 slabs, one GPU)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_default.csv python bench.py --steps 20 --warmup 5 --cpu-seconds 0 > gpurun_out/c12_launch_bench.log 2>&1
tail -2 gpurun_out/c12_launch_bench.log | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:k_frame2 -s 4 -c 2 -f -o gpurun_out/r02_frame2_16384 python bench.py --config 3 --steps 6 --warmup 4 --e2e-steps 0 --cpu-seconds 0 --no-secondary --no-graph > gpurun_out/c12_ncu1.log 2>&1; tail -1 gpurun_out/c12_ncu1.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:k_frame2 -s 4 -c 2 -f -o gpurun_out/r02_frame2_4096 python bench.py --config 2 --steps 6 --warmup 4 --e2e-steps 0 --cpu-seconds 0 --no-secondary --no-graph > gpurun_out/c12_ncu2.log 2>&1; tail -1 gpurun_out/c12_ncu2.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:k_frame2 -s 4 -c 2 -f -o gpurun_out/r02_frame2_porous python bench.py --config 5 --steps 6 --warmup 4 --e2e-steps 0 --cpu-seconds 0 --no-secondary --no-graph > gpurun_out/c12_ncu3.log 2>&1; tail -1 gpurun_out/c12_ncu3.log | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_cfg5.csv python bench.py --config 5 --steps 6 --warmup 4 --e2e-steps 0 --cpu-seconds 0 --no-secondary --no-graph > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_frame2 -s 24 -c 2 -f -o gpurun_out/r02_frame2_slab python tools/slabs_one_gpu.py 16384 16384 8 8 > gpurun_out/c12_ncu4.log 2>&1; tail -2 gpurun_out/c12_ncu4.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
