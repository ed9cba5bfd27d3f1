#!/usr/bin/env bash
# final 1-GPU evidence of the round: GPU tests, smoke, sanitizer runs, ncu captures of the final kernels, default bench
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/c17_pytest.log 2>&1; tail -2 gpurun_out/c17_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
# compute-sanitizer on the new paths (small cases): memcheck + racecheck
timeout 1500 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_frames.py tests/test_gpu_fused.py -m gpu -q -x -k "particle_frames or macro_texture or 244 or 120-40 or curl or pipelined or obstacle_painted or previous_buffer" > gpurun_out/r02_compute_sanitizer_memcheck.log 2>&1; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/r02_compute_sanitizer_memcheck.log | tail -3
timeout 1500 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_frames.py -m gpu -q -x -k "particle_frames and 120" > gpurun_out/r02_compute_sanitizer_racecheck.log 2>&1; grep -E "passed|failed|RACECHECK SUMMARY" gpurun_out/r02_compute_sanitizer_racecheck.log | tail -3
# ncu: final kernels
ncu --set full --clock-control none --import-source on -k regex:k_frame2 -s 4 -c 2 -f -o gpurun_out/r02b_frame2_16384 python bench.py --config 3 --steps 6 --warmup 4 --e2e-steps 0 --cpu-seconds 0 --no-secondary --no-graph > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_frame2 -s 4 -c 2 -f -o gpurun_out/r02b_frame2_4096 python bench.py --config 2 --steps 6 --warmup 4 --e2e-steps 0 --cpu-seconds 0 --no-secondary --no-graph > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_frame2 -s 4 -c 2 -f -o gpurun_out/r02b_frame2_porous python bench.py --config 5 --steps 6 --warmup 4 --e2e-steps 0 --cpu-seconds 0 --no-secondary --no-graph > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches_bench_default.csv python bench.py --steps 20 --warmup 5 --cpu-seconds 0 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02b_launches_cfg5.csv python bench.py --config 5 --steps 6 --warmup 4 --e2e-steps 0 --cpu-seconds 0 --no-secondary --no-graph > /dev/null 2>&1
ls -la gpurun_out/r02b_*
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/c17_bench_default.json 2> gpurun_out/c17_bench_default.err; tail -3 gpurun_out/c17_bench_default.err
( time python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/c17_bench_ref.json 2> gpurun_out/c17_bench_ref.err; tail -3 gpurun_out/c17_bench_ref.err
for cfg in 1 2 5; do python bench.py --config $cfg --steps 200 --warmup 20 --no-secondary > gpurun_out/c17_bench_cfg$cfg.json 2>/dev/null; done
python bench.py --config 2 --steps 200 --warmup 20 --no-secondary --no-fuse --e2e-steps 0 --cpu-seconds 0 > gpurun_out/c17_bench_cfg2_nofuse.json 2>/dev/null
cat gpurun_out/c17_bench_default.json | cut -c1-600
