#!/usr/bin/env bash
# session 3 of round 2: GPU suite with the colour present pass and the vectorised mass reduction, default bench line
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/s3_pytest.log 2>&1; tail -3 gpurun_out/s3_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/s3_bench_default.json 2> gpurun_out/s3_bench_default.err; tail -3 gpurun_out/s3_bench_default.err
cut -c1-300 gpurun_out/s3_bench_default.json
