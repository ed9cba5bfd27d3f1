#!/usr/bin/env bash
# session 3 sanity on 2 GPUs: cross-process parity tests and the driver's N = 2 command after the last library changes
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q > gpurun_out/s3_pytest_multirank.log 2>&1; tail -2 gpurun_out/s3_pytest_multirank.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/s3_bench_n2.json 2> gpurun_out/s3_bench_n2.err
tail -2 gpurun_out/s3_bench_n2.err; cut -c1-200 gpurun_out/s3_bench_n2.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/s3_bench_ref_n2.json 2> gpurun_out/s3_bench_ref_n2.err; cut -c1-300 gpurun_out/s3_bench_ref_n2.json
