#!/usr/bin/env bash
# session 3, 8-GPU box: the driver's commands at N = 8 and N = 1 with the final sweep geometry (packed last column, tall blocks at N = 1)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/s3_bench_n8.json 2> gpurun_out/s3_bench_n8.err
python -c "import json; d=json.loads([l for l in open('gpurun_out/s3_bench_n8.json').read().strip().splitlines() if l.startswith('{')][-1]); print('N=8', round(d['value']), round(d['macro_on']['value']), d.get('multirank_parity'), d['clocks'], round(d['e2e']['value']), d['roofline'].get('dram_frac'))" || tail -5 gpurun_out/s3_bench_n8.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --cpu-seconds 0 > gpurun_out/s3_bench_n1_8box.json 2> gpurun_out/s3_bench_n1_8box.err
python -c "import json; d=json.loads([l for l in open('gpurun_out/s3_bench_n1_8box.json').read().strip().splitlines() if l.startswith('{')][-1]); print('N=1', round(d['value']), round(d['macro_on']['value']), d['clocks'], round(d['e2e']['value']))" || tail -5 gpurun_out/s3_bench_n1_8box.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 8 --steps 200 --warmup 20 --e2e-steps 0 --cpu-seconds 0 > gpurun_out/s3_bench_long_n8.json 2>/dev/null
python -c "import json; d=json.loads([l for l in open('gpurun_out/s3_bench_long_n8.json').read().strip().splitlines() if l.startswith('{')][-1]); print('N=8 long', round(d['value']), d['clocks'])"
