#!/usr/bin/env bash
# packed last strip column on slabs: parity (slabs in one process), slab-sized lattice, then 2 GPUs with the driver's command
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_frames.py tests/test_gpu_fullsize.py tests/test_gpu_multirank.py -m gpu -q -x > gpurun_out/s3_pytest_packslabs.log 2>&1; tail -3 gpurun_out/s3_pytest_packslabs.log
python tools/slabs_one_gpu.py 16384 16384 8 16
LBM_FUSE_PACK=0 python tools/slabs_one_gpu.py 16384 16384 8 16
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/s3_bench_n2.json 2> gpurun_out/s3_bench_n2.err
python -c "import json; d=json.loads([l for l in open('gpurun_out/s3_bench_n2.json').read().strip().splitlines() if l.startswith('{')][-1]); print('N=2', round(d['value']), round(d['macro_on']['value']), d.get('multirank_parity'), d['clocks'])" || tail -5 gpurun_out/s3_bench_n2.err
LBM_FUSE_PACK=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 --e2e-steps 0 --cpu-seconds 0 2>/dev/null | python -c "import json,sys; d=json.loads([l for l in sys.stdin.read().strip().splitlines() if l.startswith('{')][-1]); print('N=2 no pack', round(d['value']), round(d['macro_on']['value']))"
