#!/usr/bin/env bash
# frames with tracer particles: particle passes on a side stream beside the next sweep (captured runs of 8 frames), A/B
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_frames.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/s3_pytest_frames.log 2>&1; tail -3 gpurun_out/s3_pytest_frames.log
for ov in 1 0; do for cfg in 1 5; do
  LBM_PARTICLE_OVERLAP=$ov python bench.py --config $cfg --steps 400 --warmup 40 --no-secondary --e2e-steps 0 --cpu-seconds 0 > gpurun_out/s3_cfg${cfg}_ov${ov}.json 2>gpurun_out/s3_cfg${cfg}_ov${ov}.err
  python -c "import json,sys; d=json.loads(open('gpurun_out/s3_cfg${cfg}_ov${ov}.json').read().strip().splitlines()[-1]); print('cfg', $cfg, 'overlap', $ov, d['value'], d['ms_per_step'], d['gpu_launches'])"
done; done
