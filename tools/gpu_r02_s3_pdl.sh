#!/usr/bin/env bash
# programmatic dependent launch of the sweeps (LBM_PDL=1): parity tests, then A/B on configs 1, 2, 3, 5
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LBM_PDL=1 timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_frames.py -m gpu -q -x > gpurun_out/s3_pytest_pdl.log 2>&1; tail -2 gpurun_out/s3_pytest_pdl.log
for pdl in 0 1; do for cfg in 1 2 3 5; do
  LBM_PDL=$pdl python bench.py --config $cfg --steps 400 --warmup 40 --no-secondary --e2e-steps 0 --cpu-seconds 0 > gpurun_out/s3_cfg${cfg}_pdl${pdl}.json 2>gpurun_out/s3_cfg${cfg}_pdl${pdl}.err
  python -c "import json,sys; d=json.loads(open('gpurun_out/s3_cfg${cfg}_pdl${pdl}.json').read().strip().splitlines()[-1]); print('cfg', $cfg, 'pdl', $pdl, round(d['value']), d['ms_per_step'], d['clocks']['sm_mhz'], d['clocks']['reasons'], (d.get('macro_on') or {}).get('value'))"
done; done
