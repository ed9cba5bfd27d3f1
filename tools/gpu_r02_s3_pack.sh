#!/usr/bin/env bash
# CTAs of a partially filled last strip column take several row blocks (-DLBM_FUSE_PACK=1): parity, then A/B
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export LBM_B200_LIB=$PWD/simuverse_b200/_native/liblbm_pack.so
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_frames.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/s3_pytest_pack.log 2>&1; tail -2 gpurun_out/s3_pytest_pack.log
for sp in 0 1; do for cfg in 1 2 5 3; do
  LBM_FUSE_PACK=$sp python bench.py --config $cfg --steps 200 --warmup 20 --no-secondary --e2e-steps 0 --cpu-seconds 0 > gpurun_out/s3_cfg${cfg}_pack${sp}.json 2>gpurun_out/s3_pack.err
  python -c "import json,sys; d=json.loads(open('gpurun_out/s3_cfg${cfg}_pack${sp}.json').read().strip().splitlines()[-1]); print('cfg', $cfg, 'pack', $sp, round(d['value']), d['clocks']['sm_mhz'], d['clocks']['reasons'], round((d.get('macro_on') or {}).get('value') or 0), d['gpu_launches'])" || tail -3 gpurun_out/s3_pack.err
done; done
