#!/usr/bin/env bash
# AA in-place variant: register budget A/B (7 CTAs per SM = 72 registers with spills ... 4 CTAs = 115 registers, none)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for lib in liblbm_b200.so liblbm_aa6.so liblbm_aa5.so liblbm_aa4.so; do for cfg in 2 3; do
  LBM_B200_LIB=$PWD/simuverse_b200/_native/$lib python bench.py --config $cfg --aa --steps 200 --warmup 20 --no-secondary --e2e-steps 0 --cpu-seconds 0 > gpurun_out/s3_aa_${lib%.so}_cfg$cfg.json 2>gpurun_out/s3_aa.err
  python -c "import json,sys; d=json.loads(open('gpurun_out/s3_aa_${lib%.so}_cfg$cfg.json').read().strip().splitlines()[-1]); print('$lib cfg', $cfg, round(d['value']), d['clocks']['sm_mhz'], d['clocks']['reasons'], d['detail'].get('kernel'))" || tail -3 gpurun_out/s3_aa.err
done; done
python bench.py --config 2 --no-fuse --steps 200 --warmup 20 --no-secondary --e2e-steps 0 --cpu-seconds 0 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('A/B single update cfg 2', round(d['value']))"
