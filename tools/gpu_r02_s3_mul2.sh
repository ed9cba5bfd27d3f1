#!/usr/bin/env bash
# instruction-count variants on the final geometry: packed multiplies + one-compare clamp together
set -u
cd "$(dirname "$0")/.."
for lib in liblbm_b200.so liblbm_both.so liblbm_clamp.so; do for cfg in 2 3 5; do
  LBM_B200_LIB=$PWD/simuverse_b200/_native/$lib python bench.py --config $cfg --steps 100 --warmup 20 --no-secondary --e2e-steps 0 --cpu-seconds 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib cfg', $cfg, round(d['value']), round((d.get('macro_on') or {}).get('value') or 0), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done; done
LBM_B200_LIB=$PWD/simuverse_b200/_native/liblbm_both.so timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_frames.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -1
