#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
b() { python bench.py "$@" --steps 100 --warmup 10 --e2e-steps 0 --cpu-seconds 0 --no-secondary 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round((d.get('macro_on') or {}).get('value') or 0), d['clocks']['sm_mhz'])"; }
for lib in b200 st1c5 st1c6 st1c7; do
  export LBM_B200_LIB=$PWD/simuverse_b200/_native/liblbm_$lib.so
  echo "$lib: cfg2 $(b --config 2) | cfg3 $(b --config 3) | cfg5 $(b --config 5) | cfg1 $(b --config 1) | 16384x2048 $(b --lattice 16384 2048)"
done
for lib in st1c6 st1c7; do
  export LBM_B200_LIB=$PWD/simuverse_b200/_native/liblbm_$lib.so
  timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_frames.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -2
done
