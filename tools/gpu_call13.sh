#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for lib in b200 vote; do
  for cfg in 5 1; do
    LBM_B200_LIB=$PWD/simuverse_b200/_native/liblbm_$lib.so timeout 600 python bench.py --config $cfg --steps 200 --warmup 20 --e2e-steps 0 --cpu-seconds 0 --no-secondary > gpurun_out/c13_bench_${lib}_cfg$cfg.json 2> gpurun_out/c13_bench_${lib}_cfg$cfg.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c13_bench_${lib}_cfg$cfg.json").read().strip().splitlines()[-1])
    print("$lib cfg$cfg", round(d["value"]), d["detail"]["kernel"], d["clocks"])
except Exception as e:
    print("$lib cfg$cfg FAILED", e); print(open("gpurun_out/c13_bench_${lib}_cfg$cfg.err").read()[-600:])
PY
  done
done
for lib in b200 vote; do
LBM_B200_LIB=$PWD/simuverse_b200/_native/liblbm_$lib.so LBM_FUSE_MASKED=1 timeout 600 python bench.py --config 2 --steps 200 --warmup 20 --e2e-steps 0 --cpu-seconds 0 --no-secondary > gpurun_out/c13_bench_${lib}_cfg2m.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/c13_bench_${lib}_cfg2m.json').read().strip().splitlines()[-1]); print('$lib cfg2 forced masked', round(d['value']))"
done
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_frames.py tests/test_gpu_fullsize.py -m gpu -q > gpurun_out/c13_pytest.log 2>&1; tail -3 gpurun_out/c13_pytest.log
