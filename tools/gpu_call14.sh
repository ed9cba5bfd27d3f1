#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_frames.py tests/test_gpu_fullsize.py -m gpu -q > gpurun_out/c14_pytest.log 2>&1; tail -3 gpurun_out/c14_pytest.log
for m in auto 0 1; do
  for cfg in 2 3 5 1; do
    if [ $m = auto ]; then unset LBM_FUSE_MASKED; else export LBM_FUSE_MASKED=$m; fi
    timeout 600 python bench.py --config $cfg --steps 200 --warmup 20 --e2e-steps 0 --cpu-seconds 0 --no-secondary > gpurun_out/c14_bench_m${m}_cfg$cfg.json 2> gpurun_out/c14_bench_m${m}_cfg$cfg.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c14_bench_m${m}_cfg$cfg.json").read().strip().splitlines()[-1])
    print("masked=$m cfg$cfg", round(d["value"]), "macro_on", round((d.get("macro_on") or {}).get("value") or 0), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("masked=$m cfg$cfg FAILED", e); print(open("gpurun_out/c14_bench_m${m}_cfg$cfg.err").read()[-600:])
PY
  done
done
