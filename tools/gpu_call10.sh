#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for args in "8192 8192 0 steps" "4096 256 0 steps" "8192 8192 1 frames"; do
  python tools/repro_porous2.py $args > gpurun_out/c10_repro.log 2>&1; echo "[$args] $(grep -E 'ok|created|mass|Error' gpurun_out/c10_repro.log | tail -2 | tr '\n' ' ' | cut -c1-200)"
done
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/c10_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c10_pytest.log
grep -E "passed|failed|^FAILED|rc=|host time" gpurun_out/c10_pytest.log | tail -30
for lib in b200 nomask; do
  for cfg in 2 3 5 1; do
    LBM_B200_LIB=$PWD/simuverse_b200/_native/liblbm_$lib.so timeout 600 python bench.py --config $cfg --steps 100 --warmup 10 --e2e-steps 0 --cpu-seconds 0 --no-secondary > gpurun_out/c10_bench_${lib}_cfg$cfg.json 2> gpurun_out/c10_bench_${lib}_cfg$cfg.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c10_bench_${lib}_cfg$cfg.json").read().strip().splitlines()[-1])
    print("$lib cfg$cfg", round(d["value"]), d["detail"]["kernel"], "macro_on", (d.get("macro_on") or {}).get("value"), d["clocks"])
except Exception as e:
    print("$lib cfg$cfg FAILED", e); print(open("gpurun_out/c10_bench_${lib}_cfg$cfg.err").read()[-600:])
PY
  done
done
