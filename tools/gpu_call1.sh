#!/usr/bin/env bash
# round-2 GPU call 1: parity of the new collide2 + A/B of the instruction cuts + the new bench.py flow
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/c1_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
tail -5 gpurun_out/c1_pytest.log
for v in v00 v10 v01 b200; do
  for cfg in 2 3; do
    LBM_B200_LIB=$PWD/simuverse_b200/_native/liblbm_$v.so timeout 600 python bench.py --config $cfg --steps 200 --warmup 20 --e2e-steps 0 --cpu-seconds 0 --no-secondary > gpurun_out/c1_bench_${v}_cfg$cfg.json 2> gpurun_out/c1_bench_${v}_cfg$cfg.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c1_bench_${v}_cfg$cfg.json").read().strip().splitlines()[-1])
    print("$v cfg$cfg", round(d["value"]), d["detail"]["kernel"], d["clocks"], "macro_on", d.get("macro_on",{}).get("value"))
except Exception as e:
    print("$v cfg$cfg FAILED", e)
PY
  done
done
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/c1_bench_default.json 2> gpurun_out/c1_bench_default.err
tail -3 gpurun_out/c1_bench_default.err
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/c1_bench_ref.json 2> gpurun_out/c1_bench_ref.err
cat gpurun_out/c1_bench_ref.json | cut -c1-400
