#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for lib in b200 nomask; do
  LBM_B200_LIB=$PWD/simuverse_b200/_native/liblbm_$lib.so python tools/repro_porous.py 8192 8192 0 > gpurun_out/c6_repro_$lib.log 2>&1; echo "$lib: $(tail -1 gpurun_out/c6_repro_$lib.log | cut -c1-200)"
done
timeout 1200 compute-sanitizer --tool memcheck --print-limit 3 python tools/repro_porous.py 8192 8192 0 > gpurun_out/c6_sanitizer_8192.log 2>&1; grep -E "Invalid|at |by thread|ERROR SUMMARY|Address|access" gpurun_out/c6_sanitizer_8192.log | head -40
for lib in b200 nomask; do
  for cfg in 2 3 5 1; do
    LBM_B200_LIB=$PWD/simuverse_b200/_native/liblbm_$lib.so timeout 600 python bench.py --config $cfg --steps 100 --warmup 10 --e2e-steps 0 --cpu-seconds 0 --no-secondary > gpurun_out/c6_bench_${lib}_cfg$cfg.json 2> gpurun_out/c6_bench_${lib}_cfg$cfg.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c6_bench_${lib}_cfg$cfg.json").read().strip().splitlines()[-1])
    print("$lib cfg$cfg", round(d["value"]), d["detail"]["kernel"], "macro_on", (d.get("macro_on") or {}).get("value"), d["clocks"])
except Exception as e:
    print("$lib cfg$cfg FAILED", e); print(open("gpurun_out/c6_bench_${lib}_cfg$cfg.err").read()[-600:])
PY
  done
done
