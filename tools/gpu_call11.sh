#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/c11_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c11_pytest.log
grep -E "passed|failed|^FAILED|rc=|host time" gpurun_out/c11_pytest.log | tail -30
for cfg in 2 3 5 1; do
    timeout 600 python bench.py --config $cfg --steps 200 --warmup 20 --cpu-seconds 0 --no-secondary > gpurun_out/c11_bench_cfg$cfg.json 2> gpurun_out/c11_bench_cfg$cfg.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c11_bench_cfg$cfg.json").read().strip().splitlines()[-1])
    print("cfg$cfg", round(d["value"]), d["detail"]["kernel"], "macro_on", (d.get("macro_on") or {}).get("value"), "e2e", round(d["e2e"]["value"]), d["e2e"]["two_update_sweeps"], round(d["e2e"]["metric_only_variant"]["value"]), d["clocks"])
except Exception as e:
    print("cfg$cfg FAILED", e); print(open("gpurun_out/c11_bench_cfg$cfg.err").read()[-600:])
PY
done
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/c11_bench_default.json 2> gpurun_out/c11_bench_default.err
tail -4 gpurun_out/c11_bench_default.err
cat gpurun_out/c11_bench_default.json
