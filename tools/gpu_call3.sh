#!/usr/bin/env bash
# round-2 GPU call: full GPU test suite + bench of every config with the sweep on the frame loop
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/c4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c4_pytest.log
tail -15 gpurun_out/c4_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c4_smoke.log 2>&1; tail -3 gpurun_out/c4_smoke.log
for cfg in 1 2 5; do
  timeout 600 python bench.py --config $cfg --steps 200 --warmup 20 --cpu-seconds 0 --no-secondary > gpurun_out/c4_bench_cfg$cfg.json 2> gpurun_out/c4_bench_cfg$cfg.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c4_bench_cfg$cfg.json").read().strip().splitlines()[-1])
    print("cfg$cfg", round(d["value"]), d["detail"]["kernel"], "launches", d["gpu_launches"], "macro_on", (d.get("macro_on") or {}).get("value"), "e2e", round(d["e2e"]["value"]), d["e2e"].get("two_update_sweeps"), round(d["e2e"]["metric_only_variant"]["value"]))
except Exception as e:
    print("cfg$cfg FAILED", e); print(open("gpurun_out/c4_bench_cfg$cfg.err").read()[-2000:])
PY
done
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/c4_bench_default.json 2> gpurun_out/c4_bench_default.err
tail -4 gpurun_out/c4_bench_default.err
python - <<PY
import json
d=json.loads(open("gpurun_out/c4_bench_default.json").read().strip().splitlines()[-1])
print("default", round(d["value"]), "macro_on", d["macro_on"]["value"], "e2e", d["e2e"]["value"], d["e2e"]["two_update_sweeps"], "secondary", d["secondary"]["value"], d["secondary"]["macro_on"])
PY
