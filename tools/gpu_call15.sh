#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/c15_pytest.log 2>&1; tail -3 gpurun_out/c15_pytest.log
for t in 1 0; do
  for cfg in 2 3 5; do
    LBM_FUSE_TAIL=$t timeout 600 python bench.py --config $cfg --steps 100 --warmup 10 --e2e-steps 0 --cpu-seconds 0 --no-secondary > gpurun_out/c15_bench_t${t}_cfg$cfg.json 2> gpurun_out/c15_bench_t${t}_cfg$cfg.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c15_bench_t${t}_cfg$cfg.json").read().strip().splitlines()[-1])
    print("tail=$t cfg$cfg", round(d["value"]), "macro_on", round((d.get("macro_on") or {}).get("value") or 0), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("tail=$t cfg$cfg FAILED", e); print(open("gpurun_out/c15_bench_t${t}_cfg$cfg.err").read()[-600:])
PY
  done
  LBM_FUSE_TAIL=$t python tools/slabs_one_gpu.py 16384 16384 8 40 | tail -1
  LBM_FUSE_TAIL=$t python bench.py --lattice 16384 2048 --steps 100 --warmup 10 --e2e-steps 0 --cpu-seconds 0 --no-secondary 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('tail=$t 16384x2048 single', round(d['value']))"
  LBM_FUSE_TAIL=$t python bench.py --lattice 8192 8192 --steps 100 --warmup 10 --e2e-steps 0 --cpu-seconds 0 --no-secondary 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('tail=$t 8192x8192 channel', round(d['value']))"
done
for lib in cta6 w3; do for cfg in 2 3 5; do LBM_B200_LIB=$PWD/simuverse_b200/_native/liblbm_$lib.so timeout 600 python bench.py --config $cfg --steps 100 --warmup 10 --e2e-steps 0 --cpu-seconds 0 --no-secondary 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(\"$lib cfg$cfg\", round(d[\"value\"]))"; done; done
