#!/usr/bin/env bash
# new single-slab block-height rule (~3 waves, up to 64 rows, last wave at quarter height): parity + numbers
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_frames.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/s3_pytest_tall.log 2>&1; tail -2 gpurun_out/s3_pytest_tall.log
for cfg in 1 2 5 3; do
  python bench.py --config $cfg --steps 200 --warmup 20 --no-secondary --e2e-steps 0 --cpu-seconds 0 > gpurun_out/s3_cfg${cfg}_tall.json 2>gpurun_out/s3_tall.err
  python -c "import json,sys; d=json.loads(open('gpurun_out/s3_cfg${cfg}_tall.json').read().strip().splitlines()[-1]); print('cfg', $cfg, round(d['value']), d['clocks']['sm_mhz'], d['clocks']['reasons'], round((d.get('macro_on') or {}).get('value') or 0), d['gpu_launches'])" || tail -3 gpurun_out/s3_tall.err
done
LBM_FUSE_WAVES=4 python bench.py --config 5 --steps 200 --warmup 20 --no-secondary --e2e-steps 0 --cpu-seconds 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg 5 waves 4', round(d['value']))"
LBM_FUSE_HMAX=32 python bench.py --config 5 --steps 200 --warmup 20 --no-secondary --e2e-steps 0 --cpu-seconds 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg 5 hmax 32', round(d['value']))"
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/s3_bench_default.json 2> gpurun_out/s3_bench_default.err; tail -3 gpurun_out/s3_bench_default.err
python -c "import json; d=json.loads(open('gpurun_out/s3_bench_default.json').read().strip().splitlines()[-1]); print(d['value'], d['macro_on']['value'], d['e2e']['value'], d['secondary']['value'], d['secondary']['macro_on']['value'], d['roofline']['dram_frac'], d['secondary']['roofline']['dram_frac'])"
