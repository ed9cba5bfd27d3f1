#!/usr/bin/env bash
# session 3 of round 2: the whole GPU suite, smoke, the driver's bench command and its reference arm on the final library
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/s3_pytest.log 2>&1; tail -3 gpurun_out/s3_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/s3_bench_default.json 2> gpurun_out/s3_bench_default.err; tail -3 gpurun_out/s3_bench_default.err
python -c "import json; d=json.loads(open('gpurun_out/s3_bench_default.json').read().strip().splitlines()[-1]); print(d['value'], d['macro_on']['value'], d['e2e']['value'], d['e2e']['d2h_GBps_per_rank'], d['secondary']['value'], d['secondary']['macro_on']['value'], d['roofline']['dram_frac'], d['secondary']['roofline']['dram_frac'], d['clocks'])"
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/s3_bench_ref.json 2>/dev/null; cut -c1-250 gpurun_out/s3_bench_ref.json
for cfg in 1 5; do python bench.py --config $cfg --steps 200 --warmup 20 --no-secondary --e2e-steps 0 --cpu-seconds 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg', $cfg, round(d['value']))"; done
