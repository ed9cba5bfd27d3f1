#!/usr/bin/env bash
# block height / tail-shaping knobs with the packed last column (200 updates)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { # cfg, label, env...
  local cfg=$1; shift; local label=$1; shift
  env "$@" python bench.py --config $cfg --steps 200 --warmup 20 --no-secondary --e2e-steps 0 --cpu-seconds 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg', $cfg, '$label', round(d['value']), round((d.get('macro_on') or {}).get('value') or 0), d['clocks']['sm_mhz'])"
}
run 2 h32_w2_d4 LBM_FUSE_HMIN=32 LBM_FUSE_TAIL_WAVES=2 LBM_FUSE_TAIL_DIV=4
run 2 h32_w3_d4 LBM_FUSE_HMIN=32 LBM_FUSE_TAIL_WAVES=3 LBM_FUSE_TAIL_DIV=4
run 2 h32_w1_d4 LBM_FUSE_HMIN=32 LBM_FUSE_TAIL_WAVES=1 LBM_FUSE_TAIL_DIV=4
run 2 h32_w2_d8 LBM_FUSE_HMIN=32 LBM_FUSE_TAIL_WAVES=2 LBM_FUSE_TAIL_DIV=8
run 2 h32_w4_d8 LBM_FUSE_HMIN=32 LBM_FUSE_TAIL_WAVES=4 LBM_FUSE_TAIL_DIV=8
run 2 h64_w2_d4 LBM_FUSE_HMIN=64 LBM_FUSE_HMAX=64 LBM_FUSE_TAIL_WAVES=2 LBM_FUSE_TAIL_DIV=4
run 2 h64_w2_d8 LBM_FUSE_HMIN=64 LBM_FUSE_HMAX=64 LBM_FUSE_TAIL_WAVES=2 LBM_FUSE_TAIL_DIV=8
run 2 h32_w2_d4_r16 LBM_FUSE_HMIN=32 LBM_FUSE_TAIL_WAVES=2 LBM_FUSE_TAIL_DIV=4 LBM_FUSE_ROWS0=16
run 5 default X=1
run 5 h32_w2_d4 LBM_FUSE_HMIN=32 LBM_FUSE_TAIL_WAVES=2 LBM_FUSE_TAIL_DIV=4
run 5 w2_d4 LBM_FUSE_TAIL_WAVES=2 LBM_FUSE_TAIL_DIV=4
run 3 default X=1
run 3 w2_d4 LBM_FUSE_TAIL_WAVES=2 LBM_FUSE_TAIL_DIV=4
run 3 h64_w2_d4 LBM_FUSE_HMIN=64 LBM_FUSE_HMAX=64 LBM_FUSE_TAIL_WAVES=2 LBM_FUSE_TAIL_DIV=4
