#!/usr/bin/env bash
# a slab-sized lattice (16384 x 2048 = one of 8 slabs) on one GPU: block-height rules
set -u
cd "$(dirname "$0")/.."
run() { local label=$1; shift
  env "$@" python bench.py --lattice 16384 2048 --steps 200 --warmup 20 --no-secondary --e2e-steps 0 --cpu-seconds 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$label', round(d['value']), round((d.get('macro_on') or {}).get('value') or 0), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
}
run new_rule X=1
run old_rule LBM_FUSE_WAVES=6 LBM_FUSE_HMAX=32 LBM_FUSE_TAIL_DIV=2 LBM_FUSE_PACK=0
run old_rule_pack LBM_FUSE_WAVES=6 LBM_FUSE_HMAX=32 LBM_FUSE_TAIL_DIV=2
run h32_div4_nopack LBM_FUSE_HMAX=32 LBM_FUSE_PACK=0
run h64_div4_nopack LBM_FUSE_PACK=0
run h64_div2_nopack LBM_FUSE_PACK=0 LBM_FUSE_TAIL_DIV=2
run h64_w2_div4_nopack LBM_FUSE_PACK=0 LBM_FUSE_TAIL_WAVES=2
run h32_w2_div4_nopack LBM_FUSE_HMAX=32 LBM_FUSE_PACK=0 LBM_FUSE_TAIL_WAVES=2
