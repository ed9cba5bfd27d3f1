#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for bx in 16 4 8 32; do LBM_PARTICLE_BLOCK_X=$bx python tools/time_particles.py 2>&1 | tail -1; done
b() { python bench.py "$@" --steps 100 --warmup 10 --e2e-steps 0 --cpu-seconds 0 --no-secondary 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round((d.get('macro_on') or {}).get('value') or 0))"; }
for wv in "1.0 2" "2.0 2" "0.5 2" "1.0 4" "1.5 2" "3.0 2"; do
  set -- $wv
  export LBM_FUSE_TAIL_WAVES=$1 LBM_FUSE_TAIL_DIV=$2
  echo "tail waves=$1 div=$2: cfg2 $(b --config 2) | 16384x2048 $(b --lattice 16384 2048) | 8192^2 $(b --lattice 8192 8192) | cfg2 again $(b --config 2)"
done
