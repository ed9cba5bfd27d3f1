#!/usr/bin/env bash
# 8-GPU box: the driver's own commands (20 steps) at N = 1, 2, 4, 8, longer runs at N = 1 and 8, weak scaling at 8,
# and the cross-process parity tests.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/m_topo.txt 2>&1
run() { # N extra-args tag
  local n=$1; shift; local tag=$1; shift
  if [ "$n" = 1 ]; then
    timeout 900 python bench.py --gpus 1 "$@" > gpurun_out/m_bench_$tag.json 2> gpurun_out/m_bench_$tag.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n "$@" > gpurun_out/m_bench_$tag.json 2> gpurun_out/m_bench_$tag.err
  fi
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/m_bench_$tag.json").read().strip().splitlines() if l.startswith("{")][-1])
    e=d.get("e2e") or {}
    print("$tag: value", round(d["value"]), "per-GPU", round(d["value"]/d["n_gpus"]), "macro_on", round((d.get("macro_on") or {}).get("value") or 0), "e2e", round(e.get("value") or 0), "d2h GB/s/rank", round(e.get("d2h_GBps_per_rank") or 0,1), e.get("host_binding"), "parity", d.get("multirank_parity"), "edge_wait", d.get("edge_wait"), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as ex:
    print("$tag FAILED", ex); print(open("gpurun_out/m_bench_$tag.err").read()[-1500:])
PY
}
for n in 1 2 4 8; do run $n driver_n$n --steps 20 --warmup 5; done
run 1 long_n1 --steps 200 --warmup 20 --cpu-seconds 0 --no-secondary
run 8 long_n8 --steps 200 --warmup 20
run 8 weak_n8 --steps 100 --warmup 10 --config 4 --e2e-steps 0
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q > gpurun_out/m_pytest.log 2>&1; tail -3 gpurun_out/m_pytest.log
