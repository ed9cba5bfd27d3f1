#!/usr/bin/env bash
# 8-GPU box, final: host D2H ceiling, then the driver's own commands at N = 1, 2, 4, 8 and a longer run at N = 8
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/r02_d2h_bandwidth.jsonl
python tools/d2h_bw.py 256 20 2>/dev/null | grep '^{' >> gpurun_out/r02_d2h_bandwidth.jsonl
for n in 2 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29612 tools/d2h_bw.py 256 20 2>/dev/null | grep '^{' >> gpurun_out/r02_d2h_bandwidth.jsonl
done
cat gpurun_out/r02_d2h_bandwidth.jsonl
run() {
  local n=$1; shift; local tag=$1; shift
  if [ "$n" = 1 ]; then
    timeout 900 python bench.py --gpus 1 "$@" > gpurun_out/m2_bench_$tag.json 2> gpurun_out/m2_bench_$tag.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n "$@" > gpurun_out/m2_bench_$tag.json 2> gpurun_out/m2_bench_$tag.err
  fi
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/m2_bench_$tag.json").read().strip().splitlines() if l.startswith("{")][-1])
    e=d.get("e2e") or {}
    print("$tag: value", round(d["value"]), "per-GPU", round(d["value"]/d["n_gpus"]), "macro_on", round((d.get("macro_on") or {}).get("value") or 0), "e2e", round(e.get("value") or 0), "d2h GB/s/rank", round(e.get("d2h_GBps_per_rank") or 0,1), "metric-only", round((e.get("metric_only_variant") or {}).get("value") or 0), "parity", d.get("multirank_parity"), "edge_wait", (d.get("edge_wait") or {}).get("mean_us_per_wait"), d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "dram_frac", d["roofline"]["dram_frac"])
except Exception as ex:
    print("$tag FAILED", ex); print(open("gpurun_out/m2_bench_$tag.err").read()[-1500:])
PY
}
for n in 1 2 4; do run $n driver_n$n --steps 20 --warmup 5; done
