#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/r02_d2h_bandwidth.jsonl
python tools/d2h_bw.py 256 20 >> gpurun_out/r02_d2h_bandwidth.jsonl 2>/dev/null
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29612 tools/d2h_bw.py 256 20 2>/dev/null | grep '^{' >> gpurun_out/r02_d2h_bandwidth.jsonl
done
cat gpurun_out/r02_d2h_bandwidth.jsonl
nproc; free -g | head -2; lscpu | grep -E "Model name|Socket|NUMA" 
