"""Aggregate device-to-host bandwidth of N GPUs copying into pinned host memory AT THE SAME TIME (the transfer bench.py's
e2e leg is bound by): python -m torch.distributed.run --nproc-per-node N tools/d2h_bw.py [MiB per copy] [copies]
Prints one JSON line on rank 0: per-rank and aggregate GB/s, plus the same for host-to-device."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist


def main():
    mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = mib << 20
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    host = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(2)]
    out = {}
    for name, fn in (("d2h", lambda k: host[k % 2].copy_(dev, non_blocking=True)),
                     ("h2d", lambda k: dev.copy_(host[k % 2], non_blocking=True))):
        for k in range(3):
            fn(k)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for k in range(reps):
            fn(k)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        out[name + "_GBps_per_rank"] = n * reps / dt / 1e9
        out[name + "_GBps_aggregate"] = n * reps * world / dt / 1e9
    if rank == 0:
        print(json.dumps({"n_gpus": world, "MiB_per_copy": mib, "copies": reps, **out}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
