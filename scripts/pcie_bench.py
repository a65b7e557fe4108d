"""PCIe micro-benchmark for the end-to-end leg: D2H / H2D of a pinned buffer the size of one step's afterstate output,
alone (one process) or concurrently from every rank (torchrun), so the e2e figure can be stated against a measured
transfer peak instead of a nominal one.

    python scripts/pcie_bench.py [--mb 171] [--reps 10]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/pcie_bench.py
"""
import argparse
import json
import os
import time

import torch

ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=float, default=171.0)
ap.add_argument("--reps", type=int, default=10)
a = ap.parse_args()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nbytes = int(a.mb * 1e6)
host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
out = {"rank": rank, "world": world, "mb": a.mb, "affinity": sorted(os.sched_getaffinity(0))[:4] + ["..."] + [len(os.sched_getaffinity(0))]}
for name, dst, src in (("d2h", host, dev), ("h2d", dev, host)):
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if dist: dist.barrier()
    t0 = time.perf_counter()
    for _ in range(a.reps):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out[name + "_GBps"] = nbytes * a.reps / dt / 1e9
if dist:
    t = torch.tensor([out["d2h_GBps"], out["h2d_GBps"]], device="cuda", dtype=torch.float64)
    lo = t.clone(); dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    sm = t.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    out["min_over_ranks"] = lo.tolist(); out["sum_over_ranks"] = sm.tolist()
if rank == 0:
    print(json.dumps(out))
if dist: dist.destroy_process_group()
